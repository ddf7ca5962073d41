#!/bin/bash
# A/B: resident warps per SM of the windowed rollout kernel
for t in 0 512 640 768 896; do
  echo "== B200MPPI_WIN_THREADS=$t"
  B200MPPI_WIN_THREADS=$t python bench.py --steps 30 --warmup 5 --no-numba --no-others 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline'].get('stage_ms'))"
  for g in 8 4; do B200MPPI_WIN_THREADS=$t python tools/rollout_cta_times.py c5 $g 2>&1 | grep -E "kernel span|warp-steps per CTA"; done
done
