#!/bin/bash
# DIAG: rollout kernel time against the way the per-lane control stream (16 B per thread-step from L2) is loaded
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pair_probe tools/pair_probe.cu && /tmp/pair_probe | tail -3
for s in 0 -1 -2 -3 -4 -5 -6 -7; do
  echo "== B200MPPI_WIN_STAGGER=$s"
  B200MPPI_WIN_STAGGER=$s python bench.py --steps 30 --warmup 5 --no-numba --no-others 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline'].get('stage_ms'))"
  B200MPPI_WIN_STAGGER=$s python tools/rollout_cta_times.py c5 8 2>&1 | tail -1
done
