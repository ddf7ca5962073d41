#!/bin/bash
# A/B at G ranks: one system fence per CTA against one per thread in the rollout kernel's epilogue
g=${1:-2}
for rep in 1 2; do for s in 0 -1; do
  echo "== B200MPPI_WIN_STAGGER=$s"
  B200MPPI_WIN_STAGGER=$s python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29600 + rep * 4 + s + 1)) bench.py --gpus $g --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],4), d['parity_check']['passed'], {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})"
done; done
