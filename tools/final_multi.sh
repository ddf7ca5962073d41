#!/bin/bash
# Multi-GPU evidence run (gpurun --gpus N): bench at every G <= N given on the command line.
# usage: tools/final_multi.sh <tag> 2 4 8
tag=$1; shift
for g in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g)) bench.py --gpus $g --steps 20 --warmup 5 > gpurun_out/${tag}_bench_g${g}.json 2> gpurun_out/${tag}_bench_g${g}.err
  echo "G=$g rc=$?"; tail -c 1800 gpurun_out/${tag}_bench_g${g}.json; echo; tail -3 gpurun_out/${tag}_bench_g${g}.err
done
