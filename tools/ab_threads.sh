#!/bin/bash
# A/B: resident warps of the windowed rollout kernel for short shares (one rank of 8 / 4 alone on one GPU)
for rep in 1 2; do for t in 0 896 768; do echo "== B200MPPI_WIN_THREADS=$t"; B200MPPI_WIN_THREADS=$t python tools/rank_stage_times.py c5 8 4 2>&1 | tail -2; done; done
