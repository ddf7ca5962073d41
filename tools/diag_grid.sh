#!/bin/bash
# per-CTA speed of the rollout kernel against the number of busy SMs (debug-hook build)
cp mppi_numba_b200/ab/lib_dbg.so mppi_numba_b200/libb200mppi.so
for g in 4 8 16 32 64 128 148; do
  echo "== grid $g"
  B200MPPI_WIN_GRID=$g B200MPPI_WIN_UNIT=1 python tools/rollout_cta_times.py c5 8 2>&1 | grep -E "kernel span|warp-steps per CTA"
done
