"""Stage times of ONE rank of a G-rank sharded solve, run alone on one GPU (no exchange: solve_local only, staged
cost layout): separates what a rank's kernels cost from what the exchange / the other ranks add.
    python tools/rank_stage_times.py [c5] [G ...]"""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                   # noqa: E402
import mppi_numba_b200 as E          # noqa: E402
from mppi_numba_b200._lib import lib, check   # noqa: E402
from bench import build_scenario     # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
gs = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
sc = build_scenario(name)
for G in gs:
    with contextlib.redirect_stdout(io.StringIO()):
        cfg = E.Config(**sc["cfg"])
        lin, ang = E.TDM_Numba(cfg, rank=0, world_size=G), E.TDM_Numba(cfg, rank=0, world_size=G)
        lin.set_TDM_from_PMF_grid(sc["pmf_lin"], sc["tdm_dict"], sc["obstacle"], sc["unknown"])
        ang.set_TDM_from_PMF_grid(sc["pmf_ang"], sc["tdm_dict"], sc["obstacle"], sc["unknown"])
        pl = E.MPPI_Numba(cfg, rank=0, world_size=G)
        pl.setup(sc["params"], lin, ang)
    pl.move_mppi_task_vars_to_device()
    pl.set_profiling(True)
    acc = {}
    for k in range(12):
        if G == 1:
            pl.solve()
        else:
            check(lib.b200mppi_planner_solve_local(pl._handle, 1))
            check(lib.b200mppi_planner_synchronize(pl._handle))
        if k >= 4:
            import ctypes as C
            ms = (C.c_float * 6)()
            check(lib.b200mppi_planner_last_timings(pl._handle, ms))
            for nm, v in zip(("sample", "noise", "rollout"), (ms[0], ms[1], ms[2])):
                acc.setdefault(nm, []).append(float(v))
    print("G", G, {k: round(float(np.median(v)), 4) for k, v in acc.items()}, flush=True)
    del pl, lin, ang
