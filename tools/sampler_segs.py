"""Sweep the sampler's row-segment count (B200MPPI_SAMPLE_SEGS) for the map counts a rank holds at 2/4/8 GPUs.
usage: python tools/sampler_segs.py   (on a GPU box)"""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mppi_numba_b200 as E
from bench import build_scenario

sc = build_scenario("c5")
for M in (32, 64, 128, 256):
    row = []
    for segs in (0, 4, 5, 6, 7, 8, 9, 10, 11, 13, 16, 22, 33):
        if segs:
            os.environ["B200MPPI_SAMPLE_SEGS"] = str(segs)
        else:
            os.environ.pop("B200MPPI_SAMPLE_SEGS", None)
        with contextlib.redirect_stdout(io.StringIO()):
            c = dict(sc["cfg"]); c["num_grid_samples"] = M; c["num_control_rollouts"] = 1024
            cfg = E.Config(**c); lin, ang = E.TDM_Numba(cfg), E.TDM_Numba(cfg)
            lin.set_TDM_from_PMF_grid(sc["pmf_lin"], sc["tdm_dict"], sc["obstacle"], sc["unknown"])
            ang.set_TDM_from_PMF_grid(sc["pmf_ang"], sc["tdm_dict"], sc["obstacle"], sc["unknown"])
            pl = E.MPPI_Numba(cfg); pl.setup(sc["params"], lin, ang)
        pl.set_profiling(True)
        ts = []
        for k in range(6):
            pl.solve()
            ts.append(pl.last_timings()["sample_grids"])
        row.append("%s:%.3f" % (segs if segs else "auto", min(ts[1:])))
        del pl, lin, ang
    print("M=%d  " % M + "  ".join(row), flush=True)
