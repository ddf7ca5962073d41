"""Short profiling target: build one workload of bench.py and run a few solves (for `ncu ... python tools/ncu_target.py c5 4`)."""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mppi_numba_b200 as E          # noqa: E402
from bench import build_scenario     # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = build_scenario(name)
with contextlib.redirect_stdout(io.StringIO()):
    cfg = E.Config(**sc["cfg"])
    lin, ang = E.TDM_Numba(cfg), E.TDM_Numba(cfg)
    lin.set_TDM_from_PMF_grid(sc["pmf_lin"], sc["tdm_dict"], sc["obstacle"], sc["unknown"])
    ang.set_TDM_from_PMF_grid(sc["pmf_ang"], sc["tdm_dict"], sc["obstacle"], sc["unknown"])
    pl = E.MPPI_Numba(cfg)
    pl.setup(sc["params"], lin, ang)
pl.set_profiling(True)
for k in range(n):
    u = pl.solve()
    print(k, pl.last_timings())
