"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN KERNELS (unmodified, imported from
/root/reference) under Numba's CUDA simulator in the build container.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden            # ~2-3 minutes, writes tests/golden/ref_*.npz

The reference has no tests or golden vectors of its own; these fixtures are what
pins the oracle (tests/test_oracle_golden.py) and, through it, the CUDA engine.  Simulator caveats
(SURVEY.md 8c): update_useq_numba is launched [1,1] (race 9-R1), costs_d is snapshotted between
kernels (9-Q1), sampled-grid VALUES are only compared for bin values that are multiples of 1/4
(NEP-50 vs compiled typing, 8c-iv).  Sizes are tiny: the simulator runs ~40 threads/s.
"""
import os
import sys
import io
import contextlib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        return fn(*a, **k)


def random_pmf(rng, B, H, W, zero_frac=0.3):
    """int (B,H,W) PMF in percent summing to 100 per cell, with some empty bins."""
    cuts = np.sort(rng.integers(0, 101, (B - 1, H, W)), axis=0)
    pmf = np.empty((B, H, W), dtype=np.int64)
    pmf[0] = cuts[0]
    pmf[1:B - 1] = cuts[1:] - cuts[:-1]
    pmf[B - 1] = 100 - cuts[B - 2]
    # move the mass of some bins into the last bin to create zero-probability bins
    kill = rng.random((B - 1, H, W)) < zero_frac
    moved = np.where(kill, pmf[:B - 1], 0)
    pmf[:B - 1] -= moved
    pmf[B - 1] += moved.sum(axis=0)
    assert (pmf.sum(axis=0) == 100).all() and (pmf >= 0).all()
    return pmf


def base_params(x0, xgoal, cvar_alpha=0.5):
    return dict(dt=0.1, x0=np.asarray(x0, dtype=float), xgoal=np.asarray(xgoal, dtype=float),
                goal_tolerance=0.5, v_post_rollout=0.01, cvar_alpha=cvar_alpha, alpha_dyn=1.0,
                dist_weight=1.0, lambda_weight=1.0, num_opt=1,
                u_std=np.array([2.0, 3.0]), vrange=np.array([0.0, 3.0]), wrange=np.array([-np.pi, np.pi]),
                obs_penalty=1e5, unknown_penalty=1e2)


class _T(object):
    """Stand-in for the reference's Terrain objects: the setter only uses them as dictionary keys."""
    def __init__(self, name):
        self.name = name


def semantic_inputs():
    rng = np.random.default_rng(23)
    H, W, B = 9, 7, 6
    sg = rng.integers(0, 3, (H, W))
    bin_values = np.linspace(0.0, 1.0, B)               # float64 on purpose (the reference uploads it uncast)
    names = {0: "grass", 1: "mud", 2: "road"}
    pmfs = {"grass": np.array([0.05, 0.1, 0.2, 0.4, 0.2, 0.05]), "mud": np.array([0.3, 0.3, 0.2, 0.1, 0.1, 0.0]),
            "road": np.array([0.0, 0.0, 0.02, 0.08, 0.3, 0.6])}
    obstacle = (rng.random((H, W)) < 0.1).astype(np.int8)
    unknown = (rng.random((H, W)) < 0.1).astype(np.int8)
    return sg, bin_values, names, pmfs, obstacle, unknown


def make_semantic_golden(Config, TDM_Numba):
    sg, bin_values, names, pmfs, obstacle, unknown = semantic_inputs()
    terr = {n: _T(n) for n in names.values()}
    t2p = {terr[n]: (bin_values, pmfs[n]) for n in terr}
    out = dict(sg=sg, bin_values=bin_values, obstacle=obstacle, unknown=unknown)
    for mode, flags, alphas in (("tdm", dict(use_tdm=True), (None,)), ("det", dict(use_det_dynamics=True), (0.3, 1.0)),
                                ("spd", dict(use_nom_dynamics_with_speed_map=True), (0.3, 1.0))):
        for alpha in alphas:
            cfg = _quiet(Config, T=1.0, dt=0.1, num_grid_samples=2, num_control_rollouts=100, seed=1,
                         max_map_dim=(14, 12), tdm_sample_thread_dim=(3, 2), max_speed_padding=5.0, **flags)
            tdm = _quiet(TDM_Numba, cfg)
            _quiet(tdm.set_TDM_from_semantic_grid, sg, 0.5, len(bin_values), bin_values, np.array([0.0, 1.0]),
                   np.array([0.0, 3.5]), np.array([0.0, 4.5]), names, terr, t2p, det_dynamics_cvar_alpha=alpha,
                   obstacle_map=obstacle, unknown_map=unknown)
            key = "%s_%s" % (mode, "none" if alpha is None else "a%02d" % int(alpha * 10))
            out[key + "_pmf_padded"] = tdm.pmf_grid_d.copy_to_host()
            out[key + "_semantic_cropped"] = np.asarray(tdm.semantic_grid)
            if mode == "spd":
                out[key + "_risk"] = tdm.risk_traction_map_d.copy_to_host()
            # sampled maps: float64 bin values are uploaded uncast here, so the simulator's float64
            # quantisation equals the compiled one (no NEP-50 caveat on this path)
            tdm.sample_grid_batch_d.copy_to_device(np.zeros(tdm.sample_grid_batch_d.shape, dtype=np.int8))
            out[key + "_grid1"] = _quiet(tdm.sample_grids, 0.9).copy_to_host().copy()
            print("semantic", key, "done")
    np.savez_compressed(os.path.join(OUT, "ref_semantic.npz"), **out)


def load_barebone_notebook():
    """Executes the code cells of the reference's barebone_mppi_numba.ipynb (its own Config / MPPI_Numba with
    the map-free kernels, cells 2-3) in a namespace, under the simulator shims installed by load_reference()."""
    import json
    from oracle.ref_loader import REFERENCE_ROOT
    nb = json.load(open(os.path.join(REFERENCE_ROOT, "barebone_mppi_numba.ipynb")))
    ns = {}
    exec("import numpy as np\nimport math\nimport copy\nimport numba\nimport time\nfrom numba import cuda\n"
         "from numba.cuda.random import create_xoroshiro128p_states, xoroshiro128p_normal_float32\n", ns)
    code = [c for c in nb["cells"] if c["cell_type"] == "code"]
    exec("".join(code[1]["source"]), ns)          # Config (barebone_mppi_numba.ipynb cell 2)
    exec("".join(code[2]["source"]), ns)          # stage/term cost + MPPI_Numba with 4 kernels (cell 3)
    return ns["Config"], ns["MPPI_Numba"]


def make_barebone_golden(cuda):
    BConfig, BMPPI = load_barebone_notebook()
    f32 = np.float32
    rng = np.random.default_rng(29)
    N, T = 40, 14
    noise = (rng.standard_normal((N, T, 2)) * np.array([1.0, 1.0])).astype(f32)
    u_cur = np.stack([rng.uniform(0.5, 1.8, T), rng.uniform(-0.6, 0.6, T)], 1).astype(f32)
    x0 = np.array([0.0, 0.0, np.pi / 4], dtype=f32)
    obs_pos = np.array([[1.4, 1.2], [0.6, 2.0], [2.4, 0.4]], dtype=f32)
    obs_r = np.array([0.5, 0.4, 0.3], dtype=f32)
    out = dict(noise=noise, u_cur=u_cur, x0=x0, obs_pos=obs_pos, obs_r=obs_r, vrange=[0.0, 2.0],
               wrange=[-np.pi, np.pi], u_std=[1.0, 1.0], dt=0.1, lam=1.0, goal_tol=0.5, dist_weight=10,
               obs_cost=1e6, goal_near=np.array([1.3, 1.5], f32), goal_far=np.array([7.0, 5.0], f32))
    dev = cuda.to_device
    for gname in ("near", "far"):
        costs_d = cuda.device_array((N,), dtype=f32)
        BMPPI.rollout_numba[N, 1](dev(np.array([0, 2], f32)), dev(np.array([-np.pi, np.pi], f32)), dev(out["goal_" + gname]),
                                  f32(1e6), dev(obs_pos), dev(obs_r), f32(0.5), f32(1.0), dev(np.array([1, 1], f32)),
                                  dev(x0), f32(0.1), 10, dev(noise), dev(u_cur), costs_d)
        out["costs_" + gname] = costs_d.copy_to_host()
    # whole solve through the notebook's public API (update launched with one thread)
    cfg = _quiet(BConfig, T=1.0, dt=0.1, num_control_rollouts=100, num_vis_state_rollouts=5, seed=1)
    pl = _quiet(BMPPI, cfg)
    params = dict(dt=0.1, x0=np.array([0.0, 0.0, np.pi / 4]), xgoal=np.array([7.0, 5.0]), goal_tolerance=0.5,
                  dist_weight=10, lambda_weight=1.0, num_opt=1, u_std=np.array([1.0, 1.0]),
                  vrange=np.array([0.0, 2.0]), wrange=np.array([-np.pi, np.pi]),
                  obstacle_positions=np.array([[5, 4.5], [2, 1]]), obstacle_radius=np.array([1.5, 1.0]), obs_penalty=1e6)
    pl.setup(params)
    orig = BMPPI.update_useq_numba

    class _One:
        def __getitem__(self, cfg_):
            return orig[1, 1]
    pl.update_useq_numba = _One()
    u1 = _quiet(pl.solve).copy()
    out["solve_u1"] = u1
    out["solve_noise1"] = pl.noise_samples_d.copy_to_host().copy()
    out["solve_states1"] = _quiet(pl.get_state_rollout).copy()
    pl.shift_and_update(np.array([0.05, 0.06, 0.8]), u1, num_shifts=1)
    out["solve_u2"] = _quiet(pl.solve).copy()
    np.savez_compressed(os.path.join(OUT, "ref_barebone.npz"), **out)
    print("barebone done")


def make_oversized_golden(MPPI_Numba, cuda):
    """rollout_oversized_numba (mppi.py:760-913), the kernel behind solve_stochastic_oversized, on the inputs
    of ref_rollout.npz.  Launched with FEWER threads than maps (4 threads, 6 maps -> 2 maps per thread), which
    is exactly the regime the reference uses it in (num_grid_samples > 1024 threads).  alpha = 1 is its
    meaningful output (the mean).  alpha < 1 cannot be recorded: besides swapping unconditionally
     its "sort" indexes thread_cost_shared[tid + ri*num_threads] without bounding it by M
    (mppi.py:879-895) -- the simulator raises IndexError, a GPU reads past the array."""
    f32 = np.float32
    g = np.load(os.path.join(OUT, "ref_rollout.npz"))
    dev = cuda.to_device
    N = g["noise"].shape[0]
    M = g["lin"].shape[0]
    out = {}
    for gname in ("near", "far"):
        goal = g["xgoal_" + gname]
        for alpha in (1.0,):
            costs_d = cuda.device_array((N,), dtype=f32)
            MPPI_Numba.rollout_oversized_numba[N, 4, 0, 4 * M](
                dev(g["lin"]), dev(g["ang"]), dev(np.array([0, 1], f32)), dev(np.array([0, 1], f32)), dev(g["obs"]),
                dev(g["unk"]), f32(g["res"]), dev(g["xlim"]), dev(g["ylim"]), dev(np.array([0, 3], f32)),
                dev(np.array([-np.pi, np.pi], f32)), dev(goal), f32(0.01), f32(1e5), f32(1e2), f32(0.5), f32(1.0),
                dev(np.array([2, 3], f32)), f32(alpha), dev(g["x0"]), f32(0.1), 1.0, dev(g["noise"]), dev(g["u_cur"]),
                costs_d)
            out["over_a%02d_%s" % (int(alpha * 10), gname)] = costs_d.copy_to_host()
        print("oversized", gname, "done")
    np.savez_compressed(os.path.join(OUT, "ref_oversized.npz"), threads=4, **out)


def make_state_rollout_golden(Config, TDM_Numba, MPPI_Numba):
    """get_state_rollout (mppi.py:545-608) and its two kernels (mppi.py:1194-1351) after the first solve() of
    the ref_solve.npz scenario, for the three planner modes.  Everything the kernels read is stored next to
    their output so that the oracle's restatement can be checked without re-running the solve."""
    g = np.load(os.path.join(OUT, "ref_solve.npz"))
    H, W = g["obstacle"].shape
    res = float(g["res"])
    out = {}
    for mode, flags in (("tdm", dict(use_tdm=True)), ("det", dict(use_det_dynamics=True)),
                        ("spd", dict(use_nom_dynamics_with_speed_map=True))):
        cfg = _quiet(Config, T=float(g["T_s"]), dt=float(g["dt"]), num_grid_samples=int(g["M"]),
                     num_control_rollouts=int(g["N"]), seed=int(g["seed"]),
                     max_map_dim=tuple(int(v) for v in g["max_map_dim"]),
                     tdm_sample_thread_dim=tuple(int(v) for v in g["thread_dim"]),
                     max_speed_padding=float(g["max_speed_padding"]), num_vis_state_rollouts=5, **flags)
        lt, at = _quiet(TDM_Numba, cfg), _quiet(TDM_Numba, cfg)
        d = dict(res=res, xlimits=np.array([0.0, W * res]), ylimits=np.array([0.0, H * res]),
                 bin_values=g["bin_values"], bin_values_bounds=np.array([0.0, 1.0]), det_dynamics_cvar_alpha=0.4)
        _quiet(lt.set_TDM_from_PMF_grid, g["pmf_lin"], d, g["obstacle"], g["unknown"])
        _quiet(at.set_TDM_from_PMF_grid, g["pmf_ang"], d, g["obstacle"], g["unknown"])
        for t_ in (lt, at):
            t_.sample_grid_batch_d.copy_to_device(np.zeros(t_.sample_grid_batch_d.shape, dtype=np.int8))
        pl = _quiet(MPPI_Numba, cfg)
        p = base_params([2.3, 3.1, 0.3], [5.0, 4.5], cvar_alpha=0.5)
        pl.setup(p, lt, at)
        orig = MPPI_Numba.update_useq_numba

        class _One:                                  # the update kernel is racy with 32 threads
            def __getitem__(self, cfg_):
                return orig[1, 1]
        pl.update_useq_numba = _One()
        u1 = _quiet(pl.solve).copy()
        assert np.array_equal(u1, g[mode + "_u1"]), mode            # same run as ref_solve.npz
        states = _quiet(pl.get_state_rollout).copy()
        out[mode + "_states"] = states
        out[mode + "_u_cur"] = pl.u_cur_d.copy_to_host().copy()
        out[mode + "_u_prev"] = pl.u_prev_d.copy_to_host().copy()
        out[mode + "_noise"] = pl.noise_samples_d.copy_to_host().copy()
        out[mode + "_lin_grid"] = lt.sample_grid_batch_d.copy_to_host().copy()
        out[mode + "_ang_grid"] = at.sample_grid_batch_d.copy_to_host().copy()
        out[mode + "_pxl"] = np.asarray(lt.padded_xlimits, dtype=np.float64)
        out[mode + "_pyl"] = np.asarray(lt.padded_ylimits, dtype=np.float64)
        out[mode + "_V"] = int(pl.num_vis_state_rollouts)
        print("state rollouts", mode, states.shape, "done")
    out.update(x0=np.array([2.3, 3.1, 0.3]), res=res, dt=0.1, vrange=[0.0, 3.0], wrange=[-np.pi, np.pi],
               bounds=[0.0, 1.0])
    np.savez_compressed(os.path.join(OUT, "ref_state_rollout.npz"), **out)


def main():
    from oracle.ref_loader import load_reference
    if "--only-state-rollout" in sys.argv:
        Config, TDM_Numba, MPPI_Numba, cuda = load_reference()
        make_state_rollout_golden(Config, TDM_Numba, MPPI_Numba)
        return
    if "--only-oversized" in sys.argv:
        Config, TDM_Numba, MPPI_Numba, cuda = load_reference()
        make_oversized_golden(MPPI_Numba, cuda)
        return
    Config, TDM_Numba, MPPI_Numba, cuda = load_reference()
    os.makedirs(OUT, exist_ok=True)
    f32 = np.float32

    # ---------------------------------------------------------------- 1. RNG + noise (mppi.py:118,1354-1370)
    cfg = _quiet(Config, T=0.8, dt=0.1, num_grid_samples=1, num_control_rollouts=100, seed=1,
                 max_map_dim=(30, 30), use_det_dynamics=True)
    pl = _quiet(MPPI_Numba, cfg)
    st0 = pl.rng_states_d.copy_to_host()
    u_std_d = cuda.to_device(np.array([2.0, 3.0], dtype=f32))
    MPPI_Numba.sample_noise_numba[100, 8](pl.rng_states_d, u_std_d, pl.noise_samples_d)
    n1 = pl.noise_samples_d.copy_to_host().copy()
    MPPI_Numba.sample_noise_numba[100, 8](pl.rng_states_d, u_std_d, pl.noise_samples_d)
    n2 = pl.noise_samples_d.copy_to_host().copy()
    st2 = pl.rng_states_d.copy_to_host()
    np.savez_compressed(os.path.join(OUT, "ref_noise.npz"), seed=1, N=100, T=8, u_std=[2.0, 3.0],
                        states0=np.stack([st0["s0"], st0["s1"]], 1), noise1=n1, noise2=n2,
                        states2=np.stack([st2["s0"], st2["s1"]], 1))
    print("ref_noise done")

    # ---------------------------------------------------------------- 2. PMF setters + grid sampling (terrain.py)
    rng = np.random.default_rng(7)
    B, H, W = 5, 14, 11
    bin_values = np.array([0.0, 0.25, 0.5, 0.75, 1.0])
    pmf_lin = random_pmf(rng, B, H, W)
    pmf_ang = random_pmf(rng, B, H, W)
    obstacle = (rng.random((H, W)) < 0.08).astype(np.int8)
    unknown = (rng.random((H, W)) < 0.08).astype(np.int8)
    res = 0.5
    tdm_dict = dict(res=res, xlimits=np.array([1.0, 1.0 + W * res]), ylimits=np.array([-2.0, -2.0 + H * res]),
                    bin_values=bin_values, bin_values_bounds=np.array([0.0, 1.0]), det_dynamics_cvar_alpha=0.3)
    out = dict(pmf_lin=pmf_lin, pmf_ang=pmf_ang, obstacle=obstacle, unknown=unknown, res=res,
               xlimits=tdm_dict["xlimits"], ylimits=tdm_dict["ylimits"], bin_values=bin_values,
               bounds=[0.0, 1.0], max_speed_padding=5.0, dt=0.1, max_map_dim=[20, 18], seed=1,
               thread_dim=[4, 3], M=3)
    for mode, flags in (("tdm", dict(use_tdm=True)), ("det", dict(use_det_dynamics=True)),
                        ("spd", dict(use_nom_dynamics_with_speed_map=True))):
        for alpha in (0.3, 1.0):
            cfg = _quiet(Config, T=1.0, dt=0.1, num_grid_samples=3, num_control_rollouts=100, seed=1,
                         max_map_dim=(20, 18), tdm_sample_thread_dim=(4, 3), max_speed_padding=5.0, **flags)
            tdm = _quiet(TDM_Numba, cfg)
            d = dict(tdm_dict)
            d["det_dynamics_cvar_alpha"] = alpha
            _quiet(tdm.set_TDM_from_PMF_grid, pmf_lin, d, obstacle, unknown)
            key = "%s_a%02d" % (mode, int(alpha * 10))
            out[key + "_pmf_padded"] = tdm.pmf_grid_d.copy_to_host()
            out[key + "_pxl"] = np.asarray(tdm.padded_xlimits)
            out[key + "_pyl"] = np.asarray(tdm.padded_ylimits)
            out[key + "_pad"] = tdm.pad_cells
            out[key + "_obs_padded"] = tdm.obstacle_map_d.copy_to_host()
            out[key + "_unk_padded"] = tdm.unknown_map_d.copy_to_host()
            if mode == "spd":
                out[key + "_risk"] = tdm.risk_traction_map_d.copy_to_host()
            Hp, Wp = out[key + "_pmf_padded"].shape[1:]
            # zero the (uninitialised) sample buffer so that unwritten cells are comparable
            tdm.sample_grid_batch_d.copy_to_device(np.zeros(tdm.sample_grid_batch_d.shape, dtype=np.int8))
            st = tdm.rng_states_d.copy_to_host()
            out[key + "_states0"] = np.stack([st["s0"], st["s1"]], 1)
            g1 = _quiet(tdm.sample_grids, 1.0).copy_to_host().copy()
            g2 = _quiet(tdm.sample_grids, 0.6).copy_to_host().copy()
            st = tdm.rng_states_d.copy_to_host()
            out[key + "_grid1"] = g1
            out[key + "_grid2"] = g2
            out[key + "_states2"] = np.stack([st["s0"], st["s1"]], 1)
            print("terrain", key, "done", g1.shape)
    np.savez_compressed(os.path.join(OUT, "ref_terrain.npz"), **out)

    # ---------------------------------------------------------------- 2b. semantic-grid setter (terrain.py:183-342)
    make_semantic_golden(Config, TDM_Numba)

    # ---------------------------------------------------------------- 2c. barebone map-free variant (notebook)
    make_barebone_golden(cuda)

    # ---------------------------------------------------------------- 3. rollouts (mppi.py:613-1111)
    rng = np.random.default_rng(11)
    M, R, C = 6, 26, 24
    Hp, Wp = 24, 22
    lin = rng.integers(0, 101, (M, R, C)).astype(np.int8)
    ang = rng.integers(0, 101, (M, R, C)).astype(np.int8)
    lin[:, :Hp, :Wp][:, [0, 1, Hp - 2, Hp - 1], :] = 0
    lin[:, :Hp, :Wp][:, :, [0, 1, Wp - 2, Wp - 1]] = 0
    obs = (rng.random((Hp, Wp)) < 0.05).astype(np.int8)
    unk = (rng.random((Hp, Wp)) < 0.05).astype(np.int8)
    risk = rng.integers(5, 101, (1, Hp, Wp)).astype(np.int8)
    res = f32(0.25)
    xlim = np.array([-1.0, -1.0 + Wp * 0.25], dtype=f32)
    ylim = np.array([2.0, 2.0 + Hp * 0.25], dtype=f32)
    N, T = 24, 12
    noise = (rng.standard_normal((N, T, 2)) * np.array([2.0, 3.0])).astype(f32)
    u_cur = np.stack([rng.uniform(0, 2, T), rng.uniform(-1, 1, T)], 1).astype(f32)
    x0 = np.array([1.7, 4.9, 0.4], dtype=f32)
    xgoal_near = np.array([2.6, 5.6], dtype=f32)       # some rollouts reach it (early break)
    xgoal_far = np.array([30.0, 30.0], dtype=f32)
    common = dict(lin=lin, ang=ang, obs=obs, unk=unk, risk=risk, res=res, xlim=xlim, ylim=ylim,
                  noise=noise, u_cur=u_cur, x0=x0, xgoal_near=xgoal_near, xgoal_far=xgoal_far,
                  lin_bounds=[0.0, 1.0], ang_bounds=[0.0, 1.0], vrange=[0.0, 3.0], wrange=[-np.pi, np.pi],
                  u_std=[2.0, 3.0], v_post=0.01, obs_cost=1e5, unk_cost=1e2, goal_tol=0.5, lam=1.0,
                  dt=0.1, dist_weight=1.0)
    dev = cuda.to_device

    def launch_sto(goal, alpha, grids_l, grids_a, block):
        costs_d = cuda.device_array((N,), dtype=f32)
        MPPI_Numba.rollout_numba[N, block, 0, 4 * block](
            dev(grids_l), dev(grids_a), dev(np.array([0, 1], f32)), dev(np.array([0, 1], f32)), dev(obs), dev(unk),
            res, dev(xlim), dev(ylim), dev(np.array([0, 3], f32)), dev(np.array([-np.pi, np.pi], f32)), dev(goal),
            f32(0.01), f32(1e5), f32(1e2), f32(0.5), f32(1.0), dev(np.array([2, 3], f32)), f32(alpha), dev(x0),
            f32(0.1), 1.0, dev(noise), dev(u_cur), costs_d)
        return costs_d.copy_to_host()

    def launch_det(goal, speed_map):
        costs_d = cuda.device_array((N,), dtype=f32)
        args = [dev(lin[:1]), dev(ang[:1])]
        if speed_map:
            args.append(dev(risk))
        args += [dev(np.array([0, 1], f32)), dev(np.array([0, 1], f32)), dev(obs), dev(unk),
                 res, dev(xlim), dev(ylim), dev(np.array([0, 3], f32)), dev(np.array([-np.pi, np.pi], f32)), dev(goal),
                 f32(0.01), f32(1e5), f32(1e2), f32(0.5), f32(1.0), dev(np.array([2, 3], f32)), dev(x0),
                 f32(0.1), 1.0, dev(noise), dev(u_cur), costs_d]
        k = MPPI_Numba.rollout_det_dyn_w_speed_map_numba if speed_map else MPPI_Numba.rollout_det_dyn_numba
        k[N, 1](*args)
        return costs_d.copy_to_host()

    for gname, goal in (("near", xgoal_near), ("far", xgoal_far)):
        # per-(n,m) costs from the reference itself: one-thread blocks on map m with alpha = 1
        cnm = np.stack([launch_sto(goal, 1.0, lin[m:m + 1], ang[m:m + 1], 1) for m in range(M)], 1)
        common["sto_cnm_" + gname] = cnm
        for alpha in (0.5, 0.9, 1.0):
            common["sto_cvar%02d_%s" % (int(alpha * 10), gname)] = launch_sto(goal, alpha, lin, ang, M)
        common["det_" + gname] = launch_det(goal, False)
        common["spd_" + gname] = launch_det(goal, True)
        print("rollouts", gname, "done")
    np.savez_compressed(os.path.join(OUT, "ref_rollout.npz"), **common)
    make_oversized_golden(MPPI_Numba, cuda)          # the M > 1024 kernel on the same inputs

    # ---------------------------------------------------------------- 4. update (mppi.py:1113-1191), launched [1,1]
    rng = np.random.default_rng(13)
    Nu, Tu = 150, 9
    costs = (rng.uniform(20, 30, Nu)).astype(f32)
    noise_u = (rng.standard_normal((Nu, Tu, 2)) * np.array([2.0, 3.0])).astype(f32)
    u0 = np.stack([rng.uniform(0, 2.9, Tu), rng.uniform(-3, 3, Tu)], 1).astype(f32)
    upd = dict(costs=costs, noise=noise_u, u0=u0, vrange=[0.0, 3.0], wrange=[-np.pi, np.pi])
    for lam in (1.0, 0.3):
        c_d, w_d, u_d = dev(costs.copy()), cuda.device_array((Nu,), dtype=f32), dev(u0.copy())
        MPPI_Numba.update_useq_numba[1, 1](f32(lam), c_d, dev(noise_u), w_d,
                                          dev(np.array([0, 3], f32)), dev(np.array([-np.pi, np.pi], f32)), u_d)
        upd["u_lam%02d" % int(lam * 10)] = u_d.copy_to_host()
        upd["w_lam%02d" % int(lam * 10)] = w_d.copy_to_host()
    np.savez_compressed(os.path.join(OUT, "ref_update.npz"), **upd)
    print("update done")

    # ---------------------------------------------------------------- 5. whole solve() through the public API
    rng = np.random.default_rng(17)
    B, H, W = 5, 12, 12
    pmf_l = random_pmf(rng, B, H, W, zero_frac=0.2)
    pmf_a = random_pmf(rng, B, H, W, zero_frac=0.2)
    obstacle = (rng.random((H, W)) < 0.05).astype(np.int8)
    unknown = (rng.random((H, W)) < 0.05).astype(np.int8)
    res = 0.5
    solve = dict(pmf_lin=pmf_l, pmf_ang=pmf_a, obstacle=obstacle, unknown=unknown, res=res,
                 bin_values=bin_values, max_map_dim=[24, 24], N=100, M=4, T_s=0.6, dt=0.1, seed=1,
                 thread_dim=[4, 4], max_speed_padding=5.0)
    for mode, flags in (("tdm", dict(use_tdm=True)), ("det", dict(use_det_dynamics=True)),
                        ("spd", dict(use_nom_dynamics_with_speed_map=True))):
        cfg = _quiet(Config, T=0.6, dt=0.1, num_grid_samples=4, num_control_rollouts=100, seed=1,
                     max_map_dim=(24, 24), tdm_sample_thread_dim=(4, 4), max_speed_padding=5.0, **flags)
        lt, at = _quiet(TDM_Numba, cfg), _quiet(TDM_Numba, cfg)
        d = dict(res=res, xlimits=np.array([0.0, W * res]), ylimits=np.array([0.0, H * res]),
                 bin_values=bin_values, bin_values_bounds=np.array([0.0, 1.0]), det_dynamics_cvar_alpha=0.4)
        _quiet(lt.set_TDM_from_PMF_grid, pmf_l, d, obstacle, unknown)
        _quiet(at.set_TDM_from_PMF_grid, pmf_a, d, obstacle, unknown)
        for t_ in (lt, at):
            t_.sample_grid_batch_d.copy_to_device(np.zeros(t_.sample_grid_batch_d.shape, dtype=np.int8))
        pl = _quiet(MPPI_Numba, cfg)
        p = base_params([2.3, 3.1, 0.3], [5.0, 4.5], cvar_alpha=0.5)
        pl.setup(p, lt, at)
        # the reference's update kernel is racy in the simulator with 32 threads:
        # run it with one thread by wrapping the launch configuration.
        orig = MPPI_Numba.update_useq_numba

        class _One:
            def __getitem__(self, cfg_):
                return orig[1, 1]
        pl.update_useq_numba = _One()
        u1 = _quiet(pl.solve).copy()
        solve[mode + "_u1"] = u1
        solve[mode + "_noise1"] = pl.noise_samples_d.copy_to_host().copy()
        solve[mode + "_lin_grid1"] = lt.sample_grid_batch_d.copy_to_host().copy()
        solve[mode + "_ang_grid1"] = at.sample_grid_batch_d.copy_to_host().copy()
        pl.shift_and_update(np.array([2.4, 3.15, 0.35]), u1, num_shifts=1)
        u2 = _quiet(pl.solve).copy()
        solve[mode + "_u2"] = u2
        solve[mode + "_weights2"] = pl.weights_d.copy_to_host().copy()
        print("solve", mode, "done")
    np.savez_compressed(os.path.join(OUT, "ref_solve.npz"), **solve)
    make_state_rollout_golden(Config, TDM_Numba, MPPI_Numba)     # get_state_rollout after the first solve


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    main()
