#!/usr/bin/env python
"""bench.py -- rollouts/s (N*M*T state-steps per solve / time) of the MPPI hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c5|c3|c2|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one MPPI_Numba.solve() (num_opt = 1): sample both traction-distribution maps (M maps
each), sample control noise, N x M x T rollouts with cost accumulation, CVaR over M, softmax update,
D2H of the T x 2 control sequence.  Workload (BASELINE.json configs[4], the one the metric and the
north-star target are quoted on; it fits one GPU): CVaR-cost MPPI, N=8192, M=256, T=128, 1024x1024
PMF grid (12 bins, res 0.1 m) -- at N GPUs the 256 sampled maps are sharded over the ranks ("strong"
scaling): every rank rolls all 8192 control sequences out on its M/N maps, the per-(n,m) costs are
exchanged all-to-all and the 2T+2-float softmax partials all-gathered, by the library's own
peer-memory kernels over NVLink (B200MPPI_EXCHANGE=nccl: by two NCCL collectives).

`value`  : device-timed (CUDA events on the planner's stream), inputs resident in HBM.
`e2e`    : the same metric through the public Python API from HOST buffers -- every step does
           shift_and_update(x0, u) (H2D of the T x 2 warm start + the params POD) and solve()
           (D2H of the T x 2 result), wall-clock, max over ranks.
`roofline`: the dominant kernel's algorithmic bytes / its CUDA-event time vs the measured HBM peak.
`cpu_baseline`: the numpy oracle (oracle/mppi_ref.py) on a bounded N-slice, on this box's host cores.
--impl reference: times that CPU path alone (the reference has no CPU implementation of its own;
its Numba-CUDA kernels cannot travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    #       mode   N     M    T    H     res  B   det_alpha
    "c5": ("tdm", 8192, 256, 128, 1024, 0.1, 12, 1.0),
    "c3": ("tdm", 1024, 64, 64, 512, 0.1, 12, 1.0),
    "c2": ("det", 1024, 1, 64, 256, 0.2, 2, 1.0),
    "c4": ("det", 4096, 1, 128, 512, 0.2, 32, 0.3),
}


def build_scenario(name):
    from tests.scenarios import make_scenario
    mode, N, M, T, H, res, B, da = WORKLOADS[name]
    return make_scenario(mode, N=N, M=M, T=T, H=H, W=H, res=res, B=B, seed=1, det_alpha=da)


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
_CPU_SHARED = {}


def _cpu_chunk(args):
    n0, n1, seed = args
    sc, maps = _CPU_SHARED["sc"], _CPU_SHARED["maps"]       # inherited through fork, not pickled
    from oracle import mppi_ref as MR
    from oracle import xoroshiro as X
    p = sc["params"]
    T = sc["T"]
    rng = np.random.default_rng(seed + n0)
    noise = (rng.standard_normal((n1 - n0, T, 2)) * p["u_std"]).astype(np.float32)
    mode = dict(tdm=MR.MODE_STOCHASTIC, det=MR.MODE_DET_DYN, spd=MR.MODE_SPEED_MAP)[sc["mode"]]
    cnm = MR.rollout_costs(mode, maps["lin"], maps["ang"], [0, 1], [0, 1], maps["obs"], maps["unk"],
                           np.float32(maps["res"]), maps["pxl"], maps["pyl"], p["vrange"], p["wrange"], p["xgoal"],
                           p["v_post_rollout"], 1e5, 1e2, p["goal_tolerance"], p["lambda_weight"], p["u_std"],
                           p["x0"], p["dt"], 1.0, noise, np.zeros((T, 2), np.float32))
    cn = MR.cvar_reduce(cnm, p["cvar_alpha"]) if sc["mode"] == "tdm" else cnm[:, 0]
    return cn, noise


def cpu_baseline_maps(sc, m_cpu):
    """Sampled maps for the CPU baseline: the oracle's PMF sampler is timed separately (below); the rollout
    sample uses m_cpu iid maps drawn with numpy from the same PMF (statistically the same workload)."""
    from oracle import terrain_ref as TR
    cfgd = sc["cfg"]
    d = sc["tdm_dict"]
    pl, pxl, pyl, pad = TR.set_padding(sc["pmf_lin"], cfgd["max_speed_padding"], cfgd["dt"], d["res"],
                                       d["xlimits"], d["ylimits"], cfgd["max_map_dim"])
    pa, _, _, _ = TR.set_padding(sc["pmf_ang"], cfgd["max_speed_padding"], cfgd["dt"], d["res"],
                                 d["xlimits"], d["ylimits"], cfgd["max_map_dim"])
    q = TR.quantise_bin_values(d["bin_values"], [0, 1])
    rng = np.random.default_rng(0)

    def draw(pmf):
        cum = np.cumsum(pmf.astype(np.int64), axis=0)
        out = np.empty((m_cpu,) + pmf.shape[1:], dtype=np.int8)
        for m in range(m_cpu):
            u = rng.integers(1, 101, pmf.shape[1:])
            out[m] = q[np.argmax(cum >= u[None], axis=0)]
        return out
    mmd = cfgd["max_map_dim"]
    obs = TR.set_padding_2d(sc["obstacle"], cfgd["max_speed_padding"], cfgd["dt"], d["res"], mmd)
    unk = TR.set_padding_2d(sc["unknown"], cfgd["max_speed_padding"], cfgd["dt"], d["res"], mmd)
    return dict(lin=draw(pl), ang=draw(pa), obs=obs, unk=unk, res=d["res"], pxl=pxl.astype(np.float32),
                pyl=pyl.astype(np.float32))


def run_cpu_baseline(sc, n_sample, m_sample, reps=1):
    """numpy oracle: rollouts + CVaR + update on an (n_sample x m_sample x T) slice of the workload, N-sharded
    over all host cores.  Returns (state-steps/s, cores, description)."""
    import multiprocessing as mp
    from oracle import mppi_ref as MR
    cores = os.cpu_count() or 1
    sc2 = dict(sc)
    sc2["M"] = m_sample if sc["mode"] == "tdm" else 1
    maps = cpu_baseline_maps(sc, sc2["M"])
    chunks = max(1, min(cores, n_sample // 8))
    bounds = [n_sample * i // chunks for i in range(chunks + 1)]
    jobs = [(bounds[i], bounds[i + 1], 99) for i in range(chunks)]
    _CPU_SHARED["sc"], _CPU_SHARED["maps"] = sc2, maps
    ctx = mp.get_context("fork")
    best = None
    with ctx.Pool(chunks) as pool:
        for _ in range(reps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_chunk, jobs)
            cn = np.concatenate([r[0] for r in res])
            noise = np.concatenate([r[1] for r in res])
            p = sc["params"]
            MR.update_useq(p["lambda_weight"], cn, noise, p["vrange"], p["wrange"], np.zeros((sc["T"], 2), np.float32))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    steps = n_sample * sc2["M"] * sc["T"]
    desc = "numpy oracle, N-slice %d of %d x M-slice %d of %d x T %d (%d state-steps), %d processes" % (
        n_sample, sc["N"], sc2["M"], sc["M"], sc["T"], steps, chunks)
    return steps / best, chunks, desc, best


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- algorithmic bytes
def algorithmic_bytes(sc, cfg, n_local, m_local):
    """Compulsory HBM bytes per solve and per kernel on ONE rank, each operand once."""
    N, M, T = n_local, m_local, sc["T"]
    B = sc["pmf_lin"].shape[0]
    Hp, Wp = cfg.max_map_dim
    p = sc["params"]
    reach = int(np.ceil(p["vrange"][1] * T * p["dt"] / sc["tdm_dict"]["res"]))
    Hw = min(Hp, 2 * reach + 3)
    sample = 2 * (B * Hp * Wp + M * Hp * Wp)                  # PMF read + sampled maps written, both TDMs
    rollout = 2 * M * Hw * Hw + 2 * Hw * Hw + 8 * N * T + 4 * N * M   # map windows + masks + noise + costs
    noise = 32 * N * T + 8 * N * T                             # RNG state R+W, noise W
    cvar = 4 * N * M + 4 * N
    update = 4 * N + 8 * N * T + 16 * T
    return dict(sample_grids=sample, rollout=rollout, noise=noise, cvar=cvar, update=update,
                total=sample + rollout + noise + cvar + update)


# ----------------------------------------------------------------------------- main arms
def run_reference(args, sc):
    """--impl reference: the CPU restatement of the path (oracle port), all host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_s, m_s = (4096, 64) if sc["mode"] == "tdm" else (min(sc["N"], 4096), 1)
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, cores, desc, dt = run_cpu_baseline(sc, n_s, m_s)
        if i >= args.warmup:
            vals.append(v)
            times.append(dt)
    v = float(np.mean(vals))
    out = {"impl": "reference", "metric": "rollouts/sec (N*M*T state-steps/s)", "value": v,
           "unit": "state-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(args, sc)},
           "cpu_baseline": {"value": v, "unit": "state-steps/s", "cores": cores, "kind": "port", "sample": desc},
           "e2e": {"value": v, "unit": "state-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    _emit(json.dumps(out))


def workload_name(args, sc):
    mode, N, M, T, H, res, B, da = WORKLOADS[args.workload]
    return "%s: %s MPPI N=%d M=%d T=%d, %dx%d PMF grid (%d bins, res %.1f m), num_opt=1" % (
        args.workload, {"tdm": "CVaR-cost", "det": "CVaR-dynamics"}[mode], N, M, T, H, H, B, res)


def run_b200(args, sc):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # CPU baseline FIRST: it forks worker processes, which must happen before CUDA is initialised
        n_s, m_s = (4096, 64) if sc["mode"] == "tdm" else (min(sc["N"], 4096), 1)
        v, cores, desc, _ = run_cpu_baseline(sc, n_s, m_s)
        cpu_base = {"value": v, "unit": "state-steps/s", "cores": cores, "kind": "port", "sample": desc}
    import torch
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run" % (args.gpus, world))
    import __graft_entry__
    __graft_entry__.build()
    import mppi_numba_b200 as E
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        cfg = E.Config(**sc["cfg"])
        lin = E.TDM_Numba(cfg, device=local, rank=rank, world_size=world)
        ang = E.TDM_Numba(cfg, device=local, rank=rank, world_size=world)
        lin.set_TDM_from_PMF_grid(sc["pmf_lin"], sc["tdm_dict"], sc["obstacle"], sc["unknown"])
        ang.set_TDM_from_PMF_grid(sc["pmf_ang"], sc["tdm_dict"], sc["obstacle"], sc["unknown"])
        pl = E.MPPI_Numba(cfg, device=local, rank=rank, world_size=world, process_group=pg)
        pl.setup(sc["params"], lin, ang)
    N, M, T = sc["N"], (sc["M"] if sc["mode"] == "tdm" else 1), sc["T"]
    units = N * M * T

    # all work on one torch stream so that torch.cuda.Event brackets exactly the engine's kernels
    import ctypes as C
    from mppi_numba_b200._lib import lib, check
    stream = torch.cuda.Stream(device=dev)
    if world == 1:
        check(lib.b200mppi_planner_set_stream(pl._handle, C.c_void_p(stream.cuda_stream)))
        for t in (lin, ang):
            check(lib.b200mppi_tdm_set_stream(t._handle, C.c_void_p(stream.cuda_stream)))
    else:
        pl._ensure_exchange()
        stream = pl._stream

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-timed region: K solves, inputs resident
    clocks = ClockSampler(local)          # started before the warm-up: nvidia-smi needs ~0.3 s to produce a sample
    t_w = time.perf_counter()
    for _ in range(args.warmup):
        pl.solve()
    per_solve = max((time.perf_counter() - t_w) / args.warmup, 1e-5) if args.warmup > 0 else 2e-3
    # keep the GPU under the benchmark load for ~0.5 s while the clock sampler starts.  Every solve() of a
    # multi-rank run contains exchanges, so the NUMBER of extra solves must be the same on every rank:
    # agree on it (max over ranks) instead of looping on each rank's own wall clock.
    n_settle = int(max_over_ranks(float(min(5000, int(0.5 / per_solve) + 1))))
    for _ in range(n_settle):
        pl.solve()
    barrier()
    clocks.lines.clear()
    l0 = pl.launch_count()          # includes the TDM kernels launched inside solve()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            u = pl.solve()
        e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = pl.launch_count() - l0      # kernels launched in the timed region
    # very short timed region: extend the load for the clock sampler only (same count on every rank)
    if max_over_ranks(1.0 if len(clocks.lines) < 3 else 0.0) > 0.0:
        for _ in range(int(min(5000, int(0.4 / (ms * 1e-3)) + 1))):
            pl.solve()
    clk = clocks.stop()
    value = units / (ms * 1e-3)

    # ---- end to end through the public API from host buffers (wall clock, H2D + D2H inside)
    x0 = sc["params"]["x0"].copy()
    for _ in range(2):
        pl.shift_and_update(x0, u, 1)
        u = pl.solve()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pl.shift_and_update(x0, u, 1)          # H2D: T x 2 warm start (+ params POD by value)
        u = pl.solve()                         # D2H: T x 2 result, synchronous
    torch.cuda.synchronize(dev)
    wall = max_over_ranks(time.perf_counter() - t0) / args.steps
    barrier()
    e2e = units / wall

    # ---- per-kernel times (CUDA events inside the library) for the roofline of the dominant kernel
    pl.set_profiling(True)
    acc = {}
    reps = max(3, min(10, args.steps))
    for _ in range(reps):
        pl.solve()
        for k, v in pl.last_timings().items():
            acc.setdefault(k, []).append(v)
    pl.set_profiling(False)
    stage_ms = {k: float(np.mean(v)) for k, v in acc.items()}
    ab = algorithmic_bytes(sc, cfg, pl.n_local, pl.m_local)
    peak, peak_src = measured_peaks()
    dom = max(("sample_grids", "rollout", "noise", "cvar", "update"), key=lambda k: stage_ms.get(k, 0.0))
    dom_ms = stage_ms[dom]            # sample_grids: ONE fused launch samples the linear and the angular maps
    dom_bytes = ab[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None                    # dram__bytes_read+write of that kernel from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "r01_dram_traffic.json")) as f:
            tj = json.load(f)
        if args.workload == tj.get("workload") and world == 1:
            traffic = tj["kernels"].get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
                "solve_algorithmic_bytes": ab["total"],
                "solve_frac_of_hbm_roofline": (ab["total"] / (ms * 1e-3) / 1e9) / peak,
                "stage_ms": stage_ms}

    out = None
    if rank == 0:
        out = {"metric": "rollouts/sec (N*M*T state-steps/s)", "value": value, "unit": "state-steps/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": workload_name(args, sc), "global_rollouts": N, "maps": M, "horizon": T,
                          "parallelism": "single GPU" if world == 1 else ("maps sharded x%d (M/G maps per rank, all N rollouts), all-to-all of N*M/G costs + all-gather of %d floats per solve" if sc["mode"] == "tdm" else "N-sharded x%d, 1 all-gather of %d floats per solve") % (world, 2 * T + 2),
                          "exchange": ("none (1 rank)" if world == 1 else
                                       "peer-memory kernels over NVLink (csrc/p2p.cu)" if getattr(pl, "_p2p", False)
                                       else "NCCL all_to_all_single + all_gather"),
                          "l2": "per-step working set (2 x %d MB sampled maps) exceeds the 126 MB L2; no explicit flush"
                                % (M * cfg.max_map_dim[0] * cfg.max_map_dim[1] // 2 ** 20)},
               "clocks": clk,
               "e2e": {"value": e2e, "unit": "state-steps/s", "ms_per_step": wall * 1e3,
                       "h2d_bytes_per_step": 8 * T + 88, "d2h_bytes_per_step": 8 * T},
               "gpu_launches": int(launches),
               "roofline": roofline}
    if cpu_base is not None:
        out["cpu_baseline"] = cpu_base
    if rank == 0:
        _emit(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


_emit = print


def main():
    # The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, the engine's allocation
    # notices) also write to fd 1, so everything is routed to stderr and only the final line goes to the
    # real stdout.
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w", buffering=1)
    global _emit

    def _emit(line):
        os.write(real_stdout, (line + "\n").encode())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    sc = build_scenario(args.workload)
    if args.impl == "reference":
        run_reference(args, sc)
    else:
        run_b200(args, sc)


if __name__ == "__main__":
    main()
