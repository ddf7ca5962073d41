"""N > 1 path on CPU: two gloo ranks shard the control sequences, each forms its softmax partial
(beta_r, S_r, V_r[2T]) exactly as the engine's update_partial/update_rank kernels define it, the
partials are all-gathered with torch.distributed and merged by the library's host combine
(b200mppi_combine_partials_host, the same math as update_apply_kernel).  Every rank must end with the
u_seq the oracle computes in a single process."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N, T, WS = 300, 11, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs():
    rng = np.random.default_rng(21)
    costs = rng.uniform(700, 712, N).astype(np.float32)
    noise = (rng.standard_normal((N, T, 2)) * [2, 3]).astype(np.float32)
    u0 = np.stack([rng.uniform(0, 2, T), rng.uniform(-1, 1, T)], 1).astype(np.float32)
    return costs, noise, u0


def _rank_main(rank, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WS)
    import __graft_entry__
    __graft_entry__.build()
    from mppi_numba_b200 import _lib
    costs, noise, u0 = _inputs()
    lam = np.float32(1.0)
    n0, n1 = N * rank // WS, N * (rank + 1) // WS            # the library's sharding rule (b200mppi.h)
    c, e = costs[n0:n1], noise[n0:n1]
    beta = c.min()
    w = np.exp((-1.0 / float(lam)) * (c - beta).astype(np.float64)).astype(np.float32)
    V = np.einsum("n,ntk->tk", w.astype(np.float64), e.astype(np.float64)).astype(np.float32)
    part = torch.from_numpy(np.concatenate([[beta, w.sum(dtype=np.float64)], V.ravel()]).astype(np.float32))
    gathered = torch.empty((WS * (2 * T + 2),), dtype=torch.float32)          # flat: gloo wants 1-D
    dist.all_gather_into_tensor(gathered, part)
    g = np.ascontiguousarray(gathered.numpy().reshape(WS, 2 * T + 2))
    out = np.empty((T, 2), dtype=np.float32)
    vr, wr = np.array([0, 3], np.float32), np.array([-np.pi, np.pi], np.float32)
    _lib.check(_lib.lib.b200mppi_combine_partials_host(_lib.ptr(g), WS, T, lam, _lib.ptr(u0), _lib.ptr(vr),
                                                       _lib.ptr(wr), _lib.ptr(out)))
    np.save(os.path.join(out_dir, "u_rank%d.npy" % rank), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_matches_single_process_oracle(tmp_path):
    from oracle import mppi_ref as MR
    port = _free_port()
    mp.spawn(_rank_main, args=(port, str(tmp_path)), nprocs=WS, join=True)
    costs, noise, u0 = _inputs()
    want, _ = MR.update_useq(1.0, costs, noise, [0, 3], [-np.pi, np.pi], u0)
    us = [np.load(os.path.join(str(tmp_path), "u_rank%d.npy" % r)) for r in range(WS)]
    assert (us[0] == us[1]).all()                      # every rank holds the identical update
    np.testing.assert_allclose(us[0], want, rtol=1e-5, atol=2e-6)


def test_generator_shards_are_slices_of_the_global_stream():
    """Rank r's noise generators are the global generators n*T+t of its rollouts (b200mppi.h:
    results do not depend on world_size)."""
    from oracle import xoroshiro as X
    full = X.create_states(N * T, 3)
    for ws in (2, 3):
        for r in range(ws):
            n0, n1 = N * r // ws, N * (r + 1) // ws
            z = X.splitmix64(3)
            s = (z, z)
            for _ in range(n0 * T):
                s = X.jump_scalar(*s)
            assert (int(full[n0 * T, 0]), int(full[n0 * T, 1])) == s
            assert n1 > n0
