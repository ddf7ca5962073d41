"""The control-update kernels' real source (online softmax in two levels: CTA partials -> rank partial ->
combine), executed on the host by tests/emu_update.py, against the oracle's restatement of update_useq_numba
(mppi.py:1113-1191) and against the reference's own result (ref_update.npz); N sharded over 'ranks' gives the
one-rank update."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import mppi_ref as MR
from tests.emu_update import build

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VR, WR = np.array([0, 3], np.float32), np.array([-np.pi, np.pi], np.float32)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    return build(str(tmp_path_factory.mktemp("emu_update")))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _partial(emu, costs, noise, lam):
    N, T = noise.shape[:2]
    ctas = emu.emu_update_num_ctas(N)
    w_raw = np.zeros(N, np.float32)
    parts = np.zeros((ctas, 2 * T + 2), np.float32)
    rank = np.zeros(2 * T + 2, np.float32)
    emu.emu_update_partial(_p(costs), _p(noise), _p(w_raw), _p(parts), _p(rank), N, T, np.float32(lam))
    return w_raw, parts, rank


def _finish(emu, gathered, w_raw, parts, u0, N, T, lam):
    u = u0.copy()
    w = np.zeros(N, np.float32)
    g = np.ascontiguousarray(gathered, dtype=np.float32)
    emu.emu_update_finish(_p(g), g.shape[0], _p(w_raw), _p(parts), _p(u), _p(w), N, T, np.float32(lam), _p(VR), _p(WR))
    return u, w


@pytest.mark.parametrize("lam", [1.0, 0.3])
def test_update_kernels_match_reference_golden(emu, lam):
    g = np.load(os.path.join(GOLDEN, "ref_update.npz"))
    costs, noise, u0 = (np.ascontiguousarray(g[k], dtype=np.float32) for k in ("costs", "noise", "u0"))
    N, T = noise.shape[:2]
    w_raw, parts, rank = _partial(emu, costs, noise, lam)
    u, w = _finish(emu, rank[None, :], w_raw, parts, u0, N, T, lam)
    key = "lam%02d" % int(lam * 10)
    np.testing.assert_allclose(u, g["u_" + key], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(w, g["w_" + key], rtol=1e-4, atol=1e-12)


@pytest.mark.parametrize("N,T", [(1000, 50), (300, 300), (37, 3), (9600, 8)])
def test_update_kernels_match_oracle(emu, N, T):
    rng = np.random.default_rng(N + T)
    costs = rng.uniform(900, 930, N).astype(np.float32)
    noise = (rng.standard_normal((N, T, 2)) * [2, 3]).astype(np.float32)
    u0 = np.stack([rng.uniform(0, 2.9, T), rng.uniform(-3, 3, T)], 1).astype(np.float32)
    w_raw, parts, rank = _partial(emu, costs, noise, 1.0)
    u, w = _finish(emu, rank[None, :], w_raw, parts, u0, N, T, 1.0)
    want_u, want_w = MR.update_useq(1.0, costs, noise, VR, WR, u0)
    np.testing.assert_allclose(u, want_u, rtol=1e-5, atol=3e-6)
    np.testing.assert_allclose(w, want_w, rtol=1e-4, atol=1e-12)
    assert abs(float(w.sum()) - 1.0) < 1e-5


def test_sharded_update_equals_single_rank(emu):
    """4 'ranks' own N/4 control sequences each; their (2T+2)-float partials, gathered, give every rank the one-rank u."""
    N, T, ws = 2048, 64, 4
    rng = np.random.default_rng(3)
    costs = rng.uniform(900, 930, N).astype(np.float32)
    noise = (rng.standard_normal((N, T, 2)) * [2, 3]).astype(np.float32)
    u0 = rng.uniform(0, 1, (T, 2)).astype(np.float32)
    w_raw, parts, rank = _partial(emu, costs, noise, 1.0)
    u1, w1 = _finish(emu, rank[None, :], w_raw, parts, u0, N, T, 1.0)
    shards = []
    for r in range(ws):
        sl = slice(N * r // ws, N * (r + 1) // ws)
        shards.append(_partial(emu, np.ascontiguousarray(costs[sl]), np.ascontiguousarray(noise[sl]), 1.0))
    gathered = np.stack([s[2] for s in shards])
    ws_w = []
    for r in range(ws):
        u, w = _finish(emu, gathered, shards[r][0], shards[r][1], u0, N // ws, T, 1.0)
        np.testing.assert_allclose(u, u1, rtol=1e-5, atol=2e-6)
        ws_w.append(w)
    np.testing.assert_allclose(np.concatenate(ws_w), w1, rtol=1e-4, atol=1e-12)


@pytest.mark.parametrize("N,T", [(1000, 50), (8192, 128), (37, 3)])
def test_one_launch_update_equals_partial_plus_finish(emu, N, T):
    """One rank: the last CTA of update_partial_kernel applies the update itself (UPD_TAIL_APPLY) -- bit-identical to
    the two-step path (rank partial, then update_apply_kernel on that single partial), and it follows the oracle."""
    rng = np.random.default_rng(N * 7 + T)
    costs = rng.uniform(900, 930, N).astype(np.float32)
    noise = (rng.standard_normal((N, T, 2)) * [2, 3]).astype(np.float32)
    u0 = np.stack([rng.uniform(0, 2.9, T), rng.uniform(-3, 3, T)], 1).astype(np.float32)
    w_raw, parts, rank = _partial(emu, costs, noise, 1.0)
    u2, w2 = _finish(emu, rank[None, :], w_raw, parts, u0, N, T, 1.0)
    ctas = emu.emu_update_num_ctas(N)
    w_raw1, parts1, rank1 = np.zeros(N, np.float32), np.zeros((ctas, 2 * T + 2), np.float32), np.zeros(2 * T + 2, np.float32)
    u1, w1 = u0.copy(), np.zeros(N, np.float32)
    assert emu.emu_update_one_rank(_p(costs), _p(noise), _p(w_raw1), _p(parts1), _p(rank1), _p(u1), _p(w1), N, T,
                                   np.float32(1.0), _p(VR), _p(WR)) == 0
    assert (u1 == u2).all() and (w1 == w2).all() and (rank1 == rank).all()
    want_u, _ = MR.update_useq(1.0, costs, noise, VR, WR, u0)
    np.testing.assert_allclose(u1, want_u, rtol=1e-5, atol=3e-6)


def test_update_tail_broadcasts_the_rank_partial_to_every_peer(emu):
    """UPD_TAIL_BCAST (peer-memory exchange): 3 'ranks', each stores its partial into slot `rank` of all three
    gather buffers and raises its own epoch flag in every peer -- the gather buffers end up identical and equal to
    the stacked rank partials; combining them gives the one-rank update."""
    N, T, ws = 1536, 32, 3
    rng = np.random.default_rng(11)
    costs = rng.uniform(900, 930, N).astype(np.float32)
    noise = (rng.standard_normal((N, T, 2)) * [2, 3]).astype(np.float32)
    u0 = rng.uniform(0, 1, (T, 2)).astype(np.float32)
    gather = np.zeros((ws, ws, 2 * T + 2), np.float32)
    flags = np.zeros((ws, ws), np.uint32)
    shards = []
    for r in range(ws):
        sl = slice(N * r // ws, N * (r + 1) // ws)
        n = N // ws
        ctas = emu.emu_update_num_ctas(n)
        w_raw, parts, rank = np.zeros(n, np.float32), np.zeros((ctas, 2 * T + 2), np.float32), np.zeros(2 * T + 2, np.float32)
        assert emu.emu_update_bcast(_p(np.ascontiguousarray(costs[sl])), _p(np.ascontiguousarray(noise[sl])), _p(w_raw),
                                    _p(parts), _p(rank), n, T, np.float32(1.0), ws, r, _p(gather), _p(flags), 5) == 0
        shards.append((w_raw, parts, rank))
    assert (flags == 5).all()
    for q in range(ws):
        assert (gather[q] == np.stack([s[2] for s in shards])).all()
    w_raw, parts, rank = _partial(emu, costs, noise, 1.0)
    u1, _ = _finish(emu, rank[None, :], w_raw, parts, u0, N, T, 1.0)
    u, _ = _finish(emu, gather[0], shards[0][0], shards[0][1], u0, N // ws, T, 1.0)
    np.testing.assert_allclose(u, u1, rtol=1e-5, atol=2e-6)
