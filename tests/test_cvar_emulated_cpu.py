"""The CVaR kernels' real source (warp radix select for M <= 1024, CTA-wide select above), executed on the host
by tests/emu_cvar.py, against the oracle's sort-based restatement of mppi.py:718-755: ties, negatives, all-equal
rows, M not a multiple of 32, the alpha edge cases, the chunked layout of the sharded exchange."""
import ctypes as C

import numpy as np
import pytest

from oracle import mppi_ref as MR
from tests.emu_cvar import build


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    return build(str(tmp_path_factory.mktemp("emu_cvar")))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("M,alpha", [(6, 0.5), (33, 0.1), (100, 0.999), (256, 0.5), (1000, 0.25), (1024, 1.0), (7, 0.01),
                                     (1, 0.5), (1025, 0.5), (3000, 0.1), (2048, 1.0)])
def test_cvar_kernel_source_matches_oracle(emu, M, alpha):
    rng = np.random.default_rng(M)
    N = 9
    c = rng.normal(0, 100, (N, M)).astype(np.float32)
    c[:, ::3] = np.round(c[:, ::3])            # many exact ties
    c[5] = 7.0                                  # all equal
    c[6] = -np.abs(c[6])                        # all negative
    out = np.zeros(N, dtype=np.float32)
    assert emu.emu_cvar(_ptr(c), _ptr(out), N, M, 1, np.float32(alpha)) == 0
    np.testing.assert_allclose(out, MR.cvar_reduce(c, alpha), rtol=2e-5, atol=2e-4)


@pytest.mark.parametrize("M,ws", [(64, 4), (1536, 2)])
def test_cvar_kernel_chunked_layout_of_the_sharded_exchange(emu, M, ws):
    """After the all-to-all a rank holds (ws, N/ws, M/ws): chunk g = rank g's maps for this rank's control sequences."""
    rng = np.random.default_rng(ws)
    n_red, Mc = 6, M // ws
    full = rng.normal(50, 20, (n_red, M)).astype(np.float32)
    chunked = np.ascontiguousarray(full.reshape(n_red, ws, Mc).transpose(1, 0, 2))
    out = np.zeros(n_red, dtype=np.float32)
    assert emu.emu_cvar(_ptr(chunked), _ptr(out), n_red, Mc, ws, np.float32(0.3)) == 0
    one = np.zeros(n_red, dtype=np.float32)
    assert emu.emu_cvar(_ptr(full), _ptr(one), n_red, M, 1, np.float32(0.3)) == 0
    assert (out == one).all()                   # same values per lane / slot: bit-identical to the unsharded layout
    np.testing.assert_allclose(out, MR.cvar_reduce(full, 0.3), rtol=2e-5, atol=2e-4)
