"""The CVaR kernels' real source (warp radix select for M <= 1024, CTA-wide select above), executed on the host
by tests/emu_cvar.py, against the oracle's sort-based restatement of mppi.py:718-755: ties, negatives, all-equal
rows, M not a multiple of 32, the alpha edge cases, the map-major layout with a row stride."""
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
    assert emu.emu_cvar(_ptr(c), _ptr(out), N, M, N, np.float32(alpha)) == 0
    np.testing.assert_allclose(out, MR.cvar_reduce(c, alpha), rtol=2e-5, atol=2e-4)


@pytest.mark.parametrize("M,N,ld", [(64, 70, 96), (256, 33, 33), (1000, 19, 40), (1536, 3, 8)])
def test_cvar_kernel_row_stride_and_ragged_blocks(emu, M, N, ld):
    """Map-major input with a row stride larger than the row (the receive buffer of a sharded solve is (M, N/ws)
    with ld = N/ws; a one-rank buffer has ld = N) and control-sequence counts that do not fill the last CTA: the
    same values reach the same lanes, so the result does not depend on ld, and follows the oracle."""
    rng = np.random.default_rng(M + N)
    c = rng.normal(50, 20, (N, M)).astype(np.float32)
    out, one = np.zeros(N, dtype=np.float32), np.zeros(N, dtype=np.float32)
    assert emu.emu_cvar(_ptr(c), _ptr(out), N, M, ld, np.float32(0.3)) == 0
    assert emu.emu_cvar(_ptr(c), _ptr(one), N, M, N, np.float32(0.3)) == 0
    assert (out == one).all()
    np.testing.assert_allclose(out, MR.cvar_reduce(c, 0.3), rtol=2e-5, atol=2e-4)
