"""The setter kernels' real source on the host: collapse_pad_kernel against the padded one-hot PMFs and risk maps the
REFERENCE's host preprocessing produced (ref_terrain.npz, det / speed-map modes, alpha 0.3 and 1.0) -- bit for bit --
and build_cum_kernel against the cumulative table the emulated sampler test assumes."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.emu_sampler import cumulative_table
from tests.emu_setter import build

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    return build(str(tmp_path_factory.mktemp("emu_setter")))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("mode", ["det", "spd"])
@pytest.mark.parametrize("alpha", [0.3, 1.0])
def test_collapse_pad_kernel_matches_reference_preprocessing(emu, mode, alpha):
    g = np.load(os.path.join(GOLDEN, "ref_terrain.npz"))
    key = "%s_a%02d" % (mode, int(alpha * 10))
    raw = np.ascontiguousarray(g["pmf_lin"], dtype=np.int8)
    B, H, W = raw.shape
    want = g[key + "_pmf_padded"]
    pad = int(g[key + "_pad"])
    Hp, Wp = want.shape[1:]
    keep_r, keep_c = Hp - 2 * pad, Wp - 2 * pad
    bv = np.ascontiguousarray(g["bin_values"], dtype=np.float32)
    bounds = np.asarray(g["bounds"], dtype=np.float32)
    out = np.full((B, Hp, Wp), -1, np.int8)
    rpitch = (Wp + 15) // 16 * 16
    risk = np.zeros((Hp, rpitch), np.int8) if mode == "spd" else None
    bad = np.zeros(1, np.int32)
    emu.emu_collapse_pad(_p(raw), _p(out), _p(risk), _p(bad), _p(bv), B, H, W, keep_r, keep_c, pad, rpitch, float(alpha),
                         np.float32(bounds[0]), np.float32(bounds[1] - bounds[0]), 1 if mode == "det" else 2)
    assert (out == want).all()
    assert bad[0] == int((raw.astype(np.int64).sum(0)[:keep_r, :keep_c] != 100).sum())
    if mode == "spd":
        assert (risk[:, :Wp] == g[key + "_risk"][0]).all()


def test_build_cum_kernel(emu):
    rng = np.random.default_rng(1)
    for B in (5, 12, 32):
        pmf = rng.integers(0, 30, (B, 9, 11)).astype(np.int8)
        pmf[0, 0, 0] = -3                                          # ill-formed entries are clamped, not wrapped
        bpad = (B + 3) // 4 * 4
        cum = np.zeros((9, 11, bpad), np.int8)
        emu.emu_build_cum(_p(pmf), _p(cum), B, bpad, 9, 11)
        assert (cum == cumulative_table(pmf, bpad)).all()
