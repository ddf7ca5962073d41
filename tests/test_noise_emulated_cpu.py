"""sample_noise_kernel's real source on the host against the REFERENCE's sample_noise_numba (ref_noise.npz, two
consecutive calls): generator states advance bit for bit, the Box-Muller noise agrees to float rounding."""
import ctypes as C
import os

import numpy as np

from tests.emu_noise import build

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_noise_kernel_source_matches_reference(tmp_path):
    emu = build(str(tmp_path))
    g = np.load(os.path.join(GOLDEN, "ref_noise.npz"))
    N, T = int(g["N"]), int(g["T"])
    states = np.ascontiguousarray(g["states0"]).copy().view(np.uint64).reshape(-1, 2)
    assert states.shape == (N * T, 2)
    for key in ("noise1", "noise2"):
        noise = np.zeros((N, T, 2), np.float32)
        emu.emu_sample_noise(states.ctypes.data_as(C.c_void_p), noise.ctypes.data_as(C.c_void_p), N * T,
                             np.float32(g["u_std"][0]), np.float32(g["u_std"][1]), None)
        np.testing.assert_allclose(noise, g[key], rtol=3e-6, atol=2e-6)
    assert (states == np.ascontiguousarray(g["states2"]).view(np.uint64).reshape(-1, 2)).all()
