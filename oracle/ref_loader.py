"""Import the UNMODIFIED reference (mit-acl/mppi_numba) under Numba's CUDA simulator so that its own
kernels can be run in the GPU-less build container (TEST INFRASTRUCTURE, see oracle/__init__.py).

Used only by oracle/make_golden.py (fixture generation in the build container; /root/reference does not
exist on the GPU box).
Nothing here copies reference source: the package is imported from where it lies.

Shims: NUMBA_ENABLE_CUDASIM=1; ``np.float`` (mppi.py:32-33 uses the removed alias);
``cuda.get_current_device`` (config.py:9-12 queries the GPU at import; the simulator lacks it);
``cuda.jit(max_registers=...)`` (mppi.py:761; the simulator's jit rejects the keyword).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MPPI_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mppi_numba"))


def load_reference():
    """Returns (Config, TDM_Numba, MPPI_Numba, cuda).  Must be called before numba is imported
    anywhere else in the process (the simulator switch is read at import time)."""
    if "numba" in sys.modules and os.environ.get("NUMBA_ENABLE_CUDASIM") != "1":
        raise RuntimeError("numba already imported without NUMBA_ENABLE_CUDASIM=1")
    os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float
    from numba import cuda
    cuda.get_current_device = lambda: types.SimpleNamespace(
        MAX_THREADS_PER_BLOCK=1024, MAX_BLOCK_DIM_X=1024, MAX_GRID_DIM_X=2 ** 31 - 1)
    if not getattr(cuda.jit, "_b200_shim", False):
        _jit = cuda.jit

        def jit(*a, **k):
            k.pop("max_registers", None)
            return _jit(*a, **k)
        jit._b200_shim = True
        cuda.jit = jit
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from mppi_numba.config import Config
    from mppi_numba.terrain import TDM_Numba
    from mppi_numba.mppi import MPPI_Numba
    return Config, TDM_Numba, MPPI_Numba, cuda
