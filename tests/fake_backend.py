"""A stand-in for libb200mppi.so that lets the HOST-side logic of mppi_numba_b200 (the Python mirrors of the
reference's Config / TDM_Numba / MPPI_Numba: validation, cropping, padding, PMF construction from semantic grids,
precondition checks, the params POD) run in a process without a GPU.  TEST INFRASTRUCTURE: every C-ABI call
returns success and is recorded with its scalar arguments and the bytes behind its pointer arguments; nothing
is computed.  The product never imports this module."""
import ctypes as C

import numpy as np


def _deref(arg):
    """byref(x) -> x"""
    return getattr(arg, "_obj", arg)


class FakeLib(object):
    def __init__(self):
        self.calls = []              # (name, args)
        self.uploads = {}            # name -> last recorded payload dict
        self._next_handle = 0x1000
        self.issued = set()          # fake handle values handed out

    # -- helpers
    @staticmethod
    def _bytes(p, n, dtype, shape):
        addr = p.value if isinstance(p, C.c_void_p) else C.cast(p, C.c_void_p).value
        return np.frombuffer(C.string_at(addr, int(n)), dtype=dtype).reshape(shape).copy()

    def _handle_out(self, out):
        self._next_handle += 0x10
        _deref(out).value = self._next_handle
        self.issued.add(self._next_handle)
        return 0

    # -- entry points with outputs or payloads worth keeping
    def b200mppi_tdm_create(self, pod, out):
        self.calls.append(("tdm_create", {f: getattr(_deref(pod), f) for f, _ in _deref(pod)._fields_}))
        return self._handle_out(out)

    def b200mppi_planner_create(self, pod, out):
        self.calls.append(("planner_create", {f: getattr(_deref(pod), f) for f, _ in _deref(pod)._fields_}))
        return self._handle_out(out)

    def b200mppi_tdm_sample_grid_view(self, h, base, pitch):
        _deref(base).value = 0xdead0000
        _deref(pitch).value = 16
        return 0

    def b200mppi_tdm_num_generators(self, h, out):
        _deref(out).value = 8
        return 0

    def b200mppi_tdm_set_pmf(self, h, pmf, B, rows, cols, bv, bb, res, pxl, pyl):
        self.uploads["set_pmf"] = dict(
            pmf=self._bytes(pmf, B * rows * cols, np.int8, (B, rows, cols)),
            bin_values=self._bytes(bv, 4 * B, np.float32, (B,)), bounds=self._bytes(bb, 8, np.float32, (2,)),
            res=float(res), pxl=self._bytes(pxl, 8, np.float32, (2,)), pyl=self._bytes(pyl, 8, np.float32, (2,)))
        self.calls.append(("tdm_set_pmf", (B, rows, cols)))
        return 0

    def b200mppi_tdm_set_masks(self, h, obs, unk, rows, cols):
        self.uploads["set_masks"] = dict(obs=self._bytes(obs, rows * cols, np.int8, (rows, cols)),
                                         unk=self._bytes(unk, rows * cols, np.int8, (rows, cols)))
        self.calls.append(("tdm_set_masks", (rows, cols)))
        return 0

    def b200mppi_tdm_set_risk_map(self, h, risk, rows, cols):
        self.uploads["set_risk_map"] = dict(risk=self._bytes(risk, rows * cols, np.int8, (rows, cols)))
        self.calls.append(("tdm_set_risk_map", (rows, cols)))
        return 0

    def b200mppi_tdm_set_bin_quantisation(self, h, q, n):
        self.uploads["set_bin_quantisation"] = dict(q=self._bytes(q, n, np.int8, (n,)))
        self.calls.append(("tdm_set_bin_quantisation", (n,)))
        return 0

    def b200mppi_planner_set_params(self, h, pod):
        p = _deref(pod)
        self.uploads["set_params"] = {f: (list(getattr(p, f)) if hasattr(getattr(p, f), "__len__") else getattr(p, f))
                                      for f, _ in p._fields_}
        self.calls.append(("planner_set_params", ()))
        return 0

    def b200mppi_tdm_set_pmf_collapsed(self, h, raw, B, H, W, keep_r, keep_c, pad, bv, bb, res, pxl, pyl, alpha,
                                       pmf_out, risk_out, bad_out):
        """The device-side collapse + crop + padding of the one-map modes, stood in for by the ORACLE's restatement
        (oracle/terrain_ref.py) so that the Python wrapper around it can be exercised without a GPU."""
        from oracle import terrain_ref as TR
        pmf = self._bytes(raw, B * H * W, np.int8, (B, H, W))
        bin_values = self._bytes(bv, 4 * B, np.float32, (B,))
        bounds = self._bytes(bb, 8, np.float32, (2,))
        speed = risk_out is not None
        if speed:
            onehot, risk = TR.risk_traction_map(pmf, bin_values, bounds, alpha)
        else:
            onehot = TR.collapse_pmf_det_dynamics(pmf, bin_values, alpha)
        Hp, Wp = keep_r + 2 * pad, keep_c + 2 * pad
        out = np.zeros((B, Hp, Wp), dtype=np.int8)
        out[0] = 100
        out[:, pad:pad + keep_r, pad:pad + keep_c] = onehot[:, :keep_r, :keep_c]
        C.memmove(C.cast(pmf_out, C.c_void_p).value, out.ctypes.data, out.nbytes)
        if speed:
            rp = np.zeros((Hp, Wp), dtype=np.int8)
            rp[pad:pad + keep_r, pad:pad + keep_c] = risk[0, :keep_r, :keep_c]
            C.memmove(C.cast(risk_out, C.c_void_p).value, rp.ctypes.data, rp.nbytes)
        _deref(bad_out).value = int((pmf.astype(np.int64).sum(0) != 100).sum())
        self.uploads["set_pmf_collapsed"] = dict(keep=(keep_r, keep_c), pad=pad, alpha=float(alpha), res=float(res),
                                                 pxl=self._bytes(pxl, 8, np.float32, (2,)),
                                                 pyl=self._bytes(pyl, 8, np.float32, (2,)))
        self.calls.append(("tdm_set_pmf_collapsed", (B, H, W)))
        return 0

    def b200mppi_planner_set_obstacles(self, h, xy, rad, count):
        self.uploads["set_obstacles"] = dict(
            xy=self._bytes(xy, 8 * count, np.float32, (count, 2)) if count else np.zeros((0, 2), np.float32),
            rad=self._bytes(rad, 4 * count, np.float32, (count,)) if count else np.zeros((0,), np.float32))
        self.calls.append(("planner_set_obstacles", (count,)))
        return 0

    def b200mppi_planner_set_u(self, h, u):
        self.calls.append(("planner_set_u", ()))
        self.uploads["set_u"] = u
        return 0

    def b200mppi_last_error(self):
        return b"fake backend"

    # -- everything else: succeed and remember the name
    def __getattr__(self, name):
        if not name.startswith("b200mppi_"):
            raise AttributeError(name)

        def ok(*args):
            self.calls.append((name[len("b200mppi_"):], ()))
            return 0
        return ok

    def names(self):
        return [c[0] for c in self.calls]


def install(monkeypatch):
    """Swap the ctypes library object in the host modules for a FakeLib; returns it."""
    import mppi_numba_b200.barebone as B
    import mppi_numba_b200.mppi as M
    import mppi_numba_b200.terrain as T
    fake = FakeLib()
    for mod in (M, T, B):
        monkeypatch.setattr(mod, "lib", fake)
    return fake


def disarm(fake):
    """Objects created against the fake hold fake handles; once the real library is back their __del__ would pass
    those to the real destroy functions.  Clear them (also when a failed test's traceback keeps them alive)."""
    import gc
    from mppi_numba_b200.barebone import MPPI_Numba as Barebone
    from mppi_numba_b200.mppi import MPPI_Numba
    from mppi_numba_b200.terrain import TDM_Numba
    for obj in gc.get_objects():
        if type(obj) in (MPPI_Numba, TDM_Numba, Barebone):         # (isinstance would poke lazy module proxies)
            h = getattr(obj, "_handle", None)
            if h is not None and getattr(h, "value", None) in fake.issued:
                obj._handle = None
