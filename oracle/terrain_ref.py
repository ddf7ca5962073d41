"""CPU restatement of the map side of the hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows mppi_numba/terrain.py of the reference:
  * get_padding_info / set_padding / set_padding_2d / set_padding_risk_traction  (terrain.py:511-583)
  * the PMF preprocessing inside set_TDM_from_PMF_grid for the three planner modes (terrain.py:380-508)
  * sample_grids_numba (terrain.py:633-694) with numba's xoroshiro128+ stream layout.
Typing follows the COMPILED kernels, not the simulator.
"""
import math
import numpy as np

from . import xoroshiro as X


# ----------------------------------------------------------------------------- padding
def get_padding_info(grid_shape, max_speed_padding, dt, res, max_map_dim):
    """terrain.py:562-583.  Returns (valid_rows, valid_cols, pad_cells)."""
    rows, cols = grid_shape[-2], grid_shape[-1]
    pad_cells = int(np.ceil(max_speed_padding * dt / res))
    max_rows = max_map_dim[0] - 2 * pad_cells
    max_cols = max_map_dim[1] - 2 * pad_cells
    assert max_rows >= 1 and max_cols >= 1
    return min(max_rows, rows), min(max_cols, cols), pad_cells


def set_padding(pmf_grid, max_speed_padding, dt, res, xlimits, ylimits, max_map_dim):
    """terrain.py:525-543: crop from the origin corner, ring of pad_cells with 100% mass on bin 0."""
    vr, vc, pad = get_padding_info(pmf_grid.shape, max_speed_padding, dt, res, max_map_dim)
    pxl = np.array([xlimits[0] - pad * res, xlimits[0] + (vc + pad) * res])
    pyl = np.array([ylimits[0] - pad * res, ylimits[0] + (vr + pad) * res])
    out = np.zeros((pmf_grid.shape[0], vr + 2 * pad, vc + 2 * pad), dtype=np.int8)
    out[0] = np.int8(100)
    out[:, pad:pad + vr, pad:pad + vc] = pmf_grid[:, :vr, :vc]
    return out, pxl, pyl, pad


def set_padding_2d(grid2d, max_speed_padding, dt, res, max_map_dim, pad_val=0):
    """terrain.py:546-559 (masks) and :511-522 (risk map, as a 2-D array here)."""
    vr, vc, pad = get_padding_info(grid2d.shape, max_speed_padding, dt, res, max_map_dim)
    out = (pad_val * np.ones((vr + 2 * pad, vc + 2 * pad))).astype(np.int8)
    out[pad:pad + vr, pad:pad + vc] = grid2d[:vr, :vc]
    return out


# ----------------------------------------------------------------------------- PMF preprocessing
def collapse_pmf_det_dynamics(pmf_grid, bin_values, alpha):
    """terrain.py:408-452 (use_det_dynamics): one-hot PMF at the first bin whose value is >= the
    mean (alpha == 1) or >= CVaR_alpha of the cell's traction distribution."""
    B, H, W = pmf_grid.shape
    bv = np.asarray(bin_values, dtype=np.float32)
    out = np.zeros((B, H, W), dtype=np.int8)
    pmf_cumsum = 0.01 * pmf_grid.cumsum(axis=0).astype(float)
    weighted = 0.01 * pmf_grid.astype(float) * bv.reshape((-1, 1, 1))
    wcum = np.cumsum(weighted, axis=0)
    r = np.repeat(np.arange(H), W)
    c = np.tile(np.arange(W), H)
    if alpha == 1.0:
        target = wcum[-1]
    else:
        upto = np.argmax(pmf_cumsum >= alpha, axis=0).ravel()
        target = (wcum[upto, r, c] / (pmf_cumsum[upto, r, c] + 1e-6)).reshape(H, W)
    layer = np.argmax(target <= bv.reshape((-1, 1, 1)), axis=0).ravel()
    out[layer, r, c] = np.int8(100)
    return out


def risk_traction_map(pmf_grid, bin_values, bounds, alpha):
    """terrain.py:455-495 (use_nom_dynamics_with_speed_map): returns (one-hot-last-bin PMF,
    int8 (1,H,W) worst-case traction map in percent of the traction range, truncated)."""
    B, H, W = pmf_grid.shape
    bv = np.asarray(bin_values, dtype=np.float32)
    bd = np.asarray(bounds, dtype=np.float32)
    onehot = np.zeros((B, H, W), dtype=np.int8)
    onehot[-1] = np.int8(100)
    pmf_cumsum = 0.01 * pmf_grid.cumsum(axis=0).astype(float)
    weighted = 0.01 * pmf_grid.astype(float) * bv.reshape((-1, 1, 1))
    wcum = np.cumsum(weighted, axis=0)
    trange = bd[1] - bd[0]
    if alpha == 1.0:
        risk = np.reshape(100 * (wcum[-1] - bd[0]) / trange, (1, H, W)).astype(np.int8)
    else:
        layer = np.argmax(pmf_cumsum >= alpha, axis=0).ravel()
        r = np.repeat(np.arange(H), W)
        c = np.tile(np.arange(W), H)
        cv = wcum[layer, r, c] / (pmf_cumsum[layer, r, c].ravel() + 1e-6)
        risk = np.reshape(100 * np.asarray((cv.reshape(H, W) - bd[0]) / trange), (1, H, W)).astype(np.int8)
    return onehot, risk


# ----------------------------------------------------------------------------- sampling
def sample_rng_states(cfg_seed, num_grid_samples, thread_dim, det_dyn):
    """terrain.py:170-177: M*tx*ty generators (tx*ty in the deterministic modes), all seeded cfg.seed."""
    tx, ty = thread_dim
    n = tx * ty if det_dyn else num_grid_samples * tx * ty
    return X.create_states(n, cfg_seed)


def quantise_bin_values(bin_values, bounds):
    """terrain.py:689 as COMPILED: int8( 100. * (f32 - f32) / f64(f32 range) ), float64 arithmetic,
    truncation toward zero (0.21f -> 20, not 21)."""
    bv = np.asarray(bin_values, dtype=np.float32)
    bd = np.asarray(bounds, dtype=np.float32)
    trange = np.float32(bd[1] - bd[0])
    d = (bv - bd[0]).astype(np.float32)
    v = 100.0 * d.astype(np.float64) / np.float64(trange)
    return np.trunc(v).astype(np.int64).astype(np.int8)


def sample_thresholds(u32, alpha_dyn):
    """terrain.py:683: int8(ceil(float64(u_f32) * 100.0 * alpha_dyn))."""
    return np.ceil(u32.astype(np.float64) * 100.0 * float(alpha_dyn)).astype(np.int64).astype(np.int8)


def sample_grids(grid_batch, pmf_padded, states, bin_values, bounds, alpha_dyn, thread_dim, num_maps):
    """terrain.py:633-694, all ``num_maps`` blocks at once.

    Thread (tid_x, tid_y) of block m owns generator ``tid_x*(ty*num_maps) + m*ty + tid_y``
    (terrain.py:657-658) and walks its ceil(rows/tx) x ceil(cols/ty) tile row-major, drawing ONE
    uniform per cell whether or not a bin is found (terrain.py:679-682).  ``grid_batch`` (num_maps,
    Rmax, Cmax) int8 and ``states`` are updated in place; cells whose column never reaches the
    threshold keep their previous content.
    """
    tx, ty = thread_dim
    B, rows, cols = pmf_padded.shape
    ncol = math.ceil(cols / ty)
    nrow = math.ceil(rows / tx)
    qvals = quantise_bin_values(bin_values, bounds)
    m_i, x_i, y_i = np.meshgrid(np.arange(num_maps), np.arange(tx), np.arange(ty), indexing="ij")
    m_i, x_i, y_i = m_i.ravel(), x_i.ravel(), y_i.ravel()
    gen = x_i * (ty * num_maps) + m_i * ty + y_i
    r0 = np.minimum(x_i * nrow, rows)
    r1 = np.minimum(r0 + nrow, rows)
    c0 = np.minimum(y_i * ncol, cols)
    c1 = np.minimum(c0 + ncol, cols)
    cum = np.cumsum(pmf_padded.astype(np.int64), axis=0)      # compiled: int64 accumulator (PTX add.s64)
    for dr in range(nrow):
        for dc in range(ncol):
            ri = r0 + dr
            ci = c0 + dc
            act = (ri < r1) & (ci < c1)
            if not act.any():
                continue
            g = gen[act]
            u = X.uniform_float32(X.next_u64(states, g))
            q = sample_thresholds(u, alpha_dyn)
            rr, cc, mm = ri[act], ci[act], m_i[act]
            hit = q[None, :].astype(np.int64) <= cum[:, rr, cc]                 # (B, k)
            found = hit.any(axis=0)
            b = np.argmax(hit, axis=0)
            grid_batch[mm[found], rr[found], cc[found]] = qvals[b[found]]
    return grid_batch
