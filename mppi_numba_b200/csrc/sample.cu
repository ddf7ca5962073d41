// sample.cu -- traction-map sampling from the PMF grid.
// Reference: TDM_Numba.sample_grids + sample_grids_numba (mppi_numba/terrain.py:610-694) and the
// generator set-up numba.cuda.random.create_xoroshiro128p_states (numba/cuda/random.py:226-264).
//
// Bit-exact contract: generator  tid_x*(ty*M) + m*ty + tid_y  walks its ceil(rows/tx) x ceil(cols/ty)
// tile row-major and draws ONE uniform per cell; q = int8(ceil(f64(u_f32)*100*alpha)); the first bin
// whose cumulative PMF reaches q is written as int8(100*(v_bin-lo)/(hi-lo)) (float64, truncated);
// if no bin reaches q the cell keeps its previous content.  The table of cumulative PMFs is built
// once per set_pmf (cell-major, so one cell's bins are contiguous) instead of re-summing B strided
// int8 loads per cell per map as the reference does.
#include "kernels.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------
// (B, rows, cols) PMF -> (rows, cols, bpad) running sums, clamped to int8 (comparisons against an
// int8 threshold are unaffected by the clamp); bins >= B repeat the last sum.
__global__ void build_cum_kernel(const int8_t* __restrict__ pmf, int8_t* __restrict__ cum, int B, int bpad,
                                 int rows, int cols) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= rows * cols) return;
  int acc = 0;
  for (int b = 0; b < bpad; ++b) {
    if (b < B) acc += pmf[(size_t)b * rows * cols + cell];
    cum[(size_t)cell * bpad + b] = (int8_t)max(-128, min(127, acc));
  }
}

void launch_build_cum(const int8_t* pmf, int8_t* cum, int num_bins, int bpad, int rows, int cols,
                      cudaStream_t st) {
  const int cells = rows * cols;
  build_cum_kernel<<<(cells + 255) / 256, 256, 0, st>>>(pmf, cum, num_bins, bpad, rows, cols);
}

// ---------------------------------------------------------------------------------------------
// v1 sampler: one thread per generator, direct global accesses.
__global__ void __launch_bounds__(128) sample_grids_kernel(const SampleGridsArgs a) {
  // thread order: ty fastest, then tx, then map -- a warp works on neighbouring tiles of one map
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_map = a.tx * a.ty;
  if (tid >= (int64_t)per_map * a.num_maps) return;
  const int m = (int)(tid / per_map);
  const int r = (int)(tid % per_map);
  const int tix = r / a.ty, tiy = r % a.ty;
  const int64_t gen = (int64_t)tix * ((int64_t)a.ty * a.num_maps) + (int64_t)m * a.ty + tiy;

  const int ncol = (a.cols + a.ty - 1) / a.ty;     // ceil(grid_cols / threads_y)
  const int nrow = (a.rows + a.tx - 1) / a.tx;
  const int r0 = min(tix * nrow, a.rows), r1 = min(r0 + nrow, a.rows);
  const int c0 = min(tiy * ncol, a.cols), c1 = min(c0 + ncol, a.cols);

  ulonglong2* sp = reinterpret_cast<ulonglong2*>(a.states) + gen;
  const ulonglong2 raw = *sp;
  Xoro s{raw.x, raw.y};
  int8_t* __restrict__ grid = a.grid + (size_t)m * a.grid_rows * a.pitch;

  for (int ri = r0; ri < r1; ++ri) {
    const int8_t* __restrict__ cum_row = a.cum + ((size_t)ri * a.cols + c0) * a.bpad;
    int8_t* __restrict__ out_row = grid + (size_t)ri * a.pitch;
    for (int ci = c0; ci < c1; ++ci, cum_row += a.bpad) {
      const float u = xoro_unit_f32(xoro_next(s));
      const double thr = ceil(((double)u * 100.0) * a.alpha_dyn);
      const int q = (int)(int8_t)(short)__double2int_rz(thr);     // cvt.rzi.s16.f64 ; low byte
      for (int b = 0; b < a.num_bins; ++b) {
        if (q <= (int)cum_row[b]) { out_row[ci] = a.qvals[b]; break; }
      }
    }
  }
  *sp = make_ulonglong2(s.s0, s.s1);
}

void launch_sample_grids(const SampleGridsArgs& a, cudaStream_t st) {
  const int64_t total = (int64_t)a.tx * a.ty * a.num_maps;
  const int threads = 128;
  sample_grids_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(a);
}

// ---------------------------------------------------------------------------------------------
// Host: numba-compatible generator states.  State 0 = splitmix64(seed) in both words; state i is
// state i-1 jumped 2^64 steps (random.py:47-69,103-126,226-241).
static inline uint64_t rotl_h(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline void next_h(uint64_t& s0, uint64_t& s1) {
  uint64_t t = s1 ^ s0;
  s0 = rotl_h(s0, 55) ^ t ^ (t << 14);
  s1 = rotl_h(t, 36);
}
static inline void jump_h(uint64_t& s0, uint64_t& s1) {
  static const uint64_t JUMP[2] = {0xbeac0467eba5facbULL, 0xd86b048b86aa9922ULL};
  uint64_t a0 = 0, a1 = 0;
  for (int i = 0; i < 2; ++i)
    for (int b = 0; b < 64; ++b) {
      if (JUMP[i] & (1ULL << b)) { a0 ^= s0; a1 ^= s1; }
      next_h(s0, s1);
    }
  s0 = a0; s1 = a1;
}

void create_xoroshiro_states(uint64_t* out, int64_t first, int64_t count, uint64_t seed) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  uint64_t s0 = z, s1 = z;
  for (int64_t i = 0; i < first; ++i) jump_h(s0, s1);
  for (int64_t i = 0; i < count; ++i) {
    out[2 * i] = s0; out[2 * i + 1] = s1;
    jump_h(s0, s1);
  }
}

}  // namespace b200
