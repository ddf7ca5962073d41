// Peer-memory (NVLink / NVSwitch) exchange for the sharded solve: the two collectives of a multi-rank
// solve() -- the all-to-all of per-(n,m) rollout costs and the all-gather of the (2T+2)-float softmax
// partials -- done by this library's own kernels with plain stores into the peers' memory plus epoch flags,
// instead of two NCCL calls (about 52 + 17 us of fixed latency at 8 GPUs against a 0.65 ms solve).
//
// Protocol (one "exchange" = one epoch e, a counter that only grows):
//   producer rank r:  data stores into peer d's buffer -> __threadfence_system() -> flags_d[r] = e
//   consumer rank d:  spin until flags_d[s] >= e for every s (bounded by a timeout), then the next kernel in
//                     the stream reads the data (kernel boundary = L1 invalidate)
// Buffers are written by exactly one producer per slot, flags only grow, the gather buffer is double
// buffered by epoch parity (a fast rank may start exchange e+1 before a slow one has consumed e).
#include "common.cuh"
#include "kernels.h"

namespace b200 {

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- all-to-all of cost blocks (only for rollouts that did not store straight into the peers -- the generic
// rollout kernel): block d of the local staged array (ws, Mc, n_red), contiguous n_red*Mc floats, goes to block
// `rank` of peer d's receive buffer (ws, Mc, n_red).  grid = (ctas_per_peer, ws).
__global__ void __launch_bounds__(256) p2p_push_kernel(const P2PPushArgs a) {
  const int d = blockIdx.y;
  const size_t block_elems = (size_t)a.n_red * a.Mc;
  const float* src = a.costs_nm + (size_t)d * block_elems;
  float* dst = a.peer_recv[d] + (size_t)a.rank * block_elems;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((block_elems & 3) == 0) {                 // 16-byte stores: both bases are 256-byte aligned
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (size_t i = gtid; i < block_elems / 4; i += stride) d4[i] = s4[i];
  } else {
    for (size_t i = gtid; i < block_elems; i += stride) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    const unsigned prev = atomicAdd(a.counter, 1u);
    if (prev == total - 1) {                     // last CTA of the grid: everything of this rank is out
      *a.counter = 0;
      __threadfence_system();
      for (int p = 0; p < a.ws; ++p) st_flag_sys(a.peer_flags[p] + a.rank, a.epoch);
    }
  }
}

// ---- all-gather of the rank partial (len = 2T+2 floats) into slot `rank` of every peer's gather buffer
__global__ void __launch_bounds__(256) p2p_bcast_kernel(const P2PBcastArgs a) {
  for (int p = 0; p < a.ws; ++p) {
    float* dst = a.peer_gather[p] + (size_t)a.rank * a.len;
    for (int j = threadIdx.x; j < a.len; j += blockDim.x) dst[j] = a.partial[j];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.ws) {
    __threadfence_system();
    st_flag_sys(a.peer_flags[threadIdx.x] + a.rank, a.epoch);
  }
}

// ---- consumer side: one warp; lane s waits for rank s.  A rank that never arrives must not hang the GPU:
// after timeout_ns the kernel gives up and records 1 + s in *status (checked by the host after the solve).
__global__ void __launch_bounds__(32) p2p_wait_kernel(const uint32_t* flags, int ws, uint32_t epoch,
                                                      unsigned long long timeout_ns, int* status) {
  const int s = threadIdx.x;
  if (s < ws) {
    const uint64_t t0 = globaltimer_ns();
    unsigned spins = 0;
    while ((int32_t)(ld_flag_sys(flags + s) - epoch) < 0) {
      if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) {
        atomicExch(status, 1 + s);
        break;
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}

void launch_p2p_push(const P2PPushArgs& a, cudaStream_t st) {
  const size_t block_elems = (size_t)a.n_red * a.Mc;
  size_t per_thread = (block_elems & 3) == 0 ? 4 : 1;
  size_t ctas = (block_elems / per_thread + 256 * 4 - 1) / (256 * 4);      // ~4 stores per thread
  if (ctas < 1) ctas = 1;
  if (ctas > 64) ctas = 64;
  p2p_push_kernel<<<dim3((unsigned)ctas, (unsigned)a.ws), 256, 0, st>>>(a);
}

void launch_p2p_bcast(const P2PBcastArgs& a, cudaStream_t st) { p2p_bcast_kernel<<<1, 256, 0, st>>>(a); }

void launch_p2p_wait(const uint32_t* flags, int ws, uint32_t epoch, unsigned long long timeout_ns, int* status,
                     cudaStream_t st) {
  p2p_wait_kernel<<<1, 32, 0, st>>>(flags, ws, epoch, timeout_ns, status);
}

}  // namespace b200
