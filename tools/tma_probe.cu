// Standalone probe: which tiled-TMA configurations does sm_100a accept?  (debug aid, not product code)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void kern(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int bytes, unsigned* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(s32(smem)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(s32(smem)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(&bar)) : "memory");
  unsigned s = 0;
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) s += smem[i];
  atomicAdd(out, s);
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  int cols = 1034, rows = 1034, maps = 4, pitch = 1040, rank = 3, bw = 128, bh = 64, esz = 1;
  CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  int c0 = 100, c1 = 7;
  switch (variant) {
    case 0: break;                                   // u8 rank3 128x64
    case 1: rank = 2; break;                         // u8 rank2
    case 2: l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE; break;
    case 3: cols = 1040; break;                      // dim0 multiple of 16
    case 4: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; esz = 4; cols = 256; pitch = 1024; bw = 32; rank = 2; break;
    case 5: c0 = 96; c1 = 8; break;                  // aligned coordinates
    case 6: bw = 240; bh = 232; break;
    case 7: rank = 2; bw = 240; bh = 232; c0 = -20; c1 = -9; break;
  }
  unsigned char* d; cudaMalloc(&d, (size_t)maps * rows * pitch); cudaMemset(d, 1, (size_t)maps * rows * pitch);
  unsigned* out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)maps};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(&tm, dt, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int bytes = bw * bh * esz;
  cudaError_t e;
  if (rank == 3) {
    cudaFuncSetAttribute(kern<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232000);
    kern<3><<<1, 256, bytes>>>(tm, c0, c1, 1, bytes, out);
  } else {
    cudaFuncSetAttribute(kern<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232000);
    kern<2><<<1, 256, bytes>>>(tm, c0, c1, 0, bytes, out);
  }
  e = cudaDeviceSynchronize();
  unsigned h = 0; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
  printf("variant %d: encode=%d run=%s sum=%u\n", variant, (int)r, cudaGetErrorString(e), h);
  return 0;
}
