// Tuning aid: cost of the MMA issue loop (smem-resident operands, no loads): cycles per MMA for tile widths N and
// different numbers of MMAs between tcgen05.commit's.
#include "../pitchextractor_b200/csrc/common.cuh"
#include <cstdio>
using namespace pe;
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__global__ void __launch_bounds__(128, 1) k(int N, int groups, int mpc, int uniform, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8]; __shared__ uint64_t done; __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) * 2 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_init(&done, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (uniform && threadIdx.x < 32) {  // the whole warp walks the loop; one elected lane issues
    const uint32_t idesc = umma_idesc(UMMA_BF16, 128, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const uint64_t da0 = umma_desc_sw128(0, 16, 1024), db0 = umma_desc_sw128(0, 16, 1024);
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      const uint32_t a = (sa + (g & 1) * 49152) >> 4, b = (sb + (g & 1) * 49152) >> 4;
      if (elect_one()) {
        for (int k = 0; k < mpc; ++k)
          tc_mma_bf16(tm, da0 | (uint64_t)(a + (k & 3) * 2), db0 | (uint64_t)(b + (k & 3) * 2), idesc, 1);
        tc_commit(&bars[g & 7]);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  } else if (!uniform && threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(UMMA_BF16, 128, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const uint64_t da0 = umma_desc_sw128(0, 16, 1024), db0 = umma_desc_sw128(0, 16, 1024);
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      const uint32_t a = (sa + (g & 1) * 49152) >> 4, b = (sb + (g & 1) * 49152) >> 4;
      for (int k = 0; k < mpc; ++k) {
        tc_mma_bf16(tm, da0 | (uint64_t)(a + (k & 3) * 2), db0 | (uint64_t)(b + (k & 3) * 2), idesc, 1);
      }
      tc_commit(&bars[g & 7]);
    }
    tc_commit(&done);
    mbar_wait(&done, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  for (int N : {64, 128, 256}) for (int mpc : {4, 16}) for (int nacc : {0, 1}) {
    const int groups = 2048 / mpc;
    k<<<148, 128, 100 * 1024>>>(N, groups, mpc, nacc, d);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    printf("N=%3d  MMAs/commit=%2d uniform=%d: %6.1f cycles per MMA (floor 128*N/256 = %3d)  %s\n", N, mpc, nacc, h[0] / 2048.0, N / 2, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
