// Tuning aid: issue rate of tcgen05.mma (cta_group::1, M=128, SS operands) for several N, no TMA in the loop.
#include "../pitchextractor_b200/csrc/common.cuh"
#include <cstdio>
using namespace pe;
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int same_k, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) * 2 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 256);
  fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(UMMA_BF16, 128, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k = same_k ? 0 : (i & 3);
      const int st = same_k ? 0 : ((i >> 2) & 1);
      tc_mma_bf16(tm, umma_desc_sw128(sa + st * 49152 + k * 32, 16, 1024), umma_desc_sw128(sb + st * 49152 + k * 32, 16, 1024), idesc, 1);
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  for (int grid : {1, 148}) for (int same : {1, 0}) for (int N : {64, 128, 256}) {
    k<<<grid, 128, 110 * 1024>>>(N, 2048, same, d);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    printf("grid %3d same_operand %d N=%3d: %.1f cycles per MMA (128xNx16)  err=%s\n", grid, same, N, h[0] / 2048.0, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
