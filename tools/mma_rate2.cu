// Tuning aid: effect of tcgen05.commit cadence and of waiting on completion between MMA groups.
#include "../pitchextractor_b200/csrc/common.cuh"
#include <cstdio>
using namespace pe;
// mode 0: commit every 4 MMAs to rotating barriers, never wait (except at the end)
// mode 1: commit every 4 MMAs and wait for that commit before issuing the next group (drain latency)
// mode 2: like 0 but with tcgen05.fence::after_thread_sync + a try_wait on an already completed barrier per group
__global__ void __launch_bounds__(128, 1) k(int N, int groups, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[8]; __shared__ uint64_t done; __shared__ uint64_t ready; __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) * 2 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_init(&done, 1); mbar_init(&ready, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 256);
  fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    mbar_arrive(&ready);  // phase 0 complete
    const uint32_t idesc = umma_idesc(UMMA_BF16, 128, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    uint32_t ph[8] = {0,0,0,0,0,0,0,0};
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (mode == 2) { mbar_wait(&ready, 0); tc_fence_after(); }
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(tm, umma_desc_sw128(sa + (g & 1) * 49152 + k * 32, 16, 1024), umma_desc_sw128(sb + (g & 1) * 49152 + k * 32, 16, 1024), idesc, 1);
      tc_commit(&bars[g & 7]);
      if (mode == 1) { mbar_wait(&bars[g & 7], ph[g & 7]); ph[g & 7] ^= 1; }
    }
    tc_commit(&done);
    mbar_wait(&done, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  for (int mode : {0, 1, 2}) for (int N : {64, 256}) {
    k<<<148, 128, 110 * 1024>>>(N, 512, mode, d);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    printf("mode %d N=%3d: %.1f cycles per group of 4 MMAs  err=%s\n", mode, N, h[0] / 512.0, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
