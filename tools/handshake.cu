// Tuning aid: cost of the producer <-> consumer mbarrier ring handshake (no data, no MMAs).
// consumer release: 0 = mbarrier.arrive, 1 = tcgen05.commit;  wait flavour: 0 = try_wait spin, 1 = test_wait spin
#include "../pitchextractor_b200/csrc/common.cuh"
#include <cstdio>
using namespace pe;
__device__ __forceinline__ uint32_t mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
template <int WAIT> __device__ __forceinline__ void wait_(uint64_t* bar, uint32_t parity) {
  if (WAIT == 0) { while (!mbar_try_wait(bar, parity)) {} } else { while (!mbar_test_wait(bar, parity)) {} }
}
template <int WAIT>
__global__ void __launch_bounds__(320, 1) k(int iters, int stages, int release, int group, int extra_warps_poll, long long* out) {
  __shared__ uint64_t full[8], empty[8], never;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_init(&never, 1); fence_barrier_init(); }
  if (warp == 2) tmem_alloc(&slot, 32);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const int gm = group - 1;
  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      wait_<WAIT>(&empty[s | gm], ph ^ 1u);
      if (elect_one()) mbar_arrive(&full[s]);
      __syncwarp();
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    int s = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      wait_<WAIT>(&full[s], ph);
      if (elect_one()) {
        if ((s & gm) == gm) { if (release) tc_commit(&empty[s]); else mbar_arrive(&empty[s]); }
      }
      __syncwarp();
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  } else if (extra_warps_poll) {
    // epilogue-like warps polling a barrier that completes only at the end
    long long t = clock64();
    while (clock64() - t < (long long)iters * 100) { if (mbar_try_wait(&never, 0)) break; __nanosleep(64); }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 2) tmem_dealloc(slot, 32);
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  for (int wait : {0, 1}) for (int release : {0, 1}) for (int group : {1, 4}) for (int poll : {0, 1}) {
    const int iters = 4096;
    if (wait == 0) k<0><<<148, 320>>>(iters, 8, release, group, poll, d); else k<1><<<148, 320>>>(iters, 8, release, group, poll, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    printf("wait=%s release=%s group=%d other-warps-polling=%d: %6.1f cycles per k-block  %s\n", wait ? "test_wait" : "try_wait ", release ? "tcgen05.commit" : "mbarrier.arrive",
           group, poll, h[0] / (double)iters, cudaGetErrorString(e));
  }
  return 0;
}
