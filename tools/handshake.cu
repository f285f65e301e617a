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
// work: bit0 = producer issues 2 TMA loads (16 KB + 8 KB) with expect_tx, bit1 = consumer issues 4 MMAs (N=64),
// bit2 = consumer executes tcgen05.fence::after_thread_sync after the wait
template <int WAIT>
__global__ void __launch_bounds__(320, 1) k(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, const __grid_constant__ CUtensorMap t4,
                                            int iters, int stages, int release, int group, int work, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8], empty[8], never;
  __shared__ uint32_t slot;
  const int extra_warps_poll = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_init(&never, 1); fence_barrier_init(); }
  if (warp == 2) tmem_alloc(&slot, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const int gm = group - 1;
  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      wait_<WAIT>(&empty[s | gm], ph ^ 1u);
      if (elect_one()) {
        if (work & 1) {
          uint8_t* sa = smem + s * 24576;
          mbar_arrive_expect_tx(&full[s], 24576);
          if (work & 8) tma_load_4d(&t4, &full[s], sa, 0, ((i % 5) * 16) - 1 + (i & 1), ((i / 5) % 24) * 8 - 1, (int)blockIdx.x % 64);
          else tma_load_2d(&ta, &full[s], sa, (i & 7) * 64, (int)blockIdx.x * 128);
          tma_load_2d(&tb, &full[s], sa + 16384, (i & 7) * 64, 0);
        } else {
          mbar_arrive(&full[s]);
        }
      }
      __syncwarp();
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    int s = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      wait_<WAIT>(&full[s], ph);
      if (work & 4) tc_fence_after();
      if (elect_one()) {
        if (work & 2) {
          const uint32_t a = smem_u32(smem + s * 24576) >> 4, b = a + 1024;
          const uint64_t d0 = umma_desc_sw128(0, 16, 1024);
          const uint32_t idesc = umma_idesc(UMMA_BF16, 128, 64, 0, 0);
          for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(slot, d0 | (uint64_t)(a + kk * 2), d0 | (uint64_t)(b + kk * 2), idesc, 1);
        }
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
  if (warp == 2) tmem_dealloc(slot, 64);
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  void *A, *B; cudaMalloc(&A, (size_t)148 * 128 * 512 * 2); cudaMalloc(&B, (size_t)64 * 512 * 2);
  cudaMemset(A, 0, (size_t)148 * 128 * 512 * 2); cudaMemset(B, 0, (size_t)64 * 512 * 2);
  CUtensorMap ta, tb;
  { uint64_t dims[2] = {512, 148 * 128}, str[1] = {1024}; uint32_t box[2] = {64, 128};
    pe_host::encode_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, str, box);
    dims[1] = 64; box[1] = 64; pe_host::encode_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, dims, str, box); }
  CUtensorMap t4;
  { void* X; cudaMalloc(&X, (size_t)64 * 192 * 80 * 64 * 2); cudaMemset(X, 0, (size_t)64 * 192 * 80 * 64 * 2);
    uint64_t dims[4] = {64, 80, 192, 64}, str[3] = {128, 80 * 128, 192 * 80 * 128}; uint32_t box[4] = {64, 16, 8, 1};
    pe_host::encode_tmap(&t4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, X, dims, str, box); }
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int wait : {0}) for (int release : {1}) for (int group : {1, 4}) for (int poll : {3, 7, 11, 15}) {
    const int iters = 4096;
    k<0><<<148, 320, 197 * 1024>>>(ta, tb, t4, iters, 8, release, group, poll, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    printf("wait=%s release=%s group=%d work(bit0 TMA, bit1 MMA, bit2 fence, bit3 4-D A box)=%d: %6.1f cycles per k-block  %s\n", wait ? "test_wait" : "try_wait ", release ? "tcgen05.commit" : "mbarrier.arrive",
           group, poll, h[0] / (double)iters, cudaGetErrorString(e));
  }
  return 0;
}
