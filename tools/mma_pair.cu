// Tuning aid: tensor-pipe throughput of a streaming K loop (TMA -> smem ring -> tcgen05.mma) with one CTA per tile
// (cta_group::1, 128 x N) against a CTA pair per tile (cta_group::2, 256 x N, each CTA holds half of B).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/mma_pair.cu pitchextractor_b200/csrc/host.cu -o tools/_bin/mma_pair
#include "../pitchextractor_b200/csrc/common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace pe;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t bar_cluster, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

// alloc_by: 0 = warp 2 of both CTAs issues the pair allocation, 1 = only the leader CTA's warp 2
template <int CTAS>
__global__ void __launch_bounds__(128, 1)
stream_k(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int bn, int kblocks, int kb_span,
         int stages, int load, int share_b, long long* out, uint32_t* info) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0;
  const int bn_local = bn / CTAS;
  const int stage_bytes = 16384 + bn_local * 128;
  uint64_t* full = (uint64_t*)(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* done = empty + stages;
  uint32_t* slot = (uint32_t*)(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!load) for (int i = threadIdx.x; i < stages * stage_bytes / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CTAS == 1) tmem_alloc(slot, 256);
    else tmem_alloc_pair(slot, 256);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) info[blockIdx.x] = tm;
  const int row_a = (blockIdx.x) * 128;  // every CTA streams its own A rows; B is shared by all CTAs
  if (warp == 0 && lane == 0 && load) {
    int s = 0; uint32_t ph = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* sa = smem + (size_t)s * stage_bytes;
      uint8_t* sb = sa + 16384;
      const int k0 = (kb % kb_span) * 64;
      if (CTAS == 1) {
        mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
        tma_load_2d(&ta, &full[s], sa, k0, row_a);
        tma_load_2d(&tb, &full[s], sb, k0, share_b ? 0 : (int)blockIdx.x * bn);
      } else {
        if (rank == 0) mbar_arrive_expect_tx(&full[s], 2u * (uint32_t)stage_bytes);
        const uint32_t lead = mapa_shared(smem_u32(&full[s]), 0);
        tma_load_2d_pair(&ta, lead, sa, k0, row_a);
        tma_load_2d_pair(&tb, lead, sb, k0, (share_b ? 0 : (int)(blockIdx.x / 2) * bn) + (int)rank * bn_local);
      }
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  }
  if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc(UMMA_BF16, 128 * CTAS, bn, 0, 0);
    int s = 0; uint32_t ph = 0;
    const long long t0 = clock64();
    for (int kb = 0; kb < kblocks; ++kb) {
      if (load) { mbar_wait(&full[s], ph); tc_fence_after(); }
      const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + 16384;
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = umma_desc_sw128(sa + k * 32, 16, 1024), db = umma_desc_sw128(sb + k * 32, 16, 1024);
        if (CTAS == 1) tc_mma_bf16(tm, da, db, idesc, 1); else tc_mma_bf16_pair(tm, da, db, idesc, 1);
      }
      if (CTAS == 1) tc_commit(&empty[s]); else tc_commit_pair(&empty[s], 3);
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    if (CTAS == 1) tc_commit(done); else tc_commit_pair(done, 3);
    mbar_wait(done, 0);
    out[blockIdx.x] = clock64() - t0;
  } else if (threadIdx.x == 64) {
    mbar_wait_relaxed(done, 0);  // the peer CTA also sees the multicast commit
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    if (CTAS == 1) tmem_dealloc(tm, 256);
    else tmem_dealloc_pair(tm, 256);
  }
}

template <int CTAS>
static void run(const CUtensorMap& ta, const CUtensorMap& tb, int bn, int kblocks, int kb_span, int load, int share_b,
                long long* d, uint32_t* info) {
  const int bn_local = bn / CTAS, stage_bytes = 16384 + bn_local * 128;
  int stages = (198 * 1024) / stage_bytes; if (stages > 8) stages = 8;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  if (getenv("STAGES")) stages = atoi(getenv("STAGES"));
  cudaFuncSetAttribute(stream_k<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaMemset(d, 0, 148 * 8);
  cudaError_t e = cudaLaunchKernelEx(&cfg, stream_k<CTAS>, ta, tb, bn, kblocks, kb_span, stages, load, share_b, d, info);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long h[148]; uint32_t hi[148];
  cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hi, info, 148 * 4, cudaMemcpyDeviceToHost);
  double mx = 0, sum = 0; int n = 0;
  for (int i = 0; i < 148; ++i) if (h[i] > 0) { sum += h[i]; if (h[i] > mx) mx = h[i]; ++n; }
  const double flops_per_kb = 2.0 * 128 * CTAS * bn * 64;
  printf("ctas=%d N=%3d load=%d share_b=%d stages=%d: %7.1f cyc/k-block (max %7.1f)  -> %5.0f TFLOP/s at 1.965 GHz  tmem[0..3]=%x %x %x %x  %s %s\n",
         CTAS, bn, load, share_b, stages, sum / n / kblocks, mx / kblocks,
         flops_per_kb * n / (mx / kblocks) * 1.965e9 / 1e12, hi[0], hi[1], hi[2], hi[3], cudaGetErrorString(e), cudaGetErrorString(e2));
}

int main(int argc, char** argv) {
  const int K = 4096, rowsA = 148 * 128, rowsB = 256 * 148;
  void *A, *B; cudaMalloc(&A, (size_t)rowsA * K * 2); cudaMalloc(&B, (size_t)rowsB * K * 2);
  cudaMemset(A, 0x3c, (size_t)rowsA * K * 2); cudaMemset(B, 0x3c, (size_t)rowsB * K * 2);
  long long* d; cudaMalloc(&d, 148 * 8); uint32_t* info; cudaMalloc(&info, 148 * 4);
    // (only the leader allocating hangs: both CTAs of the pair must issue tcgen05.alloc.cta_group::2)
  const int span = argc > 1 ? atoi(argv[1]) : 8;
  for (int bn : {256, 128}) {
    CUtensorMap ta, tb1, tb2;
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)rowsA}, str[1] = {(uint64_t)K * 2}; uint32_t box[2] = {64, 128};
    if (pe_host::encode_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, str, box)) { printf("tmap failed\n"); return 1; }
    dims[1] = rowsB; box[1] = (uint32_t)bn;
    pe_host::encode_tmap(&tb1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, dims, str, box);
    box[1] = (uint32_t)bn / 2;
    pe_host::encode_tmap(&tb2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, dims, str, box);
    run<1>(ta, tb1, bn, 512, span, 1, 1, d, info);
    run<1>(ta, tb1, bn, 512, span, 1, 0, d, info);
    run<2>(ta, tb2, bn, 512, span, 1, 1, d, info);
    run<2>(ta, tb2, bn, 512, span, 1, 0, d, info);
  }
  return 0;
}
