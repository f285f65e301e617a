// Tuning aid: achievable global-store rate of the epilogue warps for different access patterns.
// pattern 0: lane owns a row, 4 x 16-byte stores per 64-byte row segment (what a TMEM-row-per-lane epilogue does)
// pattern 1: 4 lanes per row segment (64 B contiguous per row, 8 rows per instruction)
// pattern 2: 8 lanes per 128-byte line (4 rows per instruction)
// pattern 3: fully linear (512 contiguous bytes per warp instruction)
// pattern 4: lane writes its row into a swizzled 32 x 128 B staging block in shared memory, one lane issues a TMA tensor store
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/store_rate.cu pitchextractor_b200/csrc/host.cu -o tools/_bin/store_rate
#include "../pitchextractor_b200/csrc/common.cuh"
#include <cstdio>
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(pe::smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__global__ void __launch_bounds__(1024, 1) k(const __grid_constant__ CUtensorMap tm, uint4* out, long long rows, int ld_u4 /* row pitch in uint4 */, int pattern, int tiles_per_cta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  extern __shared__ __align__(1024) uint8_t stg_all[];
  uint8_t* stg = stg_all + warp * 4096;
  // tile = 128 rows x 256 bf16 columns (512 B per row = 32 uint4); warp w handles rows (w % 4) * 32.., column chunks by w / 4
  for (int t = 0; t < tiles_per_cta; ++t) {
    const long long tile = (long long)blockIdx.x * tiles_per_cta + t;
    const long long row0 = (tile * 128) % rows;
    const int colgroups = nwarps / 4;
    for (int c0 = (warp / 4) * (pattern == 4 ? 2 : 1); c0 < 8; c0 += colgroups * (pattern == 4 ? 2 : 1)) {
      const int c = c0;      // 8 chunks of 32 columns (64 B = 4 uint4)
      const int q = warp & 3;
      const uint4 v = make_uint4(lane, t, c, warp);
      if (pattern == 4) {
        if (c & 1) continue;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4*>(stg + lane * 128 + ((u ^ (lane & 7)) << 4)) = v;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tm, stg, c * 32, (int)(row0 + q * 32));
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (pattern == 0) {
        uint4* o = out + (row0 + q * 32 + lane) * ld_u4 + c * 4;
        for (int j = 0; j < 4; ++j) o[j] = v;
      } else if (pattern == 1) {
        for (int i = 0; i < 4; ++i) out[(row0 + q * 32 + i * 8 + (lane >> 2)) * ld_u4 + c * 4 + (lane & 3)] = v;
      } else if (pattern == 2) {  // chunk pairs: 128 B per row; handle as c even only, 2x the work
        if (c & 1) continue;
        for (int i = 0; i < 8; ++i) out[(row0 + q * 32 + i * 4 + (lane >> 3)) * ld_u4 + c * 4 + (lane & 7)] = v;
      } else {
        for (int i = 0; i < 4; ++i) out[(row0 + q * 32) * ld_u4 + c * 128 + i * 32 + lane] = v;
      }
    }
  }
  if (pattern == 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
  const long long rows = 12288; const int ld = 1536 * 2 / 16;  // 37.7 MB
  uint4* d; cudaMalloc(&d, rows * ld * 16);
  CUtensorMap tm;
  { uint64_t dims[2] = {1536, (uint64_t)rows}, str[1] = {1536 * 2}; uint32_t box[2] = {64, 32};
    if (pe_host::encode_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box)) { printf("tmap failed\n"); return 1; } }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {256, 512, 1024}) for (int pattern = 0; pattern < 5; ++pattern) {
    const int tiles = 96 * 6 * 8, per = 32;  // 576 tiles of 128x256 -> 144 CTAs x 4
    k<<<tiles / per, threads, 128 * 1024>>>(tm, d, rows, ld, pattern, per);
    cudaEventRecord(e0);
    for (int r = 0; r < 20; ++r) k<<<tiles / per, threads, 128 * 1024>>>(tm, d, rows, ld, pattern, per);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)tiles * 128 * 512;
    printf("threads %4d pattern %d: %6.2f us per pass, %5.2f TB/s, %5.1f B/cycle/SM  %s\n", threads, pattern, ms / 20 * 1e3,
           bytes / (ms / 20 * 1e-3) / 1e12, bytes / 144 / (ms / 20 * 1e-3 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
