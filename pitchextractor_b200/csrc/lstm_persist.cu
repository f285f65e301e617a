// Persistent BiLSTM recurrence (reference model.py:218-228 -> torch.nn.LSTM, gate order i,f,g,o): ONE launch runs all T
// time steps of the four independent recurrences of a layer (2 sequence models x 2 directions).
//
// Work split: CTA (unit tile, batch tile of NB, recurrence); a unit tile is 64 hidden units (backward; forward when the
// batch needs more than three batch tiles) or 32 (forward otherwise).  The CTA's slice of W_hh -- forward: the 4 x 64 (32)
// gate rows of its units (192 / 96 KB bf16); backward: the 64 unit columns of all 1536 gate rows (192 KB) -- is loaded into shared
// memory ONCE and stays there for every time step; only h_{t-1} (forward) / dgates_{t+1} (backward), 12 - 96 KB per step,
// streams through a small TMA ring.  The MMA is "transposed": weights are the A operand (gate rows / units on the 128 TMEM
// lanes), the batch is the N dimension, so that a batch of 16 keeps all 128 lanes of the epilogue busy.
//   forward  D[(gate, unit), b] = W_hh[(gate, unit), :] . h_{t-1}[b, :]          (K = 384)
//            lanes of a TMEM quarter = 4 gates x 8 units; each lane adds its input projection + biases, applies ITS gate's
//            activation (tanh(x) = 2 sigmoid(2x) - 1: one formula, per-lane constants), stores the activated gate (kept for
//            the backward pass); a 4 x 4 register transpose over the four gate lanes of a unit (two rounds of warp shuffles)
//            then gives every lane all four gates of one batch column: c_t (kept in registers across the steps), h_t.
//   backward, batch tiles of 16 / 32 columns: gate-stacked formulation on a transposed weight copy, see lstm_seq_bwd_gs_kernel
//   backward D[unit, b] = sum_j W_hh[j, unit] dgates_{t+1}[b, j]                  (K = 1536; A = W_hh^T, MN-major)
//            M = 64 units are presented twice (the two 64-row atoms of the A descriptor alias each other), so lanes 64-127
//            hold a copy and the second half of the epilogue warps takes the second half of the batch columns; dc is
//            carried in registers.
// The six (twelve) unit-tile CTAs of a (recurrence, batch tile) exchange h_t / dgates_t through global memory (L2) and a per-step
// arrival counter: writers store, fence, barrier, one release-add; the loader thread of every CTA acquire-polls the
// counter before it issues the TMA loads of the next step.  All CTAs must be co-resident: the launch is cooperative and
// the grid is at most one CTA per SM.
#include "common.cuh"
#include <cstdlib>
#include "../../include/pitchextractor_b200.h"

namespace pe {

constexpr int PH = 384;          // hidden size
constexpr int PG = 4 * PH;       // gate rows per direction
constexpr int P_W_BYTES = 6 * 32768;               // resident weight slice (both directions of use: 192 KB)
constexpr int P_RING_BYTES = 32768;
// backward: the M = 128 MMA sees the CTA's 64 units twice (A descriptor atom stride 0), so that TMEM lanes 64..127 hold a
// copy for the second half of the epilogue warps.  0: the upper lanes read the next k-block instead and are ignored.
#ifndef P_BWD_DUP
#define P_BWD_DUP 1
#endif

struct LstmSeqParams {
  int B, T, nbt, b_first;      // batch rows b_first .. b_first + nbt * NB - 1 are handled by this launch
  float* gx[2];                // per model: [T][B][2*PG] fp32 (time-major): input projection in, activated gates out
  float* c[2];                 // [T][B][2*PH] fp32 cell state
  __nv_bfloat16* y[2];         // [T][B][2*PH] bf16 hidden state
  const float* b_ih[4];        // per recurrence r = model * 2 + dir
  const float* b_hh[4];
  const __nv_bfloat16* dy[2];  // backward: [T][B][2*PH] gradient w.r.t. y from above
  __nv_bfloat16* dg[2];        // backward: [T][B][2*PG] pre-activation gate gradients
  int* flags;                  // [4][nbt][T] arrival counters (zeroed before the launch)
};

struct LstmSeqMaps {
  CUtensorMap act[2];  // per model: forward y (dims 2*PH, B, T); backward dg (dims 2*PG, B, T); box {64, NB, 1}
  CUtensorMap w[4];    // per recurrence W_hh: forward 3-D view (k, unit, gate) box {64, 8, 4}; backward 2-D box {64, 64}
};

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// bounded acquire-poll of an arrival counter (a protocol bug must end in a trap, not in a hung GPU)
__device__ __forceinline__ void wait_counter(const int* p, int target) {
  uint32_t spins = 0;
  while (ld_acquire_gpu(p) < target) {
    __nanosleep(32);
    if (++spins > (1u << 24)) mbar_timeout_trap();
  }
}

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.f, __fdividef(1.f, 1.f + __expf(-2.f * x)), -1.f); }

__host__ __device__ constexpr uint32_t tmem_pow2(uint32_t cols) {  // tcgen05.alloc takes powers of two >= 32
  uint32_t c = 32;
  while (c < cols) c <<= 1;
  return c;
}

template <int NB>
struct PlCfg {
  static constexpr int STAGE_B = NB * 128;                                  // one [NB x 64] bf16 k-block
  static constexpr int STAGES = (P_RING_BYTES / STAGE_B) > 12 ? 12 : (P_RING_BYTES / STAGE_B);
  static constexpr int BWD_EPI_WARPS = NB >= 64 ? 16 : 8;                   // backward epilogue warps
  static constexpr int BWD_THREADS = 64 + 32 * BWD_EPI_WARPS;
  static constexpr size_t SMEM = (size_t)P_W_BYTES + (size_t)STAGES * STAGE_B + (2 * STAGES + 4) * 8 + 16 + 1024;
};

// =================================================================================================================
// forward
// =================================================================================================================
// UT = hidden units per CTA: 64 (two M tiles of 4 gates x 32 units) or 32 (one M tile).  At small batch tiles a time
// step is bound by the issue interval of its MMAs (>= 85 cycles each, whatever N is: 48 per step at UT = 64), so small
// batches use twice as many CTAs with half the MMAs each.
template <int NB, int UT>
__global__ void __launch_bounds__(64 + 32 * (UT / 4), 1)
lstm_seq_fwd_kernel(const __grid_constant__ LstmSeqMaps maps, const LstmSeqParams p) {
  using C = PlCfg<NB>;
  constexpr int KB = PH / 64;  // 6 k-blocks per step
  constexpr int MT = UT / 32;                 // M tiles (128 gate rows each)
  constexpr int EPI_WARPS = 8 * MT;           // M tiles x 4 lane quarters x 2 column halves
  constexpr int W_KB_BYTES = MT * 16384;      // weight bytes per k-block
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_w = smem;                        // [KB][2 M-tiles][128 rows x 128 B]
  uint8_t* s_ring = smem + P_W_BYTES;         // [STAGES][NB rows x 128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_ring + C::STAGES * C::STAGE_B);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* w_bar = empty_bar + C::STAGES;
  uint64_t* done_bar = w_bar + 1;
  uint64_t* free_bar = done_bar + 1;          // accumulator drained by the epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(free_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UT, bt = blockIdx.y, rec = blockIdx.z;
  const int b0 = p.b_first + bt * NB;
  const int model = rec >> 1, dir = rec & 1;
  int* flags = p.flags + ((size_t)rec * p.nbt + bt) * p.T;
  constexpr uint32_t TCOLS = tmem_pow2(MT * NB);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(w_bar, 1);
    mbar_init(done_bar, 1);
    mbar_init(free_bar, EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ loader: weights once, then h_{t-1} per step
    if (elect_one()) {
      mbar_arrive_expect_tx(w_bar, KB * W_KB_BYTES);
      for (int kb = 0; kb < KB; ++kb)
        for (int mq = 0; mq < UT / 8; ++mq)
          tma_load_3d(&maps.w[rec], w_bar, s_w + kb * W_KB_BYTES + mq * 4096, kb * 64, u0 + mq * 8, 0);
    }
    __syncwarp();
    if (elect_one()) {
      uint32_t it = 0;
      for (int step = 1; step < p.T; ++step) {
        const int t_prev = dir ? p.T - step : step - 1;
        wait_counter(flags + (step - 1), PH / UT);   // h_{t-1} of all unit tiles is in global memory
        fence_proxy_async_global();
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % C::STAGES;
          mbar_wait(&empty_bar[s], ((it / C::STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], C::STAGE_B);
          tma_load_3d(&maps.act[model], &full_bar[s], s_ring + s * C::STAGE_B, dir * PH + kb * 64, b0, t_prev);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t IDESC = umma_idesc(UMMA_BF16, 128, NB, 0, 0);
    if (elect_one()) mbar_wait(w_bar, 0);
    __syncwarp();
    uint32_t it = 0;
    for (int step = 1; step < p.T; ++step) {
      if (elect_one()) {
        mbar_wait(free_bar, (uint32_t)(step - 1) & 1u);   // the epilogue of step - 1 has drained the accumulator
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t s = (it + kb) % C::STAGES;
          mbar_wait(&full_bar[s], ((it + kb) / C::STAGES) & 1u);
          tc_fence_after();
          const uint32_t sb = smem_u32(s_ring + s * C::STAGE_B), sa = smem_u32(s_w + kb * W_KB_BYTES);
#pragma unroll
          for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(tm + m * NB, umma_desc_sw128(sa + m * 16384 + k * 32, 16, 1024),
                          umma_desc_sw128(sb + k * 32, 16, 1024), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&empty_bar[s]);
        }
        tc_commit(done_bar);
      }
      __syncwarp();
      it += KB;
    }
  } else {
    // ------------------------------------------------------------ epilogue: 16 warps = 2 M-tiles x 4 lane quarters x
    // 2 column halves (the epilogue is the longest part of a step at wide batch tiles: twice the warps hide its latencies)
    const int ew = warp - 2;
    const int q = warp & 3, m = MT == 2 ? (ew >> 2) & 1 : 0, ch = MT == 2 ? ew >> 3 : ew >> 2;
    const int gl = lane >> 3, ul = lane & 7;             // this lane's gate and unit inside the quarter
    const int u = u0 + (m * 4 + q) * 8 + ul;
    constexpr int NGRP = NB / 8;                         // groups of 4 batch columns handled by this warp
    const int colw = ch * (NB / 2);                      // its first column
    const uint32_t trow = tm + ((uint32_t)(q * 32) << 16) + m * NB + colw;
    const float bias = __ldg(p.b_ih[rec] + gl * PH + u) + __ldg(p.b_hh[rec] + gl * PH + u);
    const float a_s = gl == 2 ? 2.f : 1.f, a_o = gl == 2 ? -1.f : 0.f;   // act(x) = a_s * sigmoid(a_s * x) + a_o
    const bool hi1 = (gl & 2) != 0, hi0 = (gl & 1) != 0;
    float cst[NGRP];
#pragma unroll
    for (int j = 0; j < NGRP; ++j) cst[j] = 0.f;
    // Time-major tensors ([T][B][...]): a step touches one contiguous slab of B rows (DRAM pages stay open), a batch
    // column is one row (12 KB / 3 KB / 1.5 KB) further: per-step base pointer + small constant byte offsets
    const int nvalid = min(NB, p.B - b0) - colw;         // valid columns of this warp's range (may be <= 0)
    constexpr uint32_t sg = 2u * PG * 4u, sc = 2u * PH * 4u, sy = sc / 2u;
    const size_t tg = (size_t)p.B * sg, tc = (size_t)p.B * sc, ty = (size_t)p.B * sy;   // bytes per time step
    char* gx0 = reinterpret_cast<char*>(p.gx[model] + (long long)(b0 + colw) * (2 * PG) + dir * PG + gl * PH + u);
    char* c0p = reinterpret_cast<char*>(p.c[model] + (long long)(b0 + colw + gl) * (2 * PH) + dir * PH + u);
    char* y0p = reinterpret_cast<char*>(p.y[model] + (long long)(b0 + colw + gl) * (2 * PH) + dir * PH + u);
    // Input projections stream through a register FIFO PF groups (of 4 batch columns) deep, filled across step
    // boundaries: the loads of a step's first groups are in flight while the previous step finishes, so neither the DRAM
    // latency of the strided [b][t] rows nor the MMA wait is exposed per group.
    constexpr int PF = NGRP < 4 ? NGRP : 4;
    float zq[PF][4];
    auto load_group = [&](const char* base, int j, float* z) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        z[e] = (4 * j + e) < nvalid ? *reinterpret_cast<const float*>(base + (uint32_t)(4 * j + e) * sg) : 0.f;
    };
    {
      const int t_first = dir ? p.T - 1 : 0;
#pragma unroll
      for (int j = 0; j < PF; ++j) load_group(gx0 + (size_t)t_first * tg, j, zq[j]);
    }
    for (int step = 0; step < p.T; ++step) {
      const int t = dir ? p.T - 1 - step : step;
      const int t_nxt = dir ? t - 1 : t + 1;
      const bool more = step + 1 < p.T;
      char* gxt = gx0 + (size_t)t * tg;
      const char* gxn = gx0 + (size_t)(more ? t_nxt : t) * tg;
      char* ct = c0p + (size_t)t * tc;
      char* yt = y0p + (size_t)t * ty;
      if (step > 0) {
        mbar_wait(done_bar, (uint32_t)(step - 1) & 1u);
        tc_fence_after();
      }
#pragma unroll
      for (int j = 0; j < NGRP; ++j) {
        float zin[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) zin[e] = zq[j % PF][e];
        // refill this FIFO slot: a later group of this step, or an early group of the next step
        if (j + PF < NGRP) load_group(gxt, j + PF, zq[j % PF]);
        else if (more) load_group(gxn, j + PF - NGRP, zq[j % PF]);
        uint32_t acc[4] = {0u, 0u, 0u, 0u};
        if (step > 0) {
          tmem_ld4(trow + 4 * j, acc);
          tmem_ld_wait();
        }
        float a[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float z = zin[e] + bias + __uint_as_float(acc[e]);
          a[e] = fmaf(a_s, sigmoid_fast(a_s * z), a_o);
          if ((4 * j + e) < nvalid)   // activated gate, kept for the backward pass
            *reinterpret_cast<float*>(gxt + (uint32_t)(4 * j + e) * sg) = a[e];
        }
        // 4 x 4 transpose over the lanes (gate 0..3, same unit): afterwards this lane holds i, f, g, o of column 4j + gl
        {
          const float s0 = hi1 ? a[0] : a[2], s1 = hi1 ? a[1] : a[3];
          const float r0 = __shfl_xor_sync(0xffffffffu, s0, 16), r1 = __shfl_xor_sync(0xffffffffu, s1, 16);
          // rows (gates) {gl & 1, (gl & 1) + 2} x columns {2 * (gl >> 1), + 1}
          const float b00 = hi1 ? r0 : a[0], b01 = hi1 ? r1 : a[1];   // gate (gl & 1)      columns 2*hi1, 2*hi1 + 1
          const float b10 = hi1 ? a[2] : r0, b11 = hi1 ? a[3] : r1;   // gate (gl & 1) + 2
          const float t0 = hi0 ? b00 : b01, t1 = hi0 ? b10 : b11;     // what the partner (gate ^ 1) needs from here
          const float x0 = __shfl_xor_sync(0xffffffffu, t0, 8), x1 = __shfl_xor_sync(0xffffffffu, t1, 8);
          // column 2 * hi1 + hi0 == gl: gates 0..3
          const float g0 = hi0 ? x0 : b00, g1 = hi0 ? b01 : x0, g2 = hi0 ? x1 : b10, g3 = hi0 ? b11 : x1;
          const float cn = fmaf(g1, cst[j], g0 * g2);
          cst[j] = cn;
          const float hn = g3 * tanh_fast(cn);
          if ((4 * j + gl) < nvalid) {
            *reinterpret_cast<float*>(ct + (uint32_t)(4 * j) * sc) = cn;
            *reinterpret_cast<__nv_bfloat16*>(yt + (uint32_t)(4 * j) * sy) = __float2bfloat16(hn);
          }
        }
      }
      // publish h_t: every writer makes its stores visible to the async proxy, then one release-add per CTA
      fence_proxy_async_global();
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      if (threadIdx.x == 64) {
        __threadfence();
        red_release_gpu_add(flags + step, 1);
      }
      if (lane == 0) mbar_arrive(free_bar);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tm, TCOLS);
}

// =================================================================================================================
// backward
// =================================================================================================================
template <int NB>
__global__ void __launch_bounds__(PlCfg<NB>::BWD_THREADS, 1)
lstm_seq_bwd_kernel(const __grid_constant__ LstmSeqMaps maps, const LstmSeqParams p) {
  using C = PlCfg<NB>;
  constexpr int KB = PG / 64;  // 24 k-blocks (gate rows) per step
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_w = smem;                        // [KB][64 gate rows x 128 B (64 units)]
  uint8_t* s_ring = smem + P_W_BYTES;         // [STAGES][NB rows x 128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_ring + C::STAGES * C::STAGE_B);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* w_bar = empty_bar + C::STAGES;
  uint64_t* done_bar = w_bar + 1;
  uint64_t* free_bar = done_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(free_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * 64, bt = blockIdx.y, rec = blockIdx.z;
  const int b0 = p.b_first + bt * NB;
  const int model = rec >> 1, dir = rec & 1;
  int* flags = p.flags + ((size_t)rec * p.nbt + bt) * p.T;
  constexpr uint32_t TCOLS = tmem_pow2(NB);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(w_bar, 1);
    mbar_init(done_bar, 1);
    mbar_init(free_bar, C::BWD_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(w_bar, P_W_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(&maps.w[rec], w_bar, s_w + kb * 8192, u0, kb * 64);
    }
    __syncwarp();
    if (elect_one()) {
      uint32_t it = 0;
      for (int step = 1; step < p.T; ++step) {
        // the backward pass walks each direction's own time order in reverse; t_next was processed one step earlier
        const int t = dir ? step : p.T - 1 - step;
        const int t_next = dir ? t - 1 : t + 1;
        wait_counter(flags + (step - 1), 6);
        fence_proxy_async_global();
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % C::STAGES;
          mbar_wait(&empty_bar[s], ((it / C::STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], C::STAGE_B);
          tma_load_3d(&maps.act[model], &full_bar[s], s_ring + s * C::STAGE_B, dir * PG + kb * 64, b0, t_next);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // A = W_hh^T slice, MN-major: 64 units per K row; both 64-row atoms of the M = 128 operand alias the same rows
    constexpr uint32_t IDESC = umma_idesc(UMMA_BF16, 128, NB, 1, 0);
    constexpr uint32_t A_LBO = P_BWD_DUP ? 0u : 8192u;
    if (elect_one()) mbar_wait(w_bar, 0);
    __syncwarp();
    uint32_t it = 0;
    for (int step = 1; step < p.T; ++step) {
      if (elect_one()) {
        mbar_wait(free_bar, (uint32_t)(step - 1) & 1u);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t s = (it + kb) % C::STAGES;
          mbar_wait(&full_bar[s], ((it + kb) / C::STAGES) & 1u);
          tc_fence_after();
          const uint32_t sb = smem_u32(s_ring + s * C::STAGE_B), sa = smem_u32(s_w + kb * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tm, umma_desc_sw128(sa + k * 2048, A_LBO, 1024), umma_desc_sw128(sb + k * 32, 16, 1024), IDESC,
                        (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&empty_bar[s]);
        }
        tc_commit(done_bar);
      }
      __syncwarp();
      it += KB;
    }
  } else {
    // epilogue: lane quarter q -> units (q & 1) * 32 + lane (quarters 2, 3 hold the copy); the (copy, warp group)
    // combinations split the NB batch columns: 8 or 16 warps (wide batch tiles: more warps hide the epilogue's latencies)
    constexpr int PARTS = C::BWD_EPI_WARPS / 4;           // warps per lane quarter
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int u = u0 + (q & 1) * 32 + lane;
    constexpr int NC = P_BWD_DUP ? NB / (2 * PARTS) : NB / PARTS;       // columns per thread
    const int col0 = P_BWD_DUP ? ((q >> 1) * PARTS + part) * NC : part * NC;
    const bool lanes_valid = P_BWD_DUP || q < 2;
    const uint32_t trow = tm + ((uint32_t)(q * 32) << 16) + col0;
    float dcs[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) dcs[j] = 0.f;
    // time-major tensors: per-step base pointers + constant byte offsets per batch column
    const int nvalid = lanes_valid ? min(NB, p.B - b0) - col0 : 0;   // valid columns of this thread's range
    constexpr uint32_t sg = 2u * PG * 4u, sc = 2u * PH * 4u, sy = 2u * PH * 2u, sd = 2u * PG * 2u;
    const size_t tg = (size_t)p.B * sg, tcb = (size_t)p.B * sc, ty = (size_t)p.B * sy, td = (size_t)p.B * sd;
    const char* gx0 = reinterpret_cast<const char*>(p.gx[model] + (long long)(b0 + col0) * (2 * PG) + dir * PG + u);
    const char* c0p = reinterpret_cast<const char*>(p.c[model] + (long long)(b0 + col0) * (2 * PH) + dir * PH + u);
    const char* y0p = reinterpret_cast<const char*>(p.dy[model] + (long long)(b0 + col0) * (2 * PH) + dir * PH + u);
    char* d0p = reinterpret_cast<char*>(p.dg[model] + (long long)(b0 + col0) * (2 * PG) + dir * PG + u);
    // Everything the cell backward needs except the recurrent gradient streams through a register FIFO two groups (of 4
    // batch columns) deep, filled across step boundaries (7 loads per cell: dy, the four gates, c_t, c_{t-1}).
    constexpr int NG4 = NC / 4;                  // groups of 4 columns per thread
    constexpr int PF = (NG4 < 2 || C::BWD_EPI_WARPS == 16) ? 1 : 2;   // 16 warps: 96 registers per thread
    struct Cell { float dy, ig, fg, gg, og, ct, cp; };
    Cell cq[PF][4];
    auto load_group = [&](int tt, int g4, Cell* cc) {
      const int t_pf = dir ? tt + 1 : tt - 1;     // forward-order predecessor (c_{t-1})
      const bool has_prev = dir ? (tt < p.T - 1) : (tt > 0);
      const char* gp = gx0 + (size_t)tt * tg;
      const char* cp = c0p + (size_t)tt * tcb;
      const char* cpp = c0p + (size_t)(has_prev ? t_pf : tt) * tcb;
      const char* yp = y0p + (size_t)tt * ty;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = 4 * g4 + e;
        Cell c{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (col < nvalid) {
          const float* g = reinterpret_cast<const float*>(gp + (uint32_t)col * sg);
          c.ig = g[0]; c.fg = g[PH]; c.gg = g[2 * PH]; c.og = g[3 * PH];
          c.ct = *reinterpret_cast<const float*>(cp + (uint32_t)col * sc);
          c.cp = has_prev ? *reinterpret_cast<const float*>(cpp + (uint32_t)col * sc) : 0.f;
          c.dy = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(yp + (uint32_t)col * sy));
        }
        cc[e] = c;
      }
    };
    {
      const int t_first = dir ? 0 : p.T - 1;
#pragma unroll
      for (int g4 = 0; g4 < PF; ++g4) load_group(t_first, g4, cq[g4]);
    }
    for (int step = 0; step < p.T; ++step) {
      const int t = dir ? step : p.T - 1 - step;
      const int t_nxt = dir ? t + 1 : t - 1;
      const bool more = step + 1 < p.T;
      char* dgt = d0p + (size_t)t * td;
      if (step > 0) {
        mbar_wait(done_bar, (uint32_t)(step - 1) & 1u);
        tc_fence_after();
      }
#pragma unroll
      for (int g4 = 0; g4 < NG4; ++g4) {
        Cell cur[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) cur[e] = cq[g4 % PF][e];
        if (g4 + PF < NG4) load_group(t, g4 + PF, cq[g4 % PF]);
        else if (more) load_group(t_nxt, g4 + PF - NG4, cq[g4 % PF]);
        uint32_t acc[4] = {0u, 0u, 0u, 0u};
        if (step > 0) {
          tmem_ld4(trow + 4 * g4, acc);
          tmem_ld_wait();
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * g4 + e;
          if (j < nvalid) {
            const Cell c = cur[e];
            const float dh = c.dy + __uint_as_float(acc[e]);
            const float tc = tanh_fast(c.ct);
            const float dc = dcs[j] + dh * c.og * (1.f - tc * tc);
            dcs[j] = dc * c.fg;
            __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(dgt + (uint32_t)j * sd);
            dp[0] = __float2bfloat16(dc * c.gg * c.ig * (1.f - c.ig));
            dp[PH] = __float2bfloat16(dc * c.cp * c.fg * (1.f - c.fg));
            dp[2 * PH] = __float2bfloat16(dc * c.ig * (1.f - c.gg * c.gg));
            dp[3 * PH] = __float2bfloat16(dh * tc * c.og * (1.f - c.og));
          }
        }
      }
      fence_proxy_async_global();
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * C::BWD_EPI_WARPS) : "memory");
      if (threadIdx.x == 64) {
        __threadfence();
        red_release_gpu_add(flags + step, 1);
      }
      if (lane == 0) mbar_arrive(free_bar);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tm, TCOLS);
}


// =================================================================================================================
// backward, narrow batch tiles (NB <= 32): gate-stacked formulation
// =================================================================================================================
// dh[u, b] = sum_j W_hh[j, u] dgates[b, j] has K = 1536: 96 MMAs per step whatever the batch tile holds, and an MMA
// cannot be issued more often than every ~85 cycles.  Stacking the four gate blocks on BOTH output dimensions --
// D[(g, u), (g', b)] = sum_k W_hh[g*384 + k, u] dgates[b, g'*384 + k], K = 384 -- needs 48 MMAs of N = 4 NB columns
// (still at the issue floor for NB <= 32) and the wanted sum is the diagonal g == g'.  A = W_hh^T (a transposed copy made
// by the launcher) in the same [4 gates x 8 units] x 64 k layout per TMEM lane quarter as the forward kernel, so the four
// diagonal terms of a unit sit on four lanes of one warp: two shuffle rounds add them, and (as in the forward kernel) each
// of the four lanes then does the cell backward of one of four batch columns.
template <int NB, int UT>
__global__ void __launch_bounds__(64 + 32 * (UT / 4), 1)
lstm_seq_bwd_gs_kernel(const __grid_constant__ LstmSeqMaps maps, const LstmSeqParams p) {
  constexpr int KB = PH / 64;                 // 6 k-blocks per step
  constexpr int MT = UT / 32;                 // M tiles (4 gate blocks x 32 units each): 64 or 32 units per CTA, as forward
  constexpr int W_KB_BYTES = MT * 16384;
  constexpr int NN = 4 * NB;                  // MMA width: (gate, batch column)
  constexpr int STAGE_B = NN * 128;           // one [4 NB x 64] bf16 k-block
  // ring: with 32 units per CTA the weights take 96 KB, so a whole step's six k-blocks can be in flight
  constexpr int STAGES = UT == 32 ? (6 * STAGE_B <= 98304 ? 6 : 98304 / STAGE_B) : P_RING_BYTES / STAGE_B;
  constexpr int EPI_WARPS = 8 * MT;           // M tiles x 4 lane quarters x 2 column halves
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_w = smem;                        // [KB][M tiles][128 rows x 128 B]
  uint8_t* s_ring = smem + KB * W_KB_BYTES;   // [STAGES][4 gate blocks][NB rows x 128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_ring + STAGES * STAGE_B);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* w_bar = empty_bar + STAGES;
  uint64_t* done_bar = w_bar + 1;
  uint64_t* free_bar = done_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(free_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * UT, bt = blockIdx.y, rec = blockIdx.z;
  const int b0 = p.b_first + bt * NB;
  const int model = rec >> 1, dir = rec & 1;
  int* flags = p.flags + ((size_t)rec * p.nbt + bt) * p.T;
  constexpr uint32_t TCOLS = tmem_pow2(MT * NN);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(w_bar, 1);
    mbar_init(done_bar, 1);
    mbar_init(free_bar, EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ loader: W_hh^T slice once, then dgates_{t+1} per step
    if (elect_one()) {
      mbar_arrive_expect_tx(w_bar, KB * W_KB_BYTES);
      for (int kb = 0; kb < KB; ++kb)
        for (int mq = 0; mq < UT / 8; ++mq)
          tma_load_3d(&maps.w[rec], w_bar, s_w + kb * W_KB_BYTES + mq * 4096, kb * 64, u0 + mq * 8, 0);
    }
    __syncwarp();
    if (elect_one()) {
      uint32_t it = 0;
      for (int step = 1; step < p.T; ++step) {
        const int t = dir ? step : p.T - 1 - step;
        const int t_next = dir ? t - 1 : t + 1;
        wait_counter(flags + (step - 1), PH / UT);
        fence_proxy_async_global();
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], STAGE_B);
          for (int g = 0; g < 4; ++g)   // the k-block of every gate block: rows (g, b) of the B operand
            tma_load_3d(&maps.act[model], &full_bar[s], s_ring + s * STAGE_B + g * (NB * 128), dir * PG + g * PH + kb * 64, b0,
                        t_next);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t IDESC = umma_idesc(UMMA_BF16, 128, NN, 0, 0);
    if (elect_one()) mbar_wait(w_bar, 0);
    __syncwarp();
    uint32_t it = 0;
    for (int step = 1; step < p.T; ++step) {
      if (elect_one()) {
        mbar_wait(free_bar, (uint32_t)(step - 1) & 1u);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t s = (it + kb) % STAGES;
          mbar_wait(&full_bar[s], ((it + kb) / STAGES) & 1u);
          tc_fence_after();
          const uint32_t sb = smem_u32(s_ring + s * STAGE_B), sa = smem_u32(s_w + kb * W_KB_BYTES);
#pragma unroll
          for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(tm + m * NN, umma_desc_sw128(sa + m * 16384 + k * 32, 16, 1024),
                          umma_desc_sw128(sb + k * 32, 16, 1024), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&empty_bar[s]);
        }
        tc_commit(done_bar);
      }
      __syncwarp();
      it += KB;
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int ew = warp - 2;
    const int q = warp & 3, m = MT == 2 ? (ew >> 2) & 1 : 0, ch = MT == 2 ? ew >> 3 : ew >> 2;
    const int gl = lane >> 3, ul = lane & 7;             // this lane's gate block and unit inside the quarter
    const int u = u0 + (m * 4 + q) * 8 + ul;
    constexpr int NGRP = NB / 8;                         // groups of 4 batch columns handled by this warp
    const int colw = ch * (NB / 2);                      // its first column
    const uint32_t trow = tm + ((uint32_t)(q * 32) << 16) + m * NN + colw;
    float dcs[NGRP];
#pragma unroll
    for (int j = 0; j < NGRP; ++j) dcs[j] = 0.f;
    const int nvalid = min(NB, p.B - b0) - colw;         // valid columns of this warp's range (may be <= 0)
    constexpr uint32_t sg = 2u * PG * 4u, sc = 2u * PH * 4u, sy = 2u * PH * 2u, sd = 2u * PG * 2u;
    const size_t tg = (size_t)p.B * sg, tcb = (size_t)p.B * sc, ty = (size_t)p.B * sy, td = (size_t)p.B * sd;
    // this lane's cell of group j is (unit u, batch column b0 + colw + 4 j + gl)
    const char* gx0 = reinterpret_cast<const char*>(p.gx[model] + (long long)(b0 + colw + gl) * (2 * PG) + dir * PG + u);
    const char* c0p = reinterpret_cast<const char*>(p.c[model] + (long long)(b0 + colw + gl) * (2 * PH) + dir * PH + u);
    const char* y0p = reinterpret_cast<const char*>(p.dy[model] + (long long)(b0 + colw + gl) * (2 * PH) + dir * PH + u);
    char* d0p = reinterpret_cast<char*>(p.dg[model] + (long long)(b0 + colw + gl) * (2 * PG) + dir * PG + u);
    struct Cell { float dy, ig, fg, gg, og, ct, cp; };
    Cell cq[NGRP];
    auto load_cell = [&](int tt, int j) {
      Cell c{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (4 * j + gl < nvalid) {
        const int t_pf = dir ? tt + 1 : tt - 1;     // forward-order predecessor (c_{t-1})
        const bool has_prev = dir ? (tt < p.T - 1) : (tt > 0);
        const float* g = reinterpret_cast<const float*>(gx0 + (size_t)tt * tg + (uint32_t)(4 * j) * sg);
        c.ig = g[0]; c.fg = g[PH]; c.gg = g[2 * PH]; c.og = g[3 * PH];
        c.ct = *reinterpret_cast<const float*>(c0p + (size_t)tt * tcb + (uint32_t)(4 * j) * sc);
        c.cp = has_prev ? *reinterpret_cast<const float*>(c0p + (size_t)t_pf * tcb + (uint32_t)(4 * j) * sc) : 0.f;
        c.dy = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(y0p + (size_t)tt * ty + (uint32_t)(4 * j) * sy));
      }
      return c;
    };
    {
      const int t_first = dir ? 0 : p.T - 1;
#pragma unroll
      for (int j = 0; j < NGRP; ++j) cq[j] = load_cell(t_first, j);
    }
    for (int step = 0; step < p.T; ++step) {
      const int t = dir ? step : p.T - 1 - step;
      const int t_nxt = dir ? t + 1 : t - 1;
      const bool more = step + 1 < p.T;
      char* dgt = d0p + (size_t)t * td;
      if (step > 0) {
        mbar_wait(done_bar, (uint32_t)(step - 1) & 1u);
        tc_fence_after();
      }
#pragma unroll
      for (int j = 0; j < NGRP; ++j) {
        const Cell c = cq[j];
        if (more) cq[j] = load_cell(t_nxt, j);   // next step's operands are in flight during this step's arithmetic
        float dhr = 0.f;
        if (step > 0) {
          // the lane's own gate block of the four (g', b) column blocks: D[(gl, u), (gl, 4j .. 4j+3)]
          uint32_t a0[4], a1[4], a2[4], a3[4];
          tmem_ld4(trow + 0 * NB + 4 * j, a0);
          tmem_ld4(trow + 1 * NB + 4 * j, a1);
          tmem_ld4(trow + 2 * NB + 4 * j, a2);
          tmem_ld4(trow + 3 * NB + 4 * j, a3);
          tmem_ld_wait();
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            v[e] = __uint_as_float(gl == 0 ? a0[e] : gl == 1 ? a1[e] : gl == 2 ? a2[e] : a3[e]);
          // sum over the four gate lanes of the unit (lanes ul + 8 g): afterwards every lane holds dh for columns 4j..4j+3
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[e] += __shfl_xor_sync(0xffffffffu, v[e], 8);
            v[e] += __shfl_xor_sync(0xffffffffu, v[e], 16);
          }
          dhr = gl == 0 ? v[0] : gl == 1 ? v[1] : gl == 2 ? v[2] : v[3];   // this lane's column: 4j + gl
        }
        if (4 * j + gl < nvalid) {
          const float dh = c.dy + dhr;
          const float tc = tanh_fast(c.ct);
          const float dc = dcs[j] + dh * c.og * (1.f - tc * tc);
          dcs[j] = dc * c.fg;
          __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(dgt + (uint32_t)(4 * j) * sd);
          dp[0] = __float2bfloat16(dc * c.gg * c.ig * (1.f - c.ig));
          dp[PH] = __float2bfloat16(dc * c.cp * c.fg * (1.f - c.fg));
          dp[2 * PH] = __float2bfloat16(dc * c.ig * (1.f - c.gg * c.gg));
          dp[3 * PH] = __float2bfloat16(dh * tc * c.og * (1.f - c.og));
        }
      }
      fence_proxy_async_global();
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      if (threadIdx.x == 64) {
        __threadfence();
        red_release_gpu_add(flags + step, 1);
      }
      if (lane == 0) mbar_arrive(free_bar);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tm, TCOLS);
}

// W_hh [1536][384] bf16 -> W_hh^T [384][1536] (one recurrence per blockIdx.z), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
lstm_whh_transpose_kernel(const __nv_bfloat16* const w0, const __nv_bfloat16* const w1, const __nv_bfloat16* const w2,
                          const __nv_bfloat16* const w3, __nv_bfloat16* __restrict__ out) {
  __shared__ __nv_bfloat16 tile[32][33];
  const __nv_bfloat16* w = blockIdx.z == 0 ? w0 : blockIdx.z == 1 ? w1 : blockIdx.z == 2 ? w2 : w3;
  __nv_bfloat16* o = out + (size_t)blockIdx.z * PG * PH;
  const int j0 = blockIdx.y * 32, u0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = w[(size_t)(j0 + r) * PH + u0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8) o[(size_t)(u0 + r) * PG + j0 + tx] = tile[tx][r];
}

}  // namespace pe

// =================================================================================================================
// C-ABI
// =================================================================================================================
using namespace pe;

// mode: 0 forward, 1 backward (W_hh as stored, MN-major operand), 2 gate-stacked backward (w_hh = the transposed copies)
static int seq_maps(LstmSeqMaps* m, const void* const* act, int act_cols, int B, int T, int NB, const void* const* w_hh,
                    int mode) {
  for (int i = 0; i < 2; ++i) {
    uint64_t dims[3] = {(uint64_t)act_cols, (uint64_t)B, (uint64_t)T};   // time-major: [T][B][cols]
    uint64_t str[2] = {(uint64_t)act_cols * 2, (uint64_t)B * act_cols * 2};
    uint32_t box[3] = {64, (uint32_t)NB, 1};
    if (int rc = pe_host::encode_tmap(&m->act[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, act[i], dims, str, box)) return rc;
  }
  for (int r = 0; r < 4; ++r) {
    if (mode == 2) {  // W_hh^T [unit][gate * 384 + k]: the same (k, unit, gate) box as the forward kernel's
      uint64_t dims[3] = {(uint64_t)PH, (uint64_t)PH, 4};
      uint64_t str[2] = {(uint64_t)PG * 2, (uint64_t)PH * 2};
      uint32_t box[3] = {64, 8, 4};
      if (int rc = pe_host::encode_tmap(&m->w[r], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, w_hh[r], dims, str, box)) return rc;
    } else if (mode == 1) {
      uint64_t dims[2] = {(uint64_t)PH, (uint64_t)PG};
      uint64_t str[1] = {(uint64_t)PH * 2};
      uint32_t box[2] = {64, 64};
      if (int rc = pe_host::encode_tmap(&m->w[r], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_hh[r], dims, str, box)) return rc;
    } else {  // [gate][unit][k]: a box of 8 units x 4 gates fills one 32-lane TMEM quarter
      uint64_t dims[3] = {(uint64_t)PH, (uint64_t)PH, 4};
      uint64_t str[2] = {(uint64_t)PH * 2, (uint64_t)PH * PH * 2};
      uint32_t box[3] = {64, 8, 4};
      if (int rc = pe_host::encode_tmap(&m->w[r], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, w_hh[r], dims, str, box)) return rc;
    }
  }
  return PE_OK;
}

// Tiling of a launch.  A time step costs a CTA roughly in proportion to its batch-tile width NB (the h / dgates tile streams
// through a 32 KB ring, the epilogue is NB columns wide) and hardly depends on the number of batch tiles, which run side by
// side on their own CTAs -- so the narrowest tile whose batch tiles still fit ONE cooperative launch wins.  Forward: 32
// units per CTA (12 unit tiles x 4 recurrences x <= 3 batch tiles on 148 SMs; half the MMAs and half the epilogue per CTA
// and step) or 64 units (<= 6 batch tiles), whichever allows the cheaper tile.  Measured forward layer times (ms, four
// layer calls): UT = 32: 3.16 / 3.75 / 5.84 / 9.75 at NB = 16 / 32 / 64 / 128; UT = 64: 3.76 / 4.57 / 6.90 / 11.7.
// Backward (64 units): 6.9 / 8.5 / 11.4 / 20.2 ms at NB = 16 / 32 / 64 / 128.
struct SeqTiling {
  int ut, nb;
};
static const int kSeqNB[5] = {16, 32, 64, 96, 128};
static int narrowest_nb(int B, int max_tiles) {
  for (int i = 0; i < 4; ++i)
    if ((B + kSeqNB[i] - 1) / kSeqNB[i] <= max_tiles) return kSeqNB[i];
  return 128;
}
static SeqTiling seq_tiling(bool backward, int B) {
  static const int k_ut = getenv("PE_LSTM_UT") ? atoi(getenv("PE_LSTM_UT")) : 0;          // tuning knobs
  static const int k_fnb = getenv("PE_LSTM_FWD_NB") ? atoi(getenv("PE_LSTM_FWD_NB")) : 0;
  static const int k_bnb = getenv("PE_LSTM_BWD_NB") ? atoi(getenv("PE_LSTM_BWD_NB")) : 0;
  auto valid_nb = [](int v) { return v == 16 || v == 32 || v == 64 || v == 96 || v == 128; };
  const int sms = pe_host::num_sms();
  if (backward) {
    // gate-stacked kernel at 16- / 32-column tiles (measured backward layers, ms: 32 units x 16 columns 4.4, 64 x 16 5.1,
    // 64 x 32 8.1; the 64 / 128-column tiles of the plain kernel 11.4 / 20.2)
    static const int k_but = getenv("PE_LSTM_BWD_UT") ? atoi(getenv("PE_LSTM_BWD_UT")) : 0;
    const int t32 = sms / 48 > 0 ? sms / 48 : 1, t64 = sms / 24 > 0 ? sms / 24 : 1;
    SeqTiling t{64, narrowest_nb(B, t64)};
    if ((B + 15) / 16 <= t32) t = SeqTiling{32, 16};
    if (k_but == 32) t = SeqTiling{32, narrowest_nb(B, t32)};
    if (k_but == 64) t = SeqTiling{64, narrowest_nb(B, t64)};
    if (valid_nb(k_bnb)) t.nb = k_bnb;
    if (t.nb > 32) t.ut = 64;  // (only the gate-stacked kernel of the narrow tiles has the 32-unit variant)
    return t;
  }
  const int nb32 = narrowest_nb(B, sms / 48 > 0 ? sms / 48 : 1), nb64 = narrowest_nb(B, sms / 24 > 0 ? sms / 24 : 1);
  // relative cost of a step at (UT, NB), from the measurements above
  auto cost = [](int ut, int nb) {
    const float c32[5] = {3.16f, 3.75f, 5.84f, 7.8f, 9.75f}, c64[5] = {3.76f, 4.57f, 6.90f, 9.3f, 11.7f};  // (96: interpolated)
    const int i = nb == 16 ? 0 : nb == 32 ? 1 : nb == 64 ? 2 : nb == 96 ? 3 : 4;
    return ut == 32 ? c32[i] : c64[i];
  };
  // (a tile width whose batch tiles do not fit one launch is run in chunks: cost x chunks)
  auto total = [&](int ut, int nb) {
    const int max_bt = sms / (4 * (PH / ut)) > 0 ? sms / (4 * (PH / ut)) : 1;
    const int tiles = (B + nb - 1) / nb;
    return cost(ut, nb) * (float)((tiles + max_bt - 1) / max_bt);
  };
  SeqTiling t = total(32, nb32) <= total(64, nb64) ? SeqTiling{32, nb32} : SeqTiling{64, nb64};
  if (k_ut == 32 || k_ut == 64) t = SeqTiling{k_ut, k_ut == 32 ? nb32 : nb64};
  if (valid_nb(k_fnb)) t.nb = k_fnb;
  return t;
}

template <int NB>
static int launch_seq(bool backward, int ut, const LstmSeqMaps& maps, LstmSeqParams p, cudaStream_t st) {
  const void* fn = backward ? (const void*)lstm_seq_bwd_kernel<NB>
                            : (ut == 32 ? (const void*)lstm_seq_fwd_kernel<NB, 32> : (const void*)lstm_seq_fwd_kernel<NB, 64>);
  const size_t smem = PlCfg<NB>::SMEM;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PE_ERR_LAUNCH;
  void* args[2] = {(void*)&maps, (void*)&p};
  dim3 grid(PH / ut, p.nbt, 4);
  const int threads = backward ? PlCfg<NB>::BWD_THREADS : 64 + 32 * (ut / 4);
  if (cudaLaunchCooperativeKernel(fn, grid, dim3(threads), args, smem, st) != cudaSuccess) return PE_ERR_LAUNCH;
  return PE_OK;
}

template <int NB>
static int launch_seq_gs(int ut, const LstmSeqMaps& maps, LstmSeqParams p, cudaStream_t st) {
  const void* fn = ut == 32 ? (const void*)lstm_seq_bwd_gs_kernel<NB, 32> : (const void*)lstm_seq_bwd_gs_kernel<NB, 64>;
  const size_t smem = (size_t)P_W_BYTES + P_RING_BYTES + 256 + 1024;  // (32 units: 96 KB of weights + up to 96 KB of ring)
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PE_ERR_LAUNCH;
  void* args[2] = {(void*)&maps, (void*)&p};
  if (cudaLaunchCooperativeKernel(fn, dim3(PH / ut, p.nbt, 4), dim3(64 + 32 * (ut / 4)), args, smem, st) != cudaSuccess)
    return PE_ERR_LAUNCH;
  return PE_OK;
}

// bytes of the arrival counters at the head of the workspace (the transposed weights of the gate-stacked backward follow)
static size_t seq_flags_bytes(int T) {
  const int max_bt = pe_host::num_sms() / 24 > 0 ? pe_host::num_sms() / 24 : 1;
  return (((size_t)4 * max_bt * T * sizeof(int)) + 255) & ~(size_t)255;
}

static int run_seq(bool backward, LstmSeqParams p, const void* const* act, int act_cols, const void* const* w_hh,
                   int* flags, size_t flags_bytes, cudaStream_t st) {
  const int B = p.B, T = p.T;
  const SeqTiling tl = seq_tiling(backward, B);
  const int NB = tl.nb, ut = tl.ut;
  const int nbt_total = (B + NB - 1) / NB;
  const int max_bt = pe_host::num_sms() / (4 * (PH / ut));   // unit tiles x 4 recurrences per batch tile, one CTA per SM
  if (max_bt < 1) return PE_ERR_ARCH;
  const int nbt_launch = nbt_total < max_bt ? nbt_total : max_bt;
  if (!flags || flags_bytes < (size_t)4 * nbt_launch * T * sizeof(int)) return PE_ERR_WORKSPACE;
  // narrow batch tiles: gate-stacked backward on a transposed copy of the weights (kept behind the counters)
  static const bool gs_off = getenv("PE_LSTM_BWD_GS") && atoi(getenv("PE_LSTM_BWD_GS")) == 0;  // tuning knob
  const bool gs = backward && NB <= 32 && !gs_off && flags_bytes >= seq_flags_bytes(T) + (size_t)4 * PG * PH * 2;
  const void* wt[4];
  if (gs) {
    __nv_bfloat16* wbase = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(flags) + seq_flags_bytes(T));
    lstm_whh_transpose_kernel<<<dim3(PH / 32, PG / 32, 4), 256, 0, st>>>(
        (const __nv_bfloat16*)w_hh[0], (const __nv_bfloat16*)w_hh[1], (const __nv_bfloat16*)w_hh[2],
        (const __nv_bfloat16*)w_hh[3], wbase);
    for (int r = 0; r < 4; ++r) wt[r] = wbase + (size_t)r * PG * PH;
  }
  LstmSeqMaps maps;
  if (int rc = seq_maps(&maps, act, act_cols, B, T, NB, gs ? wt : w_hh, gs ? 2 : (backward ? 1 : 0))) return rc;
  for (int bt0 = 0; bt0 < nbt_total; bt0 += nbt_launch) {   // batch tiles are independent: chunk them if B is huge
    p.nbt = nbt_total - bt0 < nbt_launch ? nbt_total - bt0 : nbt_launch;
    p.b_first = bt0 * NB;
    p.flags = flags;
    if (cudaMemsetAsync(flags, 0, (size_t)4 * p.nbt * T * sizeof(int), st) != cudaSuccess) return PE_ERR_LAUNCH;
    int rc;
    if (gs) {
      rc = NB == 16 ? launch_seq_gs<16>(ut, maps, p, st) : launch_seq_gs<32>(ut, maps, p, st);
      if (rc) return rc;
      continue;
    }
    switch (NB) {
      case 16: rc = launch_seq<16>(backward, ut, maps, p, st); break;
      case 32: rc = launch_seq<32>(backward, ut, maps, p, st); break;
      case 64: rc = launch_seq<64>(backward, ut, maps, p, st); break;
      case 96: rc = launch_seq<96>(backward, ut, maps, p, st); break;
      default: rc = launch_seq<128>(backward, ut, maps, p, st); break;
    }
    if (rc) return rc;
  }
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}

extern "C" int pe_lstm_seq_fwd(int B, int T, int hidden, float* const* gx, float* const* c, void* const* y,
                               const void* const* w_hh, const float* const* b_ih, const float* const* b_hh,
                               void* workspace, size_t workspace_bytes, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (hidden != PH || B <= 0 || T <= 0 || !gx || !c || !y || !w_hh || !b_ih || !b_hh) return PE_ERR_BAD_SHAPE;
  LstmSeqParams p{};
  p.B = B; p.T = T;
  for (int i = 0; i < 2; ++i) {
    p.gx[i] = gx[i]; p.c[i] = c[i]; p.y[i] = (__nv_bfloat16*)y[i];
  }
  for (int r = 0; r < 4; ++r) {
    p.b_ih[r] = b_ih[r]; p.b_hh[r] = b_hh[r];
  }
  const void* act[2] = {y[0], y[1]};
  return run_seq(false, p, act, 2 * PH, w_hh, (int*)workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int pe_lstm_seq_bwd(int B, int T, int hidden, const float* const* gates, const float* const* c,
                               const void* const* dy, void* const* dg, const void* const* w_hh, void* workspace,
                               size_t workspace_bytes, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (hidden != PH || B <= 0 || T <= 0 || !gates || !c || !dy || !dg || !w_hh) return PE_ERR_BAD_SHAPE;
  LstmSeqParams p{};
  p.B = B; p.T = T;
  for (int i = 0; i < 2; ++i) {
    p.gx[i] = const_cast<float*>(gates[i]); p.c[i] = const_cast<float*>(c[i]);
    p.dy[i] = (const __nv_bfloat16*)dy[i]; p.dg[i] = (__nv_bfloat16*)dg[i];
  }
  const void* act[2] = {dg[0], dg[1]};
  return run_seq(true, p, act, 2 * PG, w_hh, (int*)workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}
