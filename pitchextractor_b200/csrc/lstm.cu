// BiLSTM recurrence (reference model.py:218-228 -> torch.nn.LSTM, gate order i,f,g,o, torch/nn/modules/rnn.py:842-847).
// One launch per time step covers the four independent recurrences of a layer (2 sequence models x 2 directions):
//   forward : gates[b, 4x64 units] = h_{t-1} W_hh^T on tcgen05 (accumulators in TMEM), the cell update runs in the
//             epilogue (adds the input projection + both biases, sigmoid / tanh, c_t, h_t) and writes h_t, c_t and the
//             activated gates (kept for the backward pass) -- no intermediate gate tensor round trip;
//   backward: dh_rec = dgates_{t+1} W_hh on tcgen05, the epilogue adds the gradient from above, back-propagates the
//             cell (dc carried in a small fp32 state) and writes the pre-activation gate gradients.
// Weight / input gradients are large token-contraction GEMMs run afterwards by the tile engine.
#include "common.cuh"
#include "../../include/pitchextractor_b200.h"

PE_USES_STEP_SALT()

namespace pe {

constexpr int LH = 384;         // hidden size
constexpr int LG = 4 * LH;      // gate rows per direction
constexpr int L_EPI_WARPS = 8;   // two per TMEM lane quarter: each takes half of the unit chunks
constexpr int L_THREADS = 64 + 32 * L_EPI_WARPS;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

struct LstmStepParams {
  int B, T, step, first;
  // per model (0 = classifier, 1 = detector)
  float* gx[2];              // [T][B][2*LG] fp32 (time-major): input projection in, activated gates out
  float* c[2];               // [T][B][2*LH] fp32 cell state
  __nv_bfloat16* y[2];       // [T][B][2*LH] bf16 hidden state
  // per recurrence r = model*2 + dir
  const float* b_ih[4];
  const float* b_hh[4];
  // backward only
  const __nv_bfloat16* dy[2];  // [T][B][2*LH] gradient w.r.t. y from above
  __nv_bfloat16* dg[2];        // [T][B][2*LG] pre-activation gate gradients
  float* dc[2];                // [B][2*LH] running dL/dc
};

// sigmoid / tanh through one ex2 + one fast reciprocal each (relative error ~1e-6, far below the bf16 hidden state)
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return fmaf(2.f, __fdividef(1.f, 1.f + __expf(-2.f * x)), -1.f); }

struct LstmMaps {
  CUtensorMap act[2];  // per model: forward: y (dims 2*LH, B, T); backward: dg (dims 2*LG, B, T)
  CUtensorMap w[4];    // per recurrence: W_hh [LG][LH] bf16
};

// ---------------------------------------------------------------------------------------------------------------
// forward step: tile = 128 batch rows x (4 gates x 64 units), K = 384
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(L_THREADS, 1)
lstm_step_fwd_kernel(const __grid_constant__ LstmMaps maps, const LstmStepParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int STAGES = 4, A_B = 16384, B_B = 32768, STAGE_B = A_B + B_B, KB = LH / 64;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_B);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * 64, b0 = blockIdx.y * 128, rec = blockIdx.z;
  const int model = rec >> 1, dir = rec & 1;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // let the next step's prologue start early
  const int t = dir ? p.T - 1 - p.step : p.step;
  const int t_prev = dir ? t + 1 : t - 1;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: the prologue above overlapped the previous time step; everything below reads what
  // that step wrote
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (!p.first) {
    // warp-uniform loops, one elected lane issues the TMA / tcgen05 instructions (operands stay in uniform registers)
    if (warp == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        uint8_t* sa = smem + s * STAGE_B;
        uint8_t* sb = sa + A_B;
        if (elect_one()) {
          mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], STAGE_B);
          tma_load_3d(&maps.act[model], &full_bar[s], sa, dir * LH + kb * 64, b0, t_prev);
          for (int g = 0; g < 4; ++g) tma_load_2d(&maps.w[rec], &full_bar[s], sb + g * 8192, kb * 64, g * LH + u0);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      constexpr uint32_t IDESC = umma_idesc(UMMA_BF16, 128, 256, 0, 0);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t sa = smem_u32(smem + s * STAGE_B), sb = sa + A_B;
        if (elect_one()) {
          mbar_wait(&full_bar[s], (kb / STAGES) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_base, umma_desc_sw128(sa + k * 32, 16, 1024), umma_desc_sw128(sb + k * 32, 16, 1024), IDESC,
                        (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&empty_bar[s]);
          if (kb == KB - 1) tc_commit(done_bar);
        }
        __syncwarp();
      }
    }
  }
  if (warp >= 2) {
    const int q = warp & 3;
    const int b = b0 + q * 32 + lane;
    const bool row_ok = b < p.B;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const long long tok = (long long)t * p.B + b;            // time-major tokens: row = t * B + b
    const long long tok_prev = (long long)t_prev * p.B + b;
    const int pair = (warp - 2) >> 2;
    float pre[4][16], cprev[16];
    // input projection + both biases and c_{t-1} of one 16-unit chunk (independent of this step's MMA)
    auto load_inputs = [&](int uc) {
      const int u = u0 + uc * 16;
      const float* gxp = p.gx[model] + tok * (2 * LG) + dir * LG + u;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float* bi = p.b_ih[rec] + g * LH + u;
        const float* bh = p.b_hh[rec] + g * LH + u;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 x = row_ok ? *reinterpret_cast<const float4*>(gxp + g * LH + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 y1 = __ldg(reinterpret_cast<const float4*>(bi + j));
          const float4 y2 = __ldg(reinterpret_cast<const float4*>(bh + j));
          pre[g][j] = x.x + y1.x + y2.x;
          pre[g][j + 1] = x.y + y1.y + y2.y;
          pre[g][j + 2] = x.z + y1.z + y2.z;
          pre[g][j + 3] = x.w + y1.w + y2.w;
        }
      }
      const float* cp = p.c[model] + tok_prev * (2 * LH) + dir * LH + u;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!p.first && row_ok) x = *reinterpret_cast<const float4*>(cp + j);
        cprev[j] = x.x; cprev[j + 1] = x.y; cprev[j + 2] = x.z; cprev[j + 3] = x.w;
      }
    };
    // recurrent contribution from TMEM, cell update, stores
    auto finish = [&](int uc) {
      if (!p.first) {
        uint32_t acc[4][16];
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld16(trow + g * 64 + uc * 16, acc[g]);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int j = 0; j < 16; ++j) pre[g][j] += __uint_as_float(acc[g][j]);
      }
      if (!row_ok) return;
      const int u = u0 + uc * 16;
      float cn[16], hn[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float ig = sigmoidf_(pre[0][j]), fg = sigmoidf_(pre[1][j]), gg = tanhf_(pre[2][j]), og = sigmoidf_(pre[3][j]);
        pre[0][j] = ig; pre[1][j] = fg; pre[2][j] = gg; pre[3][j] = og;
        cn[j] = fmaf(fg, cprev[j], ig * gg);
        hn[j] = og * tanhf_(cn[j]);
      }
      float* gxp = p.gx[model] + tok * (2 * LG) + dir * LG + u;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(gxp + g * LH + j) = make_float4(pre[g][j], pre[g][j + 1], pre[g][j + 2], pre[g][j + 3]);
      float* cw = p.c[model] + tok * (2 * LH) + dir * LH + u;
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(cw + j) = make_float4(cn[j], cn[j + 1], cn[j + 2], cn[j + 3]);
      __nv_bfloat16* yw = p.y[model] + tok * (2 * LH) + dir * LH + u;
#pragma unroll
      for (int j = 0; j < 16; j += 8)
        *reinterpret_cast<uint4*>(yw + j) = make_uint4(pack_bf16(hn[j], hn[j + 1]), pack_bf16(hn[j + 2], hn[j + 3]),
                                                       pack_bf16(hn[j + 4], hn[j + 5]), pack_bf16(hn[j + 6], hn[j + 7]));
    };
    load_inputs(2 * pair);  // overlaps the TMA + MMA main loop of this step
    if (!p.first) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    finish(2 * pair);
    load_inputs(2 * pair + 1);
    finish(2 * pair + 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------------------
// backward step: tile = 128 batch rows x 64 units, K = 1536 (gate rows), B operand MN-major (W_hh as stored)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(L_THREADS, 1)
lstm_step_bwd_kernel(const __grid_constant__ LstmMaps maps, const LstmStepParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int STAGES = 8, A_B = 16384, B_B = 8192, STAGE_B = A_B + B_B, KB = LG / 64;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_B);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * 64, b0 = blockIdx.y * 128, rec = blockIdx.z;
  const int model = rec >> 1, dir = rec & 1;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // the backward pass walks each direction's own time order in reverse
  const int t = dir ? p.step : p.T - 1 - p.step;
  const int t_next = dir ? t - 1 : t + 1;   // processed one launch earlier
  const int t_pf = dir ? t + 1 : t - 1;     // forward-order predecessor (c_{t-1}, h_{t-1})
  const bool has_prev = dir ? (t < p.T - 1) : (t > 0);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (!p.first) {
    if (warp == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        uint8_t* sa = smem + s * STAGE_B;
        if (elect_one()) {
          mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], STAGE_B);
          tma_load_3d(&maps.act[model], &full_bar[s], sa, dir * LG + kb * 64, b0, t_next);
          tma_load_2d(&maps.w[rec], &full_bar[s], sa + A_B, u0, kb * 64);  // [64 gate rows][64 units]: MN-major B
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      constexpr uint32_t IDESC = umma_idesc(UMMA_BF16, 128, 64, 0, 1);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t sa = smem_u32(smem + s * STAGE_B), sb = sa + A_B;
        if (elect_one()) {
          mbar_wait(&full_bar[s], (kb / STAGES) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_base, umma_desc_sw128(sa + k * 32, 16, 1024), umma_desc_sw128(sb + k * 2048, 8192, 1024),
                        IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          // N = 64 MMAs are short: one commit per two k-blocks would do, but the ring is only STAGES deep
          tc_commit(&empty_bar[s]);
          if (kb == KB - 1) tc_commit(done_bar);
        }
        __syncwarp();
      }
    }
  }
  if (warp >= 2) {
    const int q = warp & 3;
    const int b = b0 + q * 32 + lane;
    const bool row_ok = b < p.B;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const long long tok = (long long)t * p.B + b;            // time-major tokens: row = t * B + b
    const long long tok_pf = (long long)t_pf * p.B + b;
    const int pair = (warp - 2) >> 2;
    float dh[16], gate[4][16], ct[16], cp[16], dcs[16];
    // everything the cell backward needs except the recurrent gradient (independent of this step's MMA)
    auto load_inputs = [&](int uc) {
      const int u = u0 + uc * 16;
      const __nv_bfloat16* dyp = p.dy[model] + tok * (2 * LH) + dir * LH + u;
#pragma unroll
      for (int j = 0; j < 16; j += 8) {
        uint4 raw = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok) raw = *reinterpret_cast<const uint4*>(dyp + j);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(h2[k]);
          dh[j + 2 * k] = f.x;
          dh[j + 2 * k + 1] = f.y;
        }
      }
      const float* gxp = p.gx[model] + tok * (2 * LG) + dir * LG + u;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 x = row_ok ? *reinterpret_cast<const float4*>(gxp + g * LH + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          gate[g][j] = x.x; gate[g][j + 1] = x.y; gate[g][j + 2] = x.z; gate[g][j + 3] = x.w;
        }
      const float* ctp = p.c[model] + tok * (2 * LH) + dir * LH + u;
      const float* cpp = p.c[model] + tok_pf * (2 * LH) + dir * LH + u;
      const float* dcp = p.dc[model] + (long long)b * (2 * LH) + dir * LH + u;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x, z = x;
        if (row_ok) {
          x = *reinterpret_cast<const float4*>(ctp + j);
          if (has_prev) y = *reinterpret_cast<const float4*>(cpp + j);
          if (!p.first) z = *reinterpret_cast<const float4*>(dcp + j);
        }
        ct[j] = x.x; ct[j + 1] = x.y; ct[j + 2] = x.z; ct[j + 3] = x.w;
        cp[j] = y.x; cp[j + 1] = y.y; cp[j + 2] = y.z; cp[j + 3] = y.w;
        dcs[j] = z.x; dcs[j + 1] = z.y; dcs[j + 2] = z.z; dcs[j + 3] = z.w;
      }
    };
    auto finish = [&](int uc) {
      if (!p.first) {
        uint32_t acc[16];
        tmem_ld16(trow + uc * 16, acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) dh[j] += __uint_as_float(acc[j]);
      }
      if (!row_ok) return;
      const int u = u0 + uc * 16;
      float dpre[4][16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float ig = gate[0][j], fg = gate[1][j], gg = gate[2][j], og = gate[3][j];
        const float tc = tanhf_(ct[j]);
        const float dc = dcs[j] + dh[j] * og * (1.f - tc * tc);
        dpre[0][j] = dc * gg * ig * (1.f - ig);
        dpre[1][j] = dc * cp[j] * fg * (1.f - fg);
        dpre[2][j] = dc * ig * (1.f - gg * gg);
        dpre[3][j] = dh[j] * tc * og * (1.f - og);
        dcs[j] = dc * fg;
      }
      float* dcp = p.dc[model] + (long long)b * (2 * LH) + dir * LH + u;
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dcp + j) = make_float4(dcs[j], dcs[j + 1], dcs[j + 2], dcs[j + 3]);
      __nv_bfloat16* dgp = p.dg[model] + tok * (2 * LG) + dir * LG + u;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 16; j += 8)
          *reinterpret_cast<uint4*>(dgp + g * LH + j) =
              make_uint4(pack_bf16(dpre[g][j], dpre[g][j + 1]), pack_bf16(dpre[g][j + 2], dpre[g][j + 3]),
                         pack_bf16(dpre[g][j + 4], dpre[g][j + 5]), pack_bf16(dpre[g][j + 6], dpre[g][j + 7]));
    };
    load_inputs(2 * pair);  // overlaps the TMA + MMA main loop of this step
    if (!p.first) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    finish(2 * pair);
    load_inputs(2 * pair + 1);
    finish(2 * pair + 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 64);
}

// y = dropout(x) over a [rows][cols] bf16 matrix (inter-layer LSTM dropout; the same call back-propagates)
__global__ void __launch_bounds__(256)
dropout_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n8,
                    unsigned thresh, float scale, unsigned long long seed) {
  seed = pe_salted(seed);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 raw = *reinterpret_cast<const uint4*>(x + i * 8);
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
  const uint32_t km = dropout_keep8(seed, (unsigned long long)i, thresh);
  float f[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 v = __bfloat1622float2(h2[k]);
    f[2 * k] = ((km >> (2 * k)) & 1u) ? v.x * scale : 0.f;
    f[2 * k + 1] = ((km >> (2 * k + 1)) & 1u) ? v.y * scale : 0.f;
  }
  *reinterpret_cast<uint4*>(y + i * 8) =
      make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

}  // namespace pe

// =================================================================================================================
// C-ABI
// =================================================================================================================
using namespace pe;

static int lstm_maps(LstmMaps* m, const void* const* act, int act_cols, int B, int T, const void* const* w_hh,
                     bool w_mn) {
  for (int i = 0; i < 2; ++i) {
    uint64_t dims[3] = {(uint64_t)act_cols, (uint64_t)B, (uint64_t)T};   // time-major: [T][B][cols]
    uint64_t str[2] = {(uint64_t)act_cols * 2, (uint64_t)B * act_cols * 2};
    uint32_t box[3] = {64, 128, 1};
    if (int rc = pe_host::encode_tmap(&m->act[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, act[i], dims, str, box))
      return rc;
  }
  for (int r = 0; r < 4; ++r) {
    uint64_t dims[2] = {(uint64_t)LH, (uint64_t)LG};
    uint64_t str[1] = {(uint64_t)LH * 2};
    uint32_t box[2] = {64, 64};
    (void)w_mn;
    if (int rc = pe_host::encode_tmap(&m->w[r], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_hh[r], dims, str, box))
      return rc;
  }
  return PE_OK;
}

extern "C" int pe_lstm_steps_fwd(int B, int T, int hidden, int step_begin, int step_end, float* const* gx,
                                 float* const* c, void* const* y, const void* const* w_hh, const float* const* b_ih,
                                 const float* const* b_hh, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (hidden != LH || B <= 0 || T <= 0 || step_begin < 0 || step_end > T || step_begin >= step_end || !gx || !c || !y ||
      !w_hh || !b_ih || !b_hh)
    return PE_ERR_BAD_SHAPE;
  LstmMaps maps;
  const void* act[2] = {y[0], y[1]};
  if (int rc = lstm_maps(&maps, act, 2 * LH, B, T, w_hh, false)) return rc;
  LstmStepParams p{};
  p.B = B; p.T = T;
  for (int i = 0; i < 2; ++i) {
    p.gx[i] = gx[i]; p.c[i] = c[i]; p.y[i] = (__nv_bfloat16*)y[i];
  }
  for (int r = 0; r < 4; ++r) {
    p.b_ih[r] = b_ih[r]; p.b_hh[r] = b_hh[r];
  }
  const size_t smem = 4 * (16384 + 32768) + 9 * 8 + 16 + 1024;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(lstm_step_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return PE_ERR_LAUNCH;
    attr = true;
  }
  dim3 grid(LH / 64, (B + 127) / 128, 4);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(L_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr_pdl[1];
  attr_pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr_pdl[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr_pdl;
  cfg.numAttrs = 1;
  for (int s = step_begin; s < step_end; ++s) {  // the time steps are dependent launches on the same stream
    p.step = s;
    p.first = s == 0;
    if (cudaLaunchKernelEx(&cfg, lstm_step_fwd_kernel, maps, p) != cudaSuccess) return PE_ERR_LAUNCH;
  }
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}

extern "C" int pe_lstm_steps_bwd(int B, int T, int hidden, int step_begin, int step_end, const float* const* gates,
                                 const float* const* c, const void* const* dy, void* const* dg, float* const* dc,
                                 const void* const* w_hh, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (hidden != LH || B <= 0 || T <= 0 || step_begin < 0 || step_end > T || step_begin >= step_end || !gates || !c ||
      !dy || !dg || !dc || !w_hh)
    return PE_ERR_BAD_SHAPE;
  LstmMaps maps;
  const void* act[2] = {dg[0], dg[1]};
  if (int rc = lstm_maps(&maps, act, 2 * LG, B, T, w_hh, true)) return rc;
  LstmStepParams p{};
  p.B = B; p.T = T;
  for (int i = 0; i < 2; ++i) {
    p.gx[i] = const_cast<float*>(gates[i]); p.c[i] = const_cast<float*>(c[i]);
    p.dy[i] = (const __nv_bfloat16*)dy[i]; p.dg[i] = (__nv_bfloat16*)dg[i]; p.dc[i] = dc[i];
  }
  const size_t smem = 8 * (16384 + 8192) + 17 * 8 + 16 + 1024;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(lstm_step_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return PE_ERR_LAUNCH;
    attr = true;
  }
  dim3 grid(LH / 64, (B + 127) / 128, 4);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(L_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr_pdl[1];
  attr_pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr_pdl[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr_pdl;
  cfg.numAttrs = 1;
  for (int s = step_begin; s < step_end; ++s) {
    p.step = s;
    p.first = s == 0;
    if (cudaLaunchKernelEx(&cfg, lstm_step_bwd_kernel, maps, p) != cudaSuccess) return PE_ERR_LAUNCH;
  }
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}

extern "C" int pe_dropout_bf16(const void* x, void* y, long long n, unsigned drop_thresh, float drop_scale,
                               unsigned long long seed, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!x || !y || n <= 0 || (n % 8)) return PE_ERR_BAD_SHAPE;
  const long long n8 = n / 8;
  dropout_bf16_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n8, drop_thresh, drop_scale, seed);
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}
