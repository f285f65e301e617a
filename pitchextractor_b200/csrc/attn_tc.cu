// Fused multi-head attention for the JDCNet Transformer blocks on tcgen05 tensor cores (T = 192 frames, head_dim 64):
// one CTA per (head, batch item); Q/K/V (and dO) tiles arrive by TMA straight out of the packed qkv activation;
// S = QK^T, P V, and all five backward products run as tcgen05.mma with accumulators in TMEM; the softmax /
// softmax-backward arithmetic runs on the TMEM rows in registers and the bf16 P / dS operands are written back to
// shared memory in the 128-byte-swizzled K-major operand layout.  nn.MultiheadAttention inside
// nn.TransformerEncoderLayer (reference model.py:231-239) is what this replaces.
#include "common.cuh"
#include "../../include/pitchextractor_b200.h"

PE_USES_STEP_SALT()

namespace pe {

constexpr int AT = 192;            // sequence length (frames)
constexpr int AD = 64;             // head dim
constexpr int TILE_B = AT * 128;   // one [192 x 64] bf16 tile: 24 KB
constexpr int OPER_B = 3 * 16384;  // one [128 x 192] bf16 A-operand (3 k-blocks of [128 x 64]): 48 KB
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of element (row r, column k) inside a [128 x 192] bf16 K-major SW128 operand
__device__ __forceinline__ uint32_t oper_off(int r, int k) {
  const int kb = k >> 6, kin = k & 63;
  const int unit = (kin >> 3) ^ (r & 7);
  return (uint32_t)(kb * 16384 + r * 128 + unit * 16 + (kin & 7) * 2);
}
// write 32 consecutive columns [k0, k0+32) of row r (k0 multiple of 32) as bf16
__device__ __forceinline__ void oper_store32(uint8_t* base, int r, int k0, const float* v) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint4 pk = make_uint4(pack_bf16(v[8 * u], v[8 * u + 1]), pack_bf16(v[8 * u + 2], v[8 * u + 3]),
                                pack_bf16(v[8 * u + 4], v[8 * u + 5]), pack_bf16(v[8 * u + 6], v[8 * u + 7]));
    *reinterpret_cast<uint4*>(base + oper_off(r, k0 + 8 * u)) = pk;
  }
}

// write 16 consecutive columns [k0, k0+16) of row r (k0 multiple of 16) as bf16
__device__ __forceinline__ void oper_store16(uint8_t* base, int r, int k0, const float* v) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint4 pk = make_uint4(pack_bf16(v[8 * u], v[8 * u + 1]), pack_bf16(v[8 * u + 2], v[8 * u + 3]),
                                pack_bf16(v[8 * u + 4], v[8 * u + 5]), pack_bf16(v[8 * u + 6], v[8 * u + 7]));
    *reinterpret_cast<uint4*>(base + oper_off(r, k0 + 8 * u)) = pk;
  }
}
__device__ __forceinline__ void store16_bf16(__nv_bfloat16* op, const uint32_t* v) {
#pragma unroll
  for (int j = 0; j < 16; j += 8)
    *reinterpret_cast<uint4*>(op + j) =
        make_uint4(pack_bf16(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                   pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])),
                   pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])),
                   pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
}

struct AttnSync {
  uint64_t* bar;
  uint32_t phase;
  __device__ __forceinline__ void wait() {
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
  }
};
// make generic-proxy smem writes and finished TMEM reads visible / ordered before thread 0 issues the next MMAs
__device__ __forceinline__ void attn_handoff() {
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
}

// D[128 x N] (+)= A[128 x 64] B^T, A/B K-major tiles with 128-byte rows: 4 UMMA_K steps
__device__ __forceinline__ void mma_k64(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    tc_mma_bf16(d, umma_desc_sw128(a + k * 32, 16, 1024), umma_desc_sw128(b + k * 32, 16, 1024), idesc, k > 0);
}
// D[128 x 64] = A[128 x 192] B, A = K-major operand (3 k-blocks), B = MN-major [192 rows x 64] tile: 12 steps
__device__ __forceinline__ void mma_k192(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc) {
#pragma unroll
  for (int j = 0; j < 12; ++j)
    tc_mma_bf16(d, umma_desc_sw128(a + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024),
                umma_desc_sw128(b + j * 2048, 8192, 1024), idesc, j > 0);
}

// -------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------
constexpr int AFW_THREADS = 512;  // 16 warps: 8 per query tile (4 TMEM lane quarters x 2 column halves)

__global__ void __launch_bounds__(AFW_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, int H, unsigned drop_thresh, float drop_scale,
                   unsigned long long seed, __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse) {
  seed = pe_salted(seed);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                 // Q | K | V tiles, contiguous: rows past 192 of a tile read the next tile
  uint8_t* sK = smem + TILE_B;
  uint8_t* sV = smem + 2 * TILE_B;
  uint8_t* sP = smem + 3 * TILE_B + 8192;  // two [128 x 192] operands (query tiles 0 and 1)
  float* red = reinterpret_cast<float*>(sP + 2 * OPER_B);  // [2 tiles][2 halves][128] partial row max / sum exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 512);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * AD;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  AttnSync sync{&bars[1], 0};

  constexpr uint32_t IDESC_S = umma_idesc(UMMA_BF16, 128, AT, 0, 0);
  constexpr uint32_t IDESC_O = umma_idesc(UMMA_BF16, 128, AD, 0, 1);
  // single-thread instructions are issued by an elected lane of warp 0 from warp-uniform code (see elect_one())
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars[0], 3 * TILE_B);
      tma_load_2d(&tm_qkv, &bars[0], sQ, h * AD, b * AT);
      tma_load_2d(&tm_qkv, &bars[0], sK, D + h * AD, b * AT);
      tma_load_2d(&tm_qkv, &bars[0], sV, 2 * D + h * AD, b * AT);
    }
    __syncwarp();
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    if (elect_one()) {
      mma_k64(tm + 0, smem_u32(sQ), smem_u32(sK), IDESC_S);            // S rows   0..127
      mma_k64(tm + AT, smem_u32(sQ) + 16384, smem_u32(sK), IDESC_S);   // S rows 128..255 (>= 192 unused)
      tc_commit(&bars[1]);
    }
    __syncwarp();
  }
  sync.wait();

  // both query tiles are processed at once: warps 0-7 own rows 0..127, warps 8-15 rows 128..191 (+ unused)
  const int quarter = warp & 3, half = (warp >> 2) & 1, t = warp >> 3;
  const int r = quarter * 32 + lane;
  const float kscale = 0.125f * LOG2E;
  float* redt = red + t * 256;
  float inv_l;
  {
    const int q = t * 128 + r;
    const uint32_t trow = tm + ((uint32_t)(quarter * 32) << 16) + t * AT;
    uint32_t v[32];
    float mx = -INFINITY;
    for (int c = half * 3; c < half * 3 + 3; ++c) {
      tmem_ld32(trow + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    redt[half * 128 + r] = mx;
    __syncthreads();
    mx = fmaxf(redt[r], redt[128 + r]);
    __syncthreads();
    float l = 0.f;
    const uint32_t rkey = attn_row_key(seed, (unsigned long long)(b * H + h) * AT + q);
    for (int c = half * 3; c < half * 3 + 3; ++c) {
      tmem_ld32(trow + c * 32, v);
      tmem_ld_wait();
      float p[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        p[j] = ex2((__uint_as_float(v[j]) - mx) * kscale);
        l += p[j];
      }
      if (drop_thresh) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          p[j] = attn_drop_hash(rkey, (uint32_t)(c * 32 + j)) < drop_thresh ? p[j] * drop_scale : 0.f;
      }
      oper_store32(sP + t * OPER_B, r, c * 32, p);
    }
    redt[half * 128 + r] = l;
    __syncthreads();
    l = redt[r] + redt[128 + r];
    inv_l = 1.f / l;
    if (half == 0 && q < AT) lse[((long long)b * H + h) * AT + q] = mx * 0.125f + __logf(l);
  }
  attn_handoff();
  if (warp == 0) {
    tc_fence_after();
    if (elect_one()) {
      mma_k192(tm + 2 * AT, smem_u32(sP), smem_u32(sV), IDESC_O);
      mma_k192(tm + 2 * AT + AD, smem_u32(sP) + OPER_B, smem_u32(sV), IDESC_O);
      tc_commit(&bars[1]);
    }
    __syncwarp();
  }
  sync.wait();
  {
    const int q = t * 128 + r;
    uint32_t v[32];
    tmem_ld32(tm + ((uint32_t)(quarter * 32) << 16) + 2 * AT + t * AD + half * 32, v);
    tmem_ld_wait();
    if (q < AT) {
      __nv_bfloat16* op = ctx + ((long long)b * AT + q) * D + h * AD + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        *reinterpret_cast<uint4*>(op + j) =
            make_uint4(pack_bf16(__uint_as_float(v[j]) * inv_l, __uint_as_float(v[j + 1]) * inv_l),
                       pack_bf16(__uint_as_float(v[j + 2]) * inv_l, __uint_as_float(v[j + 3]) * inv_l),
                       pack_bf16(__uint_as_float(v[j + 4]) * inv_l, __uint_as_float(v[j + 5]) * inv_l),
                       pack_bf16(__uint_as_float(v[j + 6]) * inv_l, __uint_as_float(v[j + 7]) * inv_l));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// -------------------------------------------------------------------------------------------------
// backward: phase A (query-major) -> dQ;  phase B (key-major, S^T recomputed) -> dK, dV
// -------------------------------------------------------------------------------------------------
constexpr int ABW_THREADS = 512;  // 16 warps: four per TMEM lane quarter, 48 score columns each

__global__ void __launch_bounds__(ABW_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, int H,
                   unsigned drop_thresh, float drop_scale, unsigned long long seed,
                   const __nv_bfloat16* __restrict__ ctx, const __nv_bfloat16* __restrict__ dctx,
                   const float* __restrict__ lse, __nv_bfloat16* __restrict__ dqkv) {
  seed = pe_salted(seed);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_B;
  uint8_t* sV = smem + 2 * TILE_B;
  uint8_t* sDO = smem + 3 * TILE_B;
  uint8_t* sA = smem + 4 * TILE_B + 8192;  // dS (phase A) / P~^T (phase B)
  uint8_t* sB = sA + OPER_B;               // dS^T (phase B)
  float* sLse = reinterpret_cast<float*>(sB + OPER_B);  // [192] lse * log2(e)
  float* sDel = sLse + AT;                               // [192] rowsum(dO * O)
  uint32_t* sKey = reinterpret_cast<uint32_t*>(sDel + AT);  // [192] dropout row keys (phase B walks the mask by column)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKey + AT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * AD;
  const long long ld3 = 3LL * D;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  AttnSync sync{&bars[1], 0};
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars[0], 4 * TILE_B);
      tma_load_2d(&tm_qkv, &bars[0], sQ, h * AD, b * AT);
      tma_load_2d(&tm_qkv, &bars[0], sK, D + h * AD, b * AT);
      tma_load_2d(&tm_qkv, &bars[0], sV, 2 * D + h * AD, b * AT);
      tma_load_2d(&tm_do, &bars[0], sDO, h * AD, b * AT);
    }
    __syncwarp();
  }
  // delta_q = sum_d dO[q][d] * O[q][d]; lse in log2 units
  if (tid < AT) {
    const long long row = (long long)b * AT + tid;
    const uint4* po = reinterpret_cast<const uint4*>(ctx + row * D + h * AD);
    const uint4* pd = reinterpret_cast<const uint4*>(dctx + row * D + h * AD);
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint4 a = po[u], c = pd[u];
      const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* hc = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = __bfloat1622float2(ha[i]), y = __bfloat1622float2(hc[i]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
    }
    sDel[tid] = acc;
    sLse[tid] = lse[((long long)b * H + h) * AT + tid] * LOG2E;
    sKey[tid] = attn_row_key(seed, (unsigned long long)(b * H + h) * AT + tid);
  }
  __syncthreads();

  constexpr uint32_t IDESC_S = umma_idesc(UMMA_BF16, 128, AT, 0, 0);
  constexpr uint32_t IDESC_O = umma_idesc(UMMA_BF16, 128, AD, 0, 1);
  const int quarter = warp & 3, part = warp >> 2;  // part 0..3: score columns [48*part, 48*part+48)
  const int r = quarter * 32 + lane;
  const uint32_t lane_base = tm + ((uint32_t)(quarter * 32) << 16);
  const float kscale = 0.125f * LOG2E;
  const unsigned long long row_bh = (unsigned long long)(b * H + h) * AT;
  if (warp == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
  }

  // ---------------- phase A: per query tile, S and dP in TMEM -> dS operand -> dQ
  for (int t = 0; t < 2; ++t) {
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        mma_k64(tm + 0, smem_u32(sQ) + t * 16384, smem_u32(sK), IDESC_S);     // S_t  = Q_t K^T
        mma_k64(tm + AT, smem_u32(sDO) + t * 16384, smem_u32(sV), IDESC_S);   // dP_t = dO_t V^T
        tc_commit(&bars[1]);
      }
      __syncwarp();
    }
    sync.wait();
    const int q = t * 128 + r;
    const int qc = q < AT ? q : AT - 1;
    const float lq = sLse[qc], dq_delta = sDel[qc];
    const uint32_t rkey = attn_row_key(seed, row_bh + q);
    for (int c = part * 3; c < part * 3 + 3; ++c) {  // 16-column chunks
      uint32_t s[16], d[16];
      tmem_ld16(lane_base + c * 16, s);
      tmem_ld16(lane_base + AT + c * 16, d);
      tmem_ld_wait();
      float ds[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p = ex2(__uint_as_float(s[j]) * kscale - lq);
        float dp = __uint_as_float(d[j]);
        if (drop_thresh)
          dp = attn_drop_hash(rkey, (uint32_t)(c * 16 + j)) < drop_thresh ? dp * drop_scale : 0.f;
        ds[j] = 0.125f * p * (dp - dq_delta);
      }
      oper_store16(sA, r, c * 16, ds);
    }
    attn_handoff();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        mma_k192(tm + 2 * AT, smem_u32(sA), smem_u32(sK), IDESC_O);            // dQ_t = dS_t K
        tc_commit(&bars[1]);
      }
      __syncwarp();
    }
    sync.wait();
    {
      uint32_t v[16];
      tmem_ld16(lane_base + 2 * AT + part * 16, v);
      tmem_ld_wait();
      if (q < AT) store16_bf16(dqkv + ((long long)b * AT + q) * ld3 + h * AD + part * 16, v);
    }
    tc_fence_before();
    __syncthreads();
  }

  // ---------------- phase B: per key tile, S^T and dP^T in TMEM -> P~^T, dS^T operands -> dV, dK
  for (int t = 0; t < 2; ++t) {
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        mma_k64(tm + 0, smem_u32(sK) + t * 16384, smem_u32(sQ), IDESC_S);     // S^T_t  = K_t Q^T
        mma_k64(tm + AT, smem_u32(sV) + t * 16384, smem_u32(sDO), IDESC_S);   // dP^T_t = V_t dO^T
        tc_commit(&bars[1]);
      }
      __syncwarp();
    }
    sync.wait();
    const int jkey = t * 128 + r;
    for (int c = part * 3; c < part * 3 + 3; ++c) {
      uint32_t s[16], d[16];
      tmem_ld16(lane_base + c * 16, s);
      tmem_ld16(lane_base + AT + c * 16, d);
      tmem_ld_wait();
      float pt[16], dst[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int q = c * 16 + j;
        const float p = ex2(__uint_as_float(s[j]) * kscale - sLse[q]);
        float dp = __uint_as_float(d[j]);
        float pd = p;
        if (drop_thresh) {
          const bool keep = attn_drop_hash(sKey[q], (uint32_t)jkey) < drop_thresh;
          dp = keep ? dp * drop_scale : 0.f;
          pd = keep ? p * drop_scale : 0.f;
        }
        pt[j] = pd;
        dst[j] = 0.125f * p * (dp - sDel[q]);
      }
      oper_store16(sA, r, c * 16, pt);
      oper_store16(sB, r, c * 16, dst);
    }
    attn_handoff();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        mma_k192(tm + 2 * AT, smem_u32(sA), smem_u32(sDO), IDESC_O);           // dV_t = P~^T_t dO
        mma_k192(tm + 2 * AT + AD, smem_u32(sB), smem_u32(sQ), IDESC_O);       // dK_t = dS^T_t Q
        tc_commit(&bars[1]);
      }
      __syncwarp();
    }
    sync.wait();
#pragma unroll
    for (int which = 0; which < 2; ++which) {  // 0: dV, 1: dK
      uint32_t v[16];
      tmem_ld16(lane_base + 2 * AT + which * AD + part * 16, v);
      tmem_ld_wait();
      if (jkey < AT)
        store16_bf16(dqkv + ((long long)b * AT + jkey) * ld3 + (which == 0 ? 2 * D : D) + h * AD + part * 16, v);
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tm, 512);
}

}  // namespace pe

// =================================================================================================
// host
// =================================================================================================
static int token_tmap(CUtensorMap* m, const void* base, long long rows, int cols) {
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)cols * 2};
  uint32_t box[2] = {64, (uint32_t)pe::AT};
  return pe_host::encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box);
}

int pe_attn_fwd_tc(const void* qkv, int B, int H, unsigned drop_thresh, float drop_scale, unsigned long long seed,
                   void* ctx, float* lse, cudaStream_t stream) {
  CUtensorMap tq;
  if (int rc = token_tmap(&tq, qkv, (long long)B * pe::AT, 3 * H * pe::AD)) return rc;
  const size_t smem = 3 * pe::TILE_B + 8192 + 2 * pe::OPER_B + 512 * 4 + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(pe::attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess)
      return PE_ERR_LAUNCH;
    attr = true;
  }
  pe::attn_fwd_tc_kernel<<<dim3(H, B), pe::AFW_THREADS, smem, stream>>>(tq, H, drop_thresh, drop_scale, seed,
                                                            (__nv_bfloat16*)ctx, lse);
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}

int pe_attn_bwd_tc(const void* qkv, const void* ctx, const void* dctx, const float* lse, int B, int H,
                   unsigned drop_thresh, float drop_scale, unsigned long long seed, void* dqkv, cudaStream_t stream) {
  CUtensorMap tq, td;
  if (int rc = token_tmap(&tq, qkv, (long long)B * pe::AT, 3 * H * pe::AD)) return rc;
  if (int rc = token_tmap(&td, dctx, (long long)B * pe::AT, H * pe::AD)) return rc;
  const size_t smem = 4 * pe::TILE_B + 8192 + 2 * pe::OPER_B + 3 * pe::AT * 4 + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(pe::attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess)
      return PE_ERR_LAUNCH;
    attr = true;
  }
  pe::attn_bwd_tc_kernel<<<dim3(H, B), pe::ABW_THREADS, smem, stream>>>(tq, td, H, drop_thresh, drop_scale, seed,
                                                            (const __nv_bfloat16*)ctx, (const __nv_bfloat16*)dctx, lse,
                                                            (__nv_bfloat16*)dqkv);
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}
