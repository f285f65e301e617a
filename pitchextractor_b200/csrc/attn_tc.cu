// Fused multi-head attention for the JDCNet Transformer blocks on tcgen05 tensor cores (T = 192 frames, head_dim 64).
// Work is cut into small independent units so that TWO CTAs fit on every SM (<= 113 KB of shared memory and 256 TMEM
// columns each): while one CTA waits for its TMA loads or tensor-core products, the other runs its softmax arithmetic.
//   forward   unit = (query tile of 128 rows, head, item): S = Q_t K^T (192 columns in TMEM), softmax in registers,
//             P written over the dead Q / K tiles as the swizzled K-major operand, O = P V.
//   backward  four units per (head, item), each walking the other sequence dimension in six 32-wide chunks whose
//             products are double-buffered in TMEM:
//             A_t (query tile t): S_c, dP_c -> dS operand chunk -> dQ_t += dS_c K_c
//             B_t (key tile t):   S^T_c, dP^T_c (recomputed key-major, so nothing is transposed) -> P~^T_c, dS^T_c
//                                 operand chunks -> dV_t += P~^T_c dO_c, dK_t += dS^T_c Q_c
// Q/K/V (and dO) tiles arrive by TMA straight out of the packed qkv activation in 64-row boxes; all products run as
// tcgen05.mma with accumulators in TMEM; the bf16 P / dS operands are written back to shared memory in the
// 128-byte-swizzled K-major operand layout.  nn.MultiheadAttention inside nn.TransformerEncoderLayer (reference
// model.py:231-239) is what this replaces.
#include "common.cuh"
#include "../../include/pitchextractor_b200.h"

PE_USES_STEP_SALT()

namespace pe {

constexpr int AT = 192;            // sequence length (frames)
constexpr int AD = 64;             // head dim
constexpr int BOX_B = 64 * 128;    // one TMA box: 64 rows x 64 bf16 = 8 KB
constexpr float LOG2E = 1.4426950408889634f;
constexpr int ATT_THREADS = 256;   // 8 warps: 4 TMEM lane quarters x 2 column halves

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of element (row r, column k) inside a [128 x 64*n] bf16 K-major SW128 operand (64-column k-blocks of 16 KB)
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
__device__ __forceinline__ void store32_bf16(__nv_bfloat16* op, const uint32_t* v, float scale) {
#pragma unroll
  for (int j = 0; j < 32; j += 8)
    *reinterpret_cast<uint4*>(op + j) =
        make_uint4(pack_bf16(__uint_as_float(v[j]) * scale, __uint_as_float(v[j + 1]) * scale),
                   pack_bf16(__uint_as_float(v[j + 2]) * scale, __uint_as_float(v[j + 3]) * scale),
                   pack_bf16(__uint_as_float(v[j + 4]) * scale, __uint_as_float(v[j + 5]) * scale),
                   pack_bf16(__uint_as_float(v[j + 6]) * scale, __uint_as_float(v[j + 7]) * scale));
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
// make generic-proxy smem writes and finished TMEM reads visible / ordered before warp 0 issues the next MMAs
__device__ __forceinline__ void attn_handoff() {
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
}

// D[128 x N] (+)= A[128 x 64] B^T, A/B K-major tiles with 128-byte rows (B: N rows): 4 UMMA_K steps
__device__ __forceinline__ void mma_k64(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, bool acc = false) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    tc_mma_bf16(d, umma_desc_sw128(a + k * 32, 16, 1024), umma_desc_sw128(b + k * 32, 16, 1024), idesc, acc || k > 0);
}
// D[128 x 64] (+)= A[128 x 64*NKB] B, A = K-major operand (NKB k-blocks of 16 KB), B = MN-major [64*NKB rows x 64] tile
template <int NKB>
__device__ __forceinline__ void mma_kn(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, bool acc) {
#pragma unroll
  for (int j = 0; j < 4 * NKB; ++j)
    tc_mma_bf16(d, umma_desc_sw128(a + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024),
                umma_desc_sw128(b + j * 2048, 8192, 1024), idesc, acc || j > 0);
}
// rows [row0, row0 + 64 n) x 64 columns starting at column col of a token tensor -> n consecutive 8 KB boxes
__device__ __forceinline__ void load_rows(const CUtensorMap* m, uint64_t* bar, uint8_t* dst, int col, int row0, int n) {
  for (int i = 0; i < n; ++i) tma_load_2d(m, bar, dst + i * BOX_B, col, row0 + 64 * i);
}

// -------------------------------------------------------------------------------------------------
// forward: one CTA per (query tile, head, item)
// -------------------------------------------------------------------------------------------------
constexpr int AFW_SMEM = 3 * BOX_B + 2 * BOX_B + BOX_B + 3 * BOX_B + 2 * 128 * 4 + 64;  // K | Q_t | pad | V | red | bars

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, int H, unsigned drop_thresh, float drop_scale,
                   unsigned long long seed, __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* sK = smem;                      // [192 x 64]
  uint8_t* sQ = smem + 3 * BOX_B;          // [128 x 64]: query rows t*128 .. (rows past the item belong to the next one)
  uint8_t* sP = smem;                      // [128 x 192] operand, written over K / Q_t (+ 8 KB) once S is complete
  uint8_t* sV = smem + 6 * BOX_B;          // [192 x 64]
  float* red = reinterpret_cast<float*>(sV + 3 * BOX_B);  // [2 halves][128] partial row max / sum exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = H * AD;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  seed = pe_salted(seed);
  if (warp == 0) {
    if (elect_one()) {  // loads first: they run while the TMEM allocation (which may wait for a co-resident CTA) settles
      mbar_arrive_expect_tx(&bars[0], 8 * BOX_B);
      load_rows(&tm_qkv, &bars[0], sQ, h * AD, b * AT + t * 128, 2);
      load_rows(&tm_qkv, &bars[0], sK, D + h * AD, b * AT, 3);
      load_rows(&tm_qkv, &bars[0], sV, 2 * D + h * AD, b * AT, 3);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  AttnSync sync{&bars[1], 0};

  constexpr uint32_t IDESC_S = umma_idesc(UMMA_BF16, 128, AT, 0, 0);
  constexpr uint32_t IDESC_O = umma_idesc(UMMA_BF16, 128, AD, 0, 1);
  if (warp == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    if (elect_one()) {
      mma_k64(tm, smem_u32(sQ), smem_u32(sK), IDESC_S);  // S rows t*128 .. t*128+127 (>= 192 unused)
      tc_commit(&bars[1]);
    }
    __syncwarp();
  }
  sync.wait();

  const int quarter = warp & 3, half = warp >> 2;
  const int r = quarter * 32 + lane;
  const int q = t * 128 + r;
  const bool live = (t * 128 + quarter * 32) < AT;  // warp-uniform: the second tile holds 64 real query rows
  const float kscale = 0.125f * LOG2E;
  const uint32_t trow = tm + ((uint32_t)(quarter * 32) << 16);
  float inv_l = 0.f;
  {
    uint32_t v[32];
    float mx = -INFINITY;
    if (live) {
      for (int c = half * 3; c < half * 3 + 3; ++c) {
        tmem_ld32(trow + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
    }
    red[half * 128 + r] = mx;
    __syncthreads();  // (also: every warp has left the S MMA wait, K / Q_t are dead from here on)
    mx = fmaxf(red[r], red[128 + r]);
    __syncthreads();
    float l = 0.f;
    if (live) {
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
        oper_store32(sP, r, c * 32, p);
      }
    }
    red[half * 128 + r] = l;
    __syncthreads();
    l = red[r] + red[128 + r];
    if (live) {
      inv_l = 1.f / l;
      if (half == 0) lse[((long long)b * H + h) * AT + q] = mx * 0.125f + __logf(l);
    }
  }
  attn_handoff();
  if (warp == 0) {
    tc_fence_after();
    if (elect_one()) {
      mma_kn<3>(tm + AT, smem_u32(sP), smem_u32(sV), IDESC_O, false);
      tc_commit(&bars[1]);
    }
    __syncwarp();
  }
  sync.wait();
  if (live) {
    uint32_t v[32];
    tmem_ld32(trow + AT + half * 32, v);
    tmem_ld_wait();
    store32_bf16(ctx + ((long long)b * AT + q) * D + h * AD + half * 32, v, inv_l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

// -------------------------------------------------------------------------------------------------
// backward: units A_0, A_1 (query tiles -> dQ) and B_0, B_1 (key tiles -> dK, dV) of one (head, item).
// Both kinds walk the "other" sequence dimension in six 32-wide chunks; the chunk products are double-buffered in TMEM
// (the tensor pipe computes chunk c+1 / c+2 while the warps work on chunk c), the outputs accumulate in TMEM.
// Warps 0-7 do the arithmetic, warp 8 issues TMA loads and MMAs; the two sides talk through mbarriers only (a CTA-wide
// barrier per chunk had the arithmetic warps wait for the issuing warp: 26 % of all stall samples).
// delta_q = rowsum(dO * O) comes from a small pre-pass (each of the four units of a head needs it).
// -------------------------------------------------------------------------------------------------
constexpr int ABW_SMEM = 12 * BOX_B + 3 * AT * 4 + 64;  // 96 KB of tiles | lse, delta, row keys | barriers
constexpr int ABW_CH = 32, ABW_NCH = AT / ABW_CH;       // chunk width, chunks per unit
constexpr int ABW_THREADS = ATT_THREADS + 32;

// delta[(b*H + h)*T + t] = sum_d dO[b*T + t][h*64 + d] * O[b*T + t][h*64 + d]; eight lanes per (token, head), each one
// 16-byte load of both tensors (a warp reads 512 contiguous bytes of a token row), summed by three shuffles
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ ctx, const __nv_bfloat16* __restrict__ dctx, int rows, int H,
                  float* __restrict__ delta) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one per 8 bf16
  const long long total = (long long)rows * H * 8;
  float acc = 0.f;
  if (i < total) {
    const uint4 a = reinterpret_cast<const uint4*>(ctx)[i], c = reinterpret_cast<const uint4*>(dctx)[i];
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* hc = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = __bfloat1622float2(ha[k]), y = __bfloat1622float2(hc[k]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (i < total && (threadIdx.x & 7) == 0) {
    const long long rh = i >> 3;  // row * H + h
    const int row = (int)(rh / H), h = (int)(rh - (long long)row * H);
    const int b = row / AT, t = row - b * AT;
    delta[((long long)b * H + h) * AT + t] = acc;
  }
}

__global__ void __launch_bounds__(ABW_THREADS, 2)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, int H,
                   unsigned drop_thresh, float drop_scale, unsigned long long seed, const float* __restrict__ lse,
                   const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  float* sLse = reinterpret_cast<float*>(smem + 12 * BOX_B);  // [192] lse * log2(e)          (unit B only)
  float* sDel = sLse + AT;                                    // [192] rowsum(dO * O)
  uint32_t* sKey = reinterpret_cast<uint32_t*>(sDel + AT);    // [192] dropout row keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKey + AT);
  uint64_t* bar_tiles = bars;       // tiles loaded (TMA)
  uint64_t* bar_prod = bars + 1;    // [2] chunk products of TMEM buffer k complete (tcgen05.commit)
  uint64_t* bar_oper = bars + 3;    // [2] operand chunk written, TMEM buffer k read (8 warps arrive)
  uint64_t* bar_used = bars + 5;    // unit B: the operand buffer has been consumed (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int unit = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const bool unitA = unit < 2;
  const int t = unit & 1;
  const int D = H * AD;
  const long long ld3 = 3LL * D;
  const int row0 = b * AT;
  const long long bh = (long long)b * H + h;

  // shared-memory map (8 KB boxes): X_t 0-1 | Y_t 2-3 | U 4-6 | W 7-9 | operand 10-11
  //   A: X = Q, Y = dO (this query tile), U = K, W = V (all keys);   operand = dS chunks (column halves alternate)
  //   B: X = K, Y = V (this key tile),    U = Q, W = dO (all queries); operand = P~^T chunk | dS^T chunk
  uint8_t* sX = smem;
  uint8_t* sY = smem + 2 * BOX_B;
  uint8_t* sU = smem + 4 * BOX_B;
  uint8_t* sW = smem + 7 * BOX_B;
  uint8_t* sOp = smem + 10 * BOX_B;

  if (tid == 0) {
    mbar_init(bar_tiles, 1);
    mbar_init(&bar_prod[0], 1);
    mbar_init(&bar_prod[1], 1);
    mbar_init(&bar_oper[0], 8);
    mbar_init(&bar_oper[1], 8);
    mbar_init(bar_used, 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  seed = pe_salted(seed);
  if (warp == 8) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_tiles, 10 * BOX_B);
      if (unitA) {
        load_rows(&tm_qkv, bar_tiles, sX, h * AD, row0 + t * 128, 2);
        load_rows(&tm_do, bar_tiles, sY, h * AD, row0 + t * 128, 2);
        load_rows(&tm_qkv, bar_tiles, sU, D + h * AD, row0, 3);
        load_rows(&tm_qkv, bar_tiles, sW, 2 * D + h * AD, row0, 3);
      } else {
        load_rows(&tm_qkv, bar_tiles, sX, D + h * AD, row0 + t * 128, 2);
        load_rows(&tm_qkv, bar_tiles, sY, 2 * D + h * AD, row0 + t * 128, 2);
        load_rows(&tm_qkv, bar_tiles, sU, h * AD, row0, 3);
        load_rows(&tm_do, bar_tiles, sW, h * AD, row0, 3);
      }
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  } else if (!unitA && tid < AT) {
    sDel[tid] = delta[bh * AT + tid];
    sLse[tid] = lse[bh * AT + tid] * LOG2E;
    sKey[tid] = attn_row_key(seed, (unsigned long long)bh * AT + tid);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;

  constexpr uint32_t IDESC_C = umma_idesc(UMMA_BF16, 128, ABW_CH, 0, 0);  // [128 x 32] score chunk, operands K-major
  constexpr uint32_t IDESC_O = umma_idesc(UMMA_BF16, 128, AD, 0, 1);      // [128 x 64] output, B MN-major
  // TMEM columns: chunk buffer k at 64k: first product (S / S^T) +0, second (dP / dP^T) +32; outputs at 128 and 192

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA / MMA issuer
    const uint32_t aX = smem_u32(sX), aY = smem_u32(sY), aU = smem_u32(sU), aW = smem_u32(sW), aOp = smem_u32(sOp);
    // chunk c: first = X_t U_c^T, second = Y_t W_c^T   (32 rows of U / W = 4 KB)
    auto issue_chunk = [&](int c) {
      const uint32_t d = tm + 64u * (uint32_t)(c & 1);
      mma_k64(d, aX, aU + (uint32_t)c * 4096u, IDESC_C);
      mma_k64(d + 32, aY, aW + (uint32_t)c * 4096u, IDESC_C);
    };
    // out[128 x 64] (+)= operand columns [col0, col0 + 32) x rows [32 c, 32 c + 32) of the MN-major tile `bt`
    auto issue_partial = [&](uint32_t out_col, int col0, uint32_t bt, int c) {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        tc_mma_bf16(tm + out_col, umma_desc_sw128(aOp + (uint32_t)(col0 * 2 + k * 32), 16, 1024),
                    umma_desc_sw128(bt + (uint32_t)c * 4096u + (uint32_t)k * 2048u, 8192, 1024), IDESC_O,
                    (c > 0 || k > 0) ? 1u : 0u);
    };
    mbar_wait(bar_tiles, 0);
    tc_fence_after();
    if (elect_one()) {
      issue_chunk(0);
      tc_commit(&bar_prod[0]);
      issue_chunk(1);
      tc_commit(&bar_prod[1]);
    }
    __syncwarp();
    for (int c = 0; c < ABW_NCH; ++c) {
      mbar_wait(&bar_oper[c & 1], (uint32_t)((c >> 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        if (unitA) {
          issue_partial(128, (c & 1) * 32, aU, c);   // dQ_t += dS_c K_c
        } else {
          issue_partial(128, 0, aW, c);              // dV_t += P~^T_c dO_c
          issue_partial(192, 32, aU, c);             // dK_t += dS^T_c Q_c
          tc_commit(bar_used);
        }
        if (c + 2 < ABW_NCH) issue_chunk(c + 2);
        tc_commit(&bar_prod[c & 1]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ arithmetic warps
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const int row_t = t * 128 + r;                      // query (A) or key (B) index of this thread's TMEM lane
    const bool live = (t * 128 + quarter * 32) < AT;    // warp-uniform
    const uint32_t lane_base = tm + ((uint32_t)(quarter * 32) << 16);
    const float kscale = 0.125f * LOG2E;
    float lq = 0.f, dq_delta = 0.f;
    uint32_t rkey = 0;
    if (unitA && live) {  // this thread's own query row
      lq = lse[bh * AT + row_t] * LOG2E;
      dq_delta = delta[bh * AT + row_t];
      rkey = attn_row_key(seed, (unsigned long long)bh * AT + row_t);
    }
    for (int c = 0; c < ABW_NCH; ++c) {
      mbar_wait(&bar_prod[c & 1], (uint32_t)((c >> 1) & 1));
      tc_fence_after();
      const uint32_t cb = lane_base + 64u * (uint32_t)(c & 1) + (uint32_t)(half * 16);
      if (unitA) {
        if (live) {
          uint32_t s[16], d[16];
          tmem_ld16(cb, s);
          tmem_ld16(cb + 32, d);
          tmem_ld_wait();
          float ds[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p = ex2(__uint_as_float(s[j]) * kscale - lq);
            float dp = __uint_as_float(d[j]);
            if (drop_thresh)
              dp = attn_drop_hash(rkey, (uint32_t)(c * ABW_CH + half * 16 + j)) < drop_thresh ? dp * drop_scale : 0.f;
            ds[j] = 0.125f * p * (dp - dq_delta);
          }
          // chunk c uses operand columns [32 (c & 1), +32): the product that read them (chunk c - 2) has completed
          oper_store16(sOp, r, (c & 1) * 32 + half * 16, ds);
        }
      } else {
        float pt[16], dst[16];
        if (live) {
          uint32_t s[16], d[16];
          tmem_ld16(cb, s);
          tmem_ld16(cb + 32, d);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int q = c * ABW_CH + half * 16 + j;
            const float p = ex2(__uint_as_float(s[j]) * kscale - sLse[q]);
            float dp = __uint_as_float(d[j]);
            float pd = p;
            if (drop_thresh) {
              const bool keep = attn_drop_hash(sKey[q], (uint32_t)row_t) < drop_thresh;
              dp = keep ? dp * drop_scale : 0.f;
              pd = keep ? p * drop_scale : 0.f;
            }
            pt[j] = pd;
            dst[j] = 0.125f * p * (dp - sDel[q]);
          }
        }
        // the single operand buffer (P~^T in columns 0-31, dS^T in 32-63) is free once chunk c-1's products have read it
        if (c > 0) mbar_wait(bar_used, (uint32_t)((c - 1) & 1));
        if (live) {
          oper_store16(sOp, r, half * 16, pt);
          oper_store16(sOp, r, 32 + half * 16, dst);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_oper[c & 1]);
    }
    // the last commit (chunk 5's products, buffer 1, fourth completion) covers every earlier product
    mbar_wait(&bar_prod[1], 1);
    tc_fence_after();
    if (live) {
      uint32_t v[32];
      if (unitA) {
        tmem_ld32(lane_base + 128 + half * 32, v);
        tmem_ld_wait();
        store32_bf16(dqkv + ((long long)row0 + row_t) * ld3 + h * AD + half * 32, v, 1.f);
      } else {
#pragma unroll
        for (int which = 0; which < 2; ++which) {  // 0: dV, 1: dK
          tmem_ld32(lane_base + 128 + which * 64 + half * 32, v);
          tmem_ld_wait();
          store32_bf16(dqkv + ((long long)row0 + row_t) * ld3 + (which == 0 ? 2 * D : D) + h * AD + half * 32, v, 1.f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 256);
}

}  // namespace pe

// =================================================================================================
// host
// =================================================================================================
static int token_tmap(CUtensorMap* m, const void* base, long long rows, int cols) {
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)cols * 2};
  uint32_t box[2] = {64, 64};
  return pe_host::encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box);
}

template <typename K>
static bool attn_attrs(K kernel, int smem) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess &&
         cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) ==
             cudaSuccess;
}

int pe_attn_fwd_tc(const void* qkv, int B, int H, unsigned drop_thresh, float drop_scale, unsigned long long seed,
                   void* ctx, float* lse, cudaStream_t stream) {
  CUtensorMap tq;
  if (int rc = token_tmap(&tq, qkv, (long long)B * pe::AT, 3 * H * pe::AD)) return rc;
  static bool attr = false;
  if (!attr) {
    if (!attn_attrs(pe::attn_fwd_tc_kernel, pe::AFW_SMEM)) return PE_ERR_LAUNCH;
    attr = true;
  }
  pe_host::launch(pe::attn_fwd_tc_kernel, dim3(2, H, B), dim3(pe::ATT_THREADS), pe::AFW_SMEM, stream, tq, H, drop_thresh,
                  drop_scale, seed, (__nv_bfloat16*)ctx, lse);
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}

int pe_attn_bwd_tc(const void* qkv, const void* ctx, const void* dctx, const float* lse, int B, int H,
                   unsigned drop_thresh, float drop_scale, unsigned long long seed, void* dqkv, float* delta,
                   cudaStream_t stream) {
  CUtensorMap tq, td;
  if (int rc = token_tmap(&tq, qkv, (long long)B * pe::AT, 3 * H * pe::AD)) return rc;
  if (int rc = token_tmap(&td, dctx, (long long)B * pe::AT, H * pe::AD)) return rc;
  static bool attr = false;
  if (!attr) {
    if (!attn_attrs(pe::attn_bwd_tc_kernel, pe::ABW_SMEM)) return PE_ERR_LAUNCH;
    attr = true;
  }
  const int rows = B * pe::AT;
  pe_host::launch(pe::attn_delta_kernel, dim3((unsigned)(((long long)rows * H * 8 + 255) / 256)), dim3(256), 0, stream, (const __nv_bfloat16*)ctx,
                  (const __nv_bfloat16*)dctx, rows, H, delta);
  pe_host::launch(pe::attn_bwd_tc_kernel, dim3(4, H, B), dim3(pe::ABW_THREADS), pe::ABW_SMEM, stream, tq, td, H,
                  drop_thresh, drop_scale, seed, (const float*)lse, (const float*)delta, (__nv_bfloat16*)dqkv);
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}
