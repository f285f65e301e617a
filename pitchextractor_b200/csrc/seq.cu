// Sequence-model side passes (reference model.py:178-256 via torch.nn.TransformerEncoderLayer / LayerNorm, and the
// heads + losses of model.py:67-70,96-117 / trainer.py:237-239): LayerNorm forward / backward (with fused
// positional-encoding add, dropout-masked copy and bias-gradient column sums), column sums, softmax attention
// (fp32 SIMT formulation, v0) forward / backward, and the fused pitch / voicing heads with SmoothL1 + BCE losses.
#include "common.cuh"
#include "../../include/pitchextractor_b200.h"

PE_USES_STEP_SALT()

namespace pe {

__device__ __forceinline__ void ld8b(const __nv_bfloat16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(h[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ void ld8f(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8b(__nv_bfloat16* p, const float* f) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over D = 256 * NCH features, one warp per row; lane owns columns i*256 + lane*8 .. +7.
// Input is fp32 (pre-LN residual sum) or bf16 (+ positional encoding row t = m % T).
// ---------------------------------------------------------------------------------------------
template <int NCH>
__device__ __forceinline__ void ln_load_row(const float* xf, const __nv_bfloat16* xb, const float* pe, int T,
                                            long long m, int lane, float (*v)[8]) {
  constexpr int D = 256 * NCH;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = i * 256 + lane * 8;
    if (xf) ld8f(xf + m * D + c, v[i]);
    else ld8b(xb + m * D + c, v[i]);
    if (pe) {
      float p[8];
      ld8f(pe + (long long)(m % T) * D + c, p);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] += p[j];
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ xf, const __nv_bfloat16* __restrict__ xb, const float* __restrict__ pe, int T,
              const float* __restrict__ gamma, const float* __restrict__ beta, float eps, long long M,
              __nv_bfloat16* __restrict__ out, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_trigger();
  pdl_wait();
  constexpr int D = 256 * NCH;
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  float v[NCH][8];
  ln_load_row<NCH>(xf, xb, pe, T, m, lane, v);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[i][j];
  const float mean = warp_sum(s) * (1.f / D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[i][j] - mean;
      ss = fmaf(d, d, ss);
    }
  const float rstd = rsqrtf(warp_sum(ss) * (1.f / D) + eps);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = i * 256 + lane * 8;
    float g[8], b[8], o[8];
    ld8f(gamma + c, g);
    ld8f(beta + c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf((v[i][j] - mean) * rstd, g[j], b[j]);
    st8b(out + m * D + c, o);
  }
  if (lane == 0) {
    mean_out[m] = mean;
    rstd_out[m] = rstd;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.  Also: dxm = dropout-masked copy of dx (the
// gradient entering the preceding Linear), dgamma/dbeta and the Linear's bias gradient dbias = colsum(dxm).
template <int NCH>
__global__ void __launch_bounds__(256, 2)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ xf, const __nv_bfloat16* __restrict__ xb,
              const float* __restrict__ pe, int T, const float* __restrict__ gamma, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, long long M, int rows_per_cta, __nv_bfloat16* __restrict__ dx,
              __nv_bfloat16* __restrict__ dxm, unsigned drop_thresh, float drop_scale, unsigned long long seed,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias) {
  pdl_trigger();
  pdl_wait();
  seed = pe_salted(seed);
  constexpr int D = 256 * NCH;
  // column partial sums (dgamma, dbeta, dbias) live in a private shared-memory slice per warp instead of 48 registers
  // per thread: that keeps two CTAs (16 rows in flight) resident per SM
  extern __shared__ __align__(16) float sacc[];  // [8 warps][3][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* my = sacc + warp * 3 * D;
  for (int i = lane * 4; i < 3 * D; i += 128) *reinterpret_cast<float4*>(my + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  float gm[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) ld8f(gamma + i * 256 + lane * 8, gm[i]);
  const long long m0 = (long long)blockIdx.x * rows_per_cta;
  const long long m1 = min(M, m0 + rows_per_cta);
  for (long long m = m0 + warp; m < m1; m += 8) {
    float v[NCH][8], d[NCH][8];
    ln_load_row<NCH>(xf, xb, pe, T, m, lane, v);
#pragma unroll
    for (int i = 0; i < NCH; ++i) ld8b(dy + m * D + i * 256 + lane * 8, d[i]);
    const float mean = mean_in[m], rstd = rstd_in[m];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      float* pg = my + i * 256 + lane * 8;
      float4 g0 = *reinterpret_cast<float4*>(pg), g1 = *reinterpret_cast<float4*>(pg + 4);
      float4 b0 = *reinterpret_cast<float4*>(pg + D), b1 = *reinterpret_cast<float4*>(pg + D + 4);
      float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      float ba[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] = (v[i][j] - mean) * rstd;  // xhat
        const float g = d[i][j] * gm[i][j];
        s1 += g;
        s2 = fmaf(g, v[i][j], s2);
        ga[j] = fmaf(d[i][j], v[i][j], ga[j]);
        ba[j] += d[i][j];
      }
      *reinterpret_cast<float4*>(pg) = make_float4(ga[0], ga[1], ga[2], ga[3]);
      *reinterpret_cast<float4*>(pg + 4) = make_float4(ga[4], ga[5], ga[6], ga[7]);
      *reinterpret_cast<float4*>(pg + D) = make_float4(ba[0], ba[1], ba[2], ba[3]);
      *reinterpret_cast<float4*>(pg + D + 4) = make_float4(ba[4], ba[5], ba[6], ba[7]);
    }
    s1 = warp_sum(s1) * (1.f / D);
    s2 = warp_sum(s2) * (1.f / D);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = i * 256 + lane * 8;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (d[i][j] * gm[i][j] - s1 - v[i][j] * s2);
      st8b(dx + m * D + c, o);
      if (dxm) {
        if (drop_thresh) {
          const unsigned long long e0 = (unsigned long long)(m * D + c);
          const uint32_t km = dropout_keep8(seed, e0 >> 3, drop_thresh);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = ((km >> j) & 1u) ? o[j] * drop_scale : 0.f;
        }
        st8b(dxm + m * D + c, o);
      }
      if (dbias) {
        float* pb = my + 2 * D + c;
        float4 a0 = *reinterpret_cast<float4*>(pb), a1 = *reinterpret_cast<float4*>(pb + 4);
        a0.x += __bfloat162float(__float2bfloat16(o[0])); a0.y += __bfloat162float(__float2bfloat16(o[1]));
        a0.z += __bfloat162float(__float2bfloat16(o[2])); a0.w += __bfloat162float(__float2bfloat16(o[3]));
        a1.x += __bfloat162float(__float2bfloat16(o[4])); a1.y += __bfloat162float(__float2bfloat16(o[5]));
        a1.z += __bfloat162float(__float2bfloat16(o[6])); a1.w += __bfloat162float(__float2bfloat16(o[7]));
        *reinterpret_cast<float4*>(pb) = a0;
        *reinterpret_cast<float4*>(pb + 4) = a1;
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * D; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sacc[w * 3 * D + c];
    const int which = c / D, col = c - which * D;
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
    if (dst) atomicAdd(dst + col, t);
  }
}

// out[n] += sum_m x[m][n]  (bias gradients of Linear layers).  CTA = 64 column groups (512 columns) x 4 row lanes.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long M, int N, long long ld, int rows_per_cta,
              float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[4][512];
  const int cg = N >> 3;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int gcol = blockIdx.y * 64 + tx;
  const long long m0 = (long long)blockIdx.x * rows_per_cta;
  const long long m1 = min(M, m0 + rows_per_cta);
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (gcol < cg) {
    long long m = m0 + ty;
    for (; m + 12 < m1; m += 16) {  // 4 independent 16-byte loads in flight
      float v0[8], v1[8], v2[8], v3[8];
      ld8b(x + m * ld + gcol * 8, v0);
      ld8b(x + (m + 4) * ld + gcol * 8, v1);
      ld8b(x + (m + 8) * ld + gcol * 8, v2);
      ld8b(x + (m + 12) * ld + gcol * 8, v3);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
    }
    for (; m < m1; m += 4) {
      float v[8];
      ld8b(x + m * ld + gcol * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = s[j];
  __syncthreads();
  for (int c = threadIdx.x; c < 512; c += 256) {
    const int col = blockIdx.y * 512 + c;
    if (col < N) atomicAdd(out + col, red[0][c] + red[1][c] + red[2][c] + red[3][c]);
  }
}

// ---------------------------------------------------------------------------------------------
// Multi-head softmax attention, head_dim 64, one CTA per (head, batch item), one thread per query row.
// qkv: [B*T][3*D] bf16 (q | k | v, head h at columns h*64); ctx: [B*T][D] bf16; lse: [B][H][T] fp32.
// Dropout (on the softmax probabilities, as nn.MultiheadAttention does) is indexed by ((b*H+h)*T+q)*T+j.
// ---------------------------------------------------------------------------------------------
constexpr int HD = 64;

__device__ __forceinline__ float dot64(const float* a, const float* smem_row) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 k = *reinterpret_cast<const float4*>(smem_row + d);
    s = fmaf(a[d], k.x, s);
    s = fmaf(a[d + 1], k.y, s);
    s = fmaf(a[d + 2], k.z, s);
    s = fmaf(a[d + 3], k.w, s);
  }
  return s;
}
__device__ __forceinline__ void axpy64(float* acc, float a, const float* smem_row) {
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 v = *reinterpret_cast<const float4*>(smem_row + d);
    acc[d] = fmaf(a, v.x, acc[d]);
    acc[d + 1] = fmaf(a, v.y, acc[d + 1]);
    acc[d + 2] = fmaf(a, v.z, acc[d + 2]);
    acc[d + 3] = fmaf(a, v.w, acc[d + 3]);
  }
}
// cooperative load of a [T][64] bf16 slice (row stride ld) into fp32 smem, scaled
__device__ __forceinline__ void load_head(const __nv_bfloat16* src, long long ld, int T, float scale, float* dst) {
  for (int i = threadIdx.x; i < T * (HD / 8); i += blockDim.x) {
    const int r = i / (HD / 8), c = (i % (HD / 8)) * 8;
    float v[8];
    ld8b(src + (long long)r * ld + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[r * HD + c + j] = v[j] * scale;
  }
}

__global__ void __launch_bounds__(256)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int T, int H, float scale, unsigned drop_thresh,
                float drop_scale, unsigned long long seed, __nv_bfloat16* __restrict__ ctx, float* __restrict__ lse) {
  pdl_trigger();
  pdl_wait();
  seed = pe_salted(seed);
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;
  float* Vs = sm + T * HD;
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * HD;
  const long long ld = 3LL * D;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld + h * HD;
  load_head(base + D, ld, T, 1.f, Ks);
  load_head(base + 2 * D, ld, T, 1.f, Vs);
  __syncthreads();
  const int q = threadIdx.x;
  if (q >= T) return;
  float qr[HD], o[HD];
#pragma unroll
  for (int c = 0; c < HD; c += 8) {
    float v[8];
    ld8b(base + (long long)q * ld + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      qr[c + j] = v[j] * scale;
      o[c + j] = 0.f;
    }
  }
  float mx = -INFINITY, l = 0.f;
  const uint32_t rkey = attn_row_key(seed, (unsigned long long)(b * H + h) * T + q);
  for (int j0 = 0; j0 < T; j0 += 8) {
    float s[8];
    float bm = mx;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      s[jj] = (j0 + jj < T) ? dot64(qr, Ks + (j0 + jj) * HD) : -INFINITY;
      bm = fmaxf(bm, s[jj]);
    }
    const float corr = __expf(mx - bm);
    l *= corr;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] *= corr;
    mx = bm;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      if (j0 + jj >= T) continue;
      const float p = __expf(s[jj] - mx);
      l += p;
      float pd = p;
      if (drop_thresh) pd = attn_drop_hash(rkey, (uint32_t)(j0 + jj)) < drop_thresh ? p * drop_scale : 0.f;
      axpy64(o, pd, Vs + (j0 + jj) * HD);
    }
  }
  const float inv = 1.f / l;
  __nv_bfloat16* op = ctx + ((long long)b * T + q) * D + h * HD;
#pragma unroll
  for (int c = 0; c < HD; c += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = o[c + j] * inv;
    st8b(op + c, v);
  }
  lse[((long long)b * H + h) * T + q] = mx + __logf(l);
}

// dQ (thread per query) and delta = rowsum(dO * O)
__global__ void __launch_bounds__(256)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ ctx,
                   const __nv_bfloat16* __restrict__ dctx, const float* __restrict__ lse, int T, int H, float scale,
                   unsigned drop_thresh, float drop_scale, unsigned long long seed, __nv_bfloat16* __restrict__ dqkv,
                   float* __restrict__ delta) {
  pdl_trigger();
  pdl_wait();
  seed = pe_salted(seed);
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;
  float* Vs = sm + T * HD;
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * HD;
  const long long ld = 3LL * D;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld + h * HD;
  load_head(base + D, ld, T, 1.f, Ks);
  load_head(base + 2 * D, ld, T, 1.f, Vs);
  __syncthreads();
  const int q = threadIdx.x;
  if (q >= T) return;
  float qr[HD], dor[HD], dq[HD];
  float dl = 0.f;
  const long long row = (long long)b * T + q;
#pragma unroll
  for (int c = 0; c < HD; c += 8) {
    float v[8], d[8], o[8];
    ld8b(base + (long long)q * ld + c, v);
    ld8b(dctx + row * D + h * HD + c, d);
    ld8b(ctx + row * D + h * HD + c, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      qr[c + j] = v[j] * scale;
      dor[c + j] = d[j];
      dq[c + j] = 0.f;
      dl = fmaf(d[j], o[j], dl);
    }
  }
  const float L = lse[((long long)b * H + h) * T + q];
  const uint32_t rkey = attn_row_key(seed, (unsigned long long)(b * H + h) * T + q);
  for (int j = 0; j < T; ++j) {
    const float p = __expf(dot64(qr, Ks + j * HD) - L);
    float dp = dot64(dor, Vs + j * HD);
    if (drop_thresh) dp = attn_drop_hash(rkey, (uint32_t)j) < drop_thresh ? dp * drop_scale : 0.f;
    axpy64(dq, p * (dp - dl), Ks + j * HD);
  }
  __nv_bfloat16* op = dqkv + row * ld + h * HD;
#pragma unroll
  for (int c = 0; c < HD; c += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = dq[c + j] * scale;
    st8b(op + c, v);
  }
  delta[((long long)b * H + h) * T + q] = dl;
}

// dK, dV (thread per key)
__global__ void __launch_bounds__(256, 1)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dctx,
                    const float* __restrict__ lse, const float* __restrict__ delta, int T, int H, float scale,
                    unsigned drop_thresh, float drop_scale, unsigned long long seed,
                    __nv_bfloat16* __restrict__ dqkv) {
  pdl_trigger();
  pdl_wait();
  seed = pe_salted(seed);
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                 // pre-scaled queries
  float* dOs = sm + T * HD;
  float* Ls = sm + 2 * T * HD;
  float* Dl = Ls + T;
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * HD;
  const long long ld = 3LL * D;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld + h * HD;
  load_head(base, ld, T, scale, Qs);
  load_head(dctx + (long long)b * T * D + h * HD, D, T, 1.f, dOs);
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    Ls[i] = lse[((long long)b * H + h) * T + i];
    Dl[i] = delta[((long long)b * H + h) * T + i];
  }
  __syncthreads();
  const int j = threadIdx.x;
  if (j >= T) return;
  float kr[HD], acc[HD];
#pragma unroll
  for (int c = 0; c < HD; c += 8) {
    float v[8];
    ld8b(base + D + (long long)j * ld + c, v);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      kr[c + t] = v[t];
      acc[c + t] = 0.f;
    }
  }
  const unsigned long long row_bh = (unsigned long long)(b * H + h) * T;
  // pass A: dV_j = sum_q dropout(P)_qj dO_q
  for (int q = 0; q < T; ++q) {
    float p = __expf(dot64(kr, Qs + q * HD) - Ls[q]);
    if (drop_thresh) p = attn_drop_hash(attn_row_key(seed, row_bh + q), (uint32_t)j) < drop_thresh ? p * drop_scale : 0.f;
    axpy64(acc, p, dOs + q * HD);
  }
  __nv_bfloat16* ov = dqkv + ((long long)b * T + j) * ld + 2 * D + h * HD;
#pragma unroll
  for (int c = 0; c < HD; c += 8) st8b(ov + c, acc + c);
  // pass B: dK_j = sum_q dS_qj Q_q   (Qs already carries the 1/sqrt(d) factor)
  float vr[HD];
#pragma unroll
  for (int c = 0; c < HD; c += 8) {
    float v[8];
    ld8b(base + 2 * D + (long long)j * ld + c, v);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      vr[c + t] = v[t];
      acc[c + t] = 0.f;
    }
  }
  for (int q = 0; q < T; ++q) {
    const float p = __expf(dot64(kr, Qs + q * HD) - Ls[q]);
    float dp = dot64(vr, dOs + q * HD);
    if (drop_thresh) dp = attn_drop_hash(attn_row_key(seed, row_bh + q), (uint32_t)j) < drop_thresh ? dp * drop_scale : 0.f;
    axpy64(acc, p * (dp - Dl[q]), Qs + q * HD);
  }
  __nv_bfloat16* ok = dqkv + ((long long)b * T + j) * ld + D + h * HD;
#pragma unroll
  for (int c = 0; c < HD; c += 8) st8b(ok + c, acc + c);
}

// ---------------------------------------------------------------------------------------------
// Heads + losses (num_class == 1): f0 = hc . wc + bc;  logit = hd . (wd0 + wd1) + bd0 + bd1 (model.py:96-98,115-117)
// loss_f0 = lambda * mean SmoothL1(beta=1)(f0, target), loss_sil = mean BCEWithLogits(logit, sil) (trainer.py:237-239)
// One warp per frame; optional gradients w.r.t. hc / hd (bf16) and the head parameters (atomic fp32).
// ---------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256)
heads_loss_kernel(const __nv_bfloat16* __restrict__ hc, const __nv_bfloat16* __restrict__ hd, long long M,
                  int rows_per_cta, const float* __restrict__ wc, const float* __restrict__ bc,
                  const float* __restrict__ wd, const float* __restrict__ bd, const float* __restrict__ f0_t,
                  const float* __restrict__ sil_t, float lambda_f0, float inv_count, float grad_scale,
                  float* __restrict__ f0_pred, float* __restrict__ sil_logit, double* __restrict__ loss_acc,
                  const float* __restrict__ gc_ext, const float* __restrict__ gd_ext,
                  __nv_bfloat16* __restrict__ dhc, __nv_bfloat16* __restrict__ dhd, float* __restrict__ dwc,
                  float* __restrict__ dbc, float* __restrict__ dwd, float* __restrict__ dbd) {
  pdl_trigger();
  pdl_wait();
  constexpr int D = 256 * NCH;
  __shared__ float red[2][D];
  __shared__ float redb[4];
  const bool want_grad = dhc != nullptr;
  for (int i = threadIdx.x; i < 2 * D; i += 256) (&red[0][0])[i] = 0.f;
  if (threadIdx.x < 4) redb[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float wcr[NCH][8], wdr[NCH][8], awc[NCH][8], awd[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = i * 256 + lane * 8;
    float w0[8], w1[8];
    ld8f(wc + c, wcr[i]);
    ld8f(wd + c, w0);
    ld8f(wd + D + c, w1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      wdr[i][j] = w0[j] + w1[j];
      awc[i][j] = awd[i][j] = 0.f;
    }
  }
  const float bias_c = bc[0], bias_d = bd[0] + bd[1];
  float l1 = 0.f, l2 = 0.f, gbc = 0.f, gbd = 0.f;
  const long long m0 = (long long)blockIdx.x * rows_per_cta;
  const long long m1 = min(M, m0 + rows_per_cta);
  for (long long m = m0 + warp; m < m1; m += 8) {
    float xc[NCH][8], xd[NCH][8];
    float zc = 0.f, zd = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = i * 256 + lane * 8;
      ld8b(hc + m * D + c, xc[i]);
      ld8b(hd + m * D + c, xd[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        zc = fmaf(xc[i][j], wcr[i][j], zc);
        zd = fmaf(xd[i][j], wdr[i][j], zd);
      }
    }
    zc = warp_sum(zc) + bias_c;
    zd = warp_sum(zd) + bias_d;
    const float y = f0_t ? f0_t[m] : 0.f, s = sil_t ? sil_t[m] : 0.f;
    const float d = zc - y, ad = fabsf(d);
    const float lf = ad < 1.f ? 0.5f * d * d : ad - 0.5f;
    const float gf = ad < 1.f ? d : (d > 0.f ? 1.f : -1.f);
    const float lb = fmaxf(zd, 0.f) - zd * s + log1pf(__expf(-fabsf(zd)));
    const float gb = 1.f / (1.f + __expf(-zd)) - s;
    if (lane == 0) {
      l1 += lf;
      l2 += lb;
      if (f0_pred) f0_pred[m] = zc;
      if (sil_logit) sil_logit[m] = zd;
    }
    if (want_grad) {
      const float gc = gc_ext ? gc_ext[m] : grad_scale * lambda_f0 * gf * inv_count;
      const float gd = gd_ext ? gd_ext[m] : grad_scale * gb * inv_count;
      if (lane == 0) {
        gbc += gc;
        gbd += gd;
      }
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = i * 256 + lane * 8;
        float oc[8], od[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          oc[j] = gc * wcr[i][j];
          od[j] = gd * wdr[i][j];
          awc[i][j] = fmaf(gc, xc[i][j], awc[i][j]);
          awd[i][j] = fmaf(gd, xd[i][j], awd[i][j]);
        }
        st8b(dhc + m * D + c, oc);
        st8b(dhd + m * D + c, od);
      }
    }
  }
  if (lane == 0) {
    atomicAdd(&redb[0], l1);
    atomicAdd(&redb[1], l2);
    atomicAdd(&redb[2], gbc);
    atomicAdd(&redb[3], gbd);
  }
  if (want_grad) {
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = i * 256 + lane * 8 + j;
        atomicAdd(&red[0][c], awc[i][j]);
        atomicAdd(&red[1][c], awd[i][j]);
      }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (loss_acc) {
      atomicAdd(loss_acc, (double)redb[0]);
      atomicAdd(loss_acc + 1, (double)redb[1]);
    }
    if (want_grad) {
      atomicAdd(dbc, redb[2]);
      atomicAdd(dbd, redb[3]);
      atomicAdd(dbd + 1, redb[3]);
    }
  }
  if (want_grad) {
    for (int c = threadIdx.x; c < D; c += 256) {
      atomicAdd(dwc + c, red[0][c]);
      atomicAdd(dwd + c, red[1][c]);
      atomicAdd(dwd + D + c, red[1][c]);
    }
  }
}

// loss_out[0] = lambda*sum1/count + sum2/count, [1] = lambda*sum1/count, [2] = sum2/count
__global__ void loss_finalize_kernel(const double* acc, float lambda_f0, double inv_count, float* loss_out) {
  pdl_trigger();
  pdl_wait();
  const double f = (double)lambda_f0 * acc[0] * inv_count, s = acc[1] * inv_count;
  loss_out[0] = (float)(f + s);
  loss_out[1] = (float)f;
  loss_out[2] = (float)s;
}

}  // namespace pe

// =================================================================================================
// C-ABI
// =================================================================================================
using namespace pe;
#define PE_ST(s) reinterpret_cast<cudaStream_t>(s)
#define PE_LAUNCH_RC() (cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH)

extern "C" int pe_layernorm_fwd(const float* x_f32, const void* x_bf16, const float* pe_table, int T, int D,
                                const float* gamma, const float* beta, float eps, long long M, void* out, float* mean,
                                float* rstd, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if ((!x_f32 && !x_bf16) || !gamma || !beta || !out || !mean || !rstd || M <= 0 || (pe_table && T <= 0))
    return PE_ERR_BAD_SHAPE;
  const unsigned grid = (unsigned)((M + 7) / 8);
  if (D == 512)
    pe_host::launch(ln_fwd_kernel<2>, dim3(grid), dim3(256), 0, PE_ST(stream), x_f32, (const __nv_bfloat16*)x_bf16, pe_table, T, gamma, beta,
                                                      eps, M, (__nv_bfloat16*)out, mean, rstd);
  else if (D == 768)
    pe_host::launch(ln_fwd_kernel<3>, dim3(grid), dim3(256), 0, PE_ST(stream), x_f32, (const __nv_bfloat16*)x_bf16, pe_table, T, gamma, beta,
                                                      eps, M, (__nv_bfloat16*)out, mean, rstd);
  else if (D == 256)
    pe_host::launch(ln_fwd_kernel<1>, dim3(grid), dim3(256), 0, PE_ST(stream), x_f32, (const __nv_bfloat16*)x_bf16, pe_table, T, gamma, beta,
                                                      eps, M, (__nv_bfloat16*)out, mean, rstd);
  else
    return PE_ERR_BAD_SHAPE;
  return PE_LAUNCH_RC();
}

extern "C" int pe_layernorm_bwd(const void* dy, const float* x_f32, const void* x_bf16, const float* pe_table, int T,
                                int D, const float* gamma, const float* mean, const float* rstd, long long M, void* dx,
                                void* dx_masked, unsigned drop_thresh, float drop_scale, unsigned long long seed,
                                float* dgamma, float* dbeta, float* dbias, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!dy || (!x_f32 && !x_bf16) || !gamma || !mean || !rstd || !dx || M <= 0 || (pe_table && T <= 0))
    return PE_ERR_BAD_SHAPE;
  // two CTAs per SM, one wave
  long long per_ll = (M + 2LL * pe_host::num_sms() - 1) / (2LL * pe_host::num_sms());
  if (per_ll < 8) per_ll = 8;
  const int per = (int)per_ll;
  const unsigned grid = (unsigned)((M + per - 1) / per);
  const size_t smem = (size_t)8 * 3 * D * sizeof(float);
#define PE_LN_BWD(N)                                                                                               \
  do {                                                                                                             \
    static bool attr = false;                                                                                      \
    if (!attr) {                                                                                                   \
      cudaFuncSetAttribute(ln_bwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 768 * 4);        \
      attr = true;                                                                                                 \
    }                                                                                                              \
    pe_host::launch(ln_bwd_kernel<N>, dim3(grid), dim3(256), smem, PE_ST(stream), (const __nv_bfloat16*)dy, x_f32, (const __nv_bfloat16*)x_bf16, \
                                                         pe_table, T, gamma, mean, rstd, M, per, (__nv_bfloat16*)dx, \
                                                         (__nv_bfloat16*)dx_masked, drop_thresh, drop_scale, seed, \
                                                         dgamma, dbeta, dbias);                                    \
  } while (0)
  if (D == 512) PE_LN_BWD(2);
  else if (D == 768) PE_LN_BWD(3);
  else if (D == 256) PE_LN_BWD(1);
  else return PE_ERR_BAD_SHAPE;
#undef PE_LN_BWD
  return PE_LAUNCH_RC();
}

extern "C" int pe_colsum_bf16(const void* x, long long M, int N, long long ld, float* out, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!x || !out || M <= 0 || N <= 0 || (N % 8) || (ld % 8)) return PE_ERR_BAD_SHAPE;
  // ~6 CTAs per SM in flight (each keeps 16 KB of loads outstanding), at least 32 rows each
  const unsigned gy = (unsigned)((N / 8 + 63) / 64);
  long long per_ll = (M * gy + 6LL * pe_host::num_sms() - 1) / (6LL * pe_host::num_sms());
  per_ll = ((per_ll + 15) / 16) * 16;
  if (per_ll < 32) per_ll = 32;
  const int per = (int)per_ll;
  dim3 grid((unsigned)((M + per - 1) / per), gy);
  pe_host::launch(colsum_kernel, dim3(grid), dim3(256), 0, PE_ST(stream), (const __nv_bfloat16*)x, M, N, ld, per, out);
  return PE_LAUNCH_RC();
}

static int attn_threads(int T) { return ((T + 31) / 32) * 32; }

// tcgen05 kernels for the T = 192 segment length (attn_tc.cu); other lengths use the SIMT kernels above
int pe_attn_fwd_tc(const void* qkv, int B, int H, unsigned drop_thresh, float drop_scale, unsigned long long seed,
                   void* ctx, float* lse, cudaStream_t stream);
int pe_attn_bwd_tc(const void* qkv, const void* ctx, const void* dctx, const float* lse, int B, int H,
                   unsigned drop_thresh, float drop_scale, unsigned long long seed, void* dqkv, float* delta,
                   cudaStream_t stream);
extern "C" int pe_attn_fwd(const void* qkv, int B, int T, int H, int head_dim, unsigned drop_thresh, float drop_scale,
                           unsigned long long seed, void* ctx, float* lse, int force_simt, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!qkv || !ctx || !lse || B <= 0 || T <= 0 || T > 256 || H <= 0 || head_dim != HD) return PE_ERR_BAD_SHAPE;
  if (T == 192 && !force_simt)
    return pe_attn_fwd_tc(qkv, B, H, drop_thresh, drop_scale, seed, ctx, lse, PE_ST(stream));
  const size_t smem = 2ull * T * HD * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  pe_host::launch(attn_fwd_kernel, dim3(dim3(H, B)), dim3(attn_threads(T)), smem, PE_ST(stream), (const __nv_bfloat16*)qkv, T, H, 0.125f,
                                                                        drop_thresh, drop_scale, seed,
                                                                        (__nv_bfloat16*)ctx, lse);
  return PE_LAUNCH_RC();
}

extern "C" int pe_attn_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse, int B, int T, int H,
                           int head_dim, unsigned drop_thresh, float drop_scale, unsigned long long seed, void* dqkv,
                           float* delta, int force_simt, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!qkv || !ctx || !dctx || !lse || !dqkv || !delta || B <= 0 || T <= 0 || T > 256 || H <= 0 || head_dim != HD)
    return PE_ERR_BAD_SHAPE;
  if (T == 192 && !force_simt)
    return pe_attn_bwd_tc(qkv, ctx, dctx, lse, B, H, drop_thresh, drop_scale, seed, dqkv, delta, PE_ST(stream));
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  const size_t smem1 = 2ull * T * HD * sizeof(float);
  pe_host::launch(attn_bwd_dq_kernel, dim3(dim3(H, B)), dim3(attn_threads(T)), smem1, PE_ST(stream), 
      (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)ctx, (const __nv_bfloat16*)dctx, lse, T, H, 0.125f, drop_thresh,
      drop_scale, seed, (__nv_bfloat16*)dqkv, delta);
  const size_t smem2 = (2ull * T * HD + 2ull * T) * sizeof(float);
  pe_host::launch(attn_bwd_dkv_kernel, dim3(dim3(H, B)), dim3(attn_threads(T)), smem2, PE_ST(stream), 
      (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dctx, lse, delta, T, H, 0.125f, drop_thresh, drop_scale, seed,
      (__nv_bfloat16*)dqkv);
  return PE_LAUNCH_RC();
}

extern "C" int pe_heads_loss(const void* hc, const void* hd, long long M, int D, const float* wc, const float* bc,
                             const float* wd, const float* bd, const float* f0_target, const float* sil_target,
                             float lambda_f0, float grad_scale, float* f0_pred, float* sil_logit,
                             double* loss_acc /* [2], zeroed */, float* loss_out /* [3] */, const float* gc_ext,
                             const float* gd_ext, void* dhc, void* dhd,
                             float* dwc, float* dbc, float* dwd, float* dbd, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!hc || !hd || !wc || !bc || !wd || !bd || M <= 0) return PE_ERR_BAD_SHAPE;
  const bool ext = gc_ext && gd_ext;
  if (!ext && (!f0_target || !sil_target || !loss_acc || !loss_out)) return PE_ERR_BAD_SHAPE;
  if (dhc && (!dhd || !dwc || !dbc || !dwd || !dbd)) return PE_ERR_BAD_SHAPE;
  const int per = 64;
  const unsigned grid = (unsigned)((M + per - 1) / per);
  const float inv = 1.0f / (float)M;
#define PE_HEADS(N)                                                                                                  \
  pe_host::launch(heads_loss_kernel<N>, dim3(grid), dim3(256), 0, PE_ST(stream), (const __nv_bfloat16*)hc, (const __nv_bfloat16*)hd, M, per, \
                                                        wc, bc, wd, bd, f0_target, sil_target, lambda_f0, inv,      \
                                                        grad_scale, f0_pred, sil_logit, loss_acc, gc_ext, gd_ext,   \
                                                        (__nv_bfloat16*)dhc, (__nv_bfloat16*)dhd, dwc, dbc, dwd, dbd)
  if (D == 512) PE_HEADS(2);
  else if (D == 768) PE_HEADS(3);
  else return PE_ERR_BAD_SHAPE;
#undef PE_HEADS
  if (loss_acc && loss_out) pe_host::launch(loss_finalize_kernel, dim3(1), dim3(1), 0, PE_ST(stream), loss_acc, lambda_f0, 1.0 / (double)M, loss_out);
  return PE_LAUNCH_RC();
}
