// Memory-bound passes of the JDCNet conv trunk over NHWC bf16 activations (reference model.py:23-57,143-175):
// stem convolution (Cin = 1), BatchNorm batch statistics / finalize, fused BN-apply + LeakyReLU + MaxPool(1,k)
// (+ Dropout) forward, the matching two-pass BatchNorm backward, and the auxiliary max-pools of the detector
// branch.  All kernels are coalesced 16-byte-vector passes; per-channel reductions use fp32 partials per CTA and
// fp64 atomics across CTAs.
#include "common.cuh"
#include <cstdlib>
#include "../../include/pitchextractor_b200.h"

PE_USES_STEP_SALT()

namespace pe {

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(h[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* f) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ float bf16_at(const uint4& q, int i) {
  const uint32_t w = i < 2 ? q.x : i < 4 ? q.y : i < 6 ? q.z : q.w;
  return (i & 1) ? bf16_hi(w) : bf16_lo(w);
}
// streaming 16-byte load: read once, do not keep in L1
__device__ __forceinline__ uint4 ld_stream(const __nv_bfloat16* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// ---------------------------------------------------------------------------------------------
// stem: Conv2d(1 -> 64, 3x3, pad 1, no bias) on x[b][t][f] (arbitrary strides) -> y[b][t][f][64] bf16
// ---------------------------------------------------------------------------------------------
// the 3x3 input neighbourhood of pixel (b, t, f), zero padded; branch-free (clamped addresses, selected values)
__device__ __forceinline__ void stem_taps(const float* __restrict__ x, long long sb, long long st, long long sf, int b,
                                          int t, int f, int T, int F, float* xv) {
  const float* xb = x + b * sb;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int tt = t + kh - 1;
    const bool tok = tt >= 0 && tt < T;
    const float* xr = xb + (long long)min(max(tt, 0), T - 1) * st;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ff = f + kw - 1;
      const float v = __ldg(xr + (long long)min(max(ff, 0), F - 1) * sf);
      xv[kh * 3 + kw] = (tok && ff >= 0 && ff < F) ? v : 0.f;
    }
  }
}
// pixel index -> (b, t, f) once, then stepped by 32 pixels per iteration without divisions
struct StemPos {
  int b, t, f;
  __device__ __forceinline__ StemPos(long long p, int T, int F) {
    f = (int)((unsigned)p % (unsigned)F);
    const unsigned bt = (unsigned)p / (unsigned)F;
    t = (int)(bt % (unsigned)T);
    b = (int)(bt / (unsigned)T);
  }
  __device__ __forceinline__ void advance(int n, int T, int F) {
    f += n;
    while (f >= F) {
      f -= F;
      if (++t == T) {
        t = 0;
        ++b;
      }
    }
  }
};

// A thread owns one 8-channel group (its 72 weights stay in registers) and walks pixels 32 apart; the BatchNorm
// batch statistics of the bf16-rounded output (model.py:24) are accumulated on the way when `stats` is given.
__global__ void __launch_bounds__(256)
stem_conv_fwd_kernel(const float* __restrict__ x, long long sb, long long st, long long sf, int B, int T, int F,
                     const float* __restrict__ w /*[64][9]*/, __nv_bfloat16* __restrict__ y, int pixels_per_cta,
                     double* __restrict__ stats /* [2][64] or NULL */) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[2][64];
  if (threadIdx.x < 128) (&red[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  const int g = threadIdx.x & 7;
  const int ty = threadIdx.x >> 3;  // 32 pixel lanes
  float wr[8][9];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[c][k] = __ldg(w + (g * 8 + c) * 9 + k);
  float s1[8], s2[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) s1[c] = s2[c] = 0.f;
  const long long P = (long long)B * T * F;
  const long long p0 = (long long)blockIdx.x * pixels_per_cta;
  const long long p1 = min(P, p0 + pixels_per_cta);
  StemPos pos(p0 + ty, T, F);
  for (long long p = p0 + ty; p < p1; p += 32, pos.advance(32, T, F)) {
    float xv[9];
    stem_taps(x, sb, st, sf, pos.b, pos.t, pos.f, T, F, xv);
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) a = fmaf(xv[k], wr[c][k], a);
      o[c] = a;
      const float r = __bfloat162float(__float2bfloat16(a));
      s1[c] += r;
      s2[c] = fmaf(r, r, s2[c]);
    }
    st8(y + p * 64 + g * 8, o);
  }
  if (stats) {
    // lanes with equal g inside a warp: lane = g + 8*j
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a = s1[c], q = s2[c];
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      q += __shfl_xor_sync(0xffffffffu, q, 8);
      q += __shfl_xor_sync(0xffffffffu, q, 16);
      if ((threadIdx.x & 31) < 8) {
        atomicAdd(&red[0][g * 8 + c], a);
        atomicAdd(&red[1][g * 8 + c], q);
      }
    }
    __syncthreads();
    if (threadIdx.x < 128) atomicAdd(stats + threadIdx.x, (double)(&red[0][0])[threadIdx.x]);
  }
}

// dw[c][tap] += sum_p dy[p][c] * x[p + tap]; two pixels per thread and iteration in flight
__global__ void __launch_bounds__(256)
stem_conv_wgrad_kernel(const float* __restrict__ x, long long sb, long long st, long long sf, int B, int T, int F,
                       const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw, int pixels_per_cta) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[576];
  for (int i = threadIdx.x; i < 576; i += 256) red[i] = 0.f;
  __syncthreads();
  const int g = threadIdx.x & 7;
  const int ty = threadIdx.x >> 3;  // 32 pixel lanes
  const long long P = (long long)B * T * F;
  const long long p0 = (long long)blockIdx.x * pixels_per_cta;
  const long long p1 = min(P, p0 + pixels_per_cta);
  float acc[8][9];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[c][k] = 0.f;
  StemPos pos(p0 + ty, T, F);
  for (long long pp = p0 + ty; pp < p1; pp += 64) {
    uint4 dq[2];
    float xv[2][9];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long p = pp + 32 * u;
      // (past the end: dy = 0 makes the contribution vanish; the clamped tap addresses stay inside the last image)
      dq[u] = p < p1 ? ld_stream(dy + p * 64 + g * 8) : make_uint4(0u, 0u, 0u, 0u);
      stem_taps(x, sb, st, sf, min(pos.b, B - 1), pos.t, pos.f, T, F, xv[u]);
      pos.advance(32, T, F);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float d = bf16_at(dq[u], c);
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[c][k] = fmaf(d, xv[u][k], acc[c][k]);
      }
  }
  // lanes with equal g inside a warp: lane = g + 8*j, j = 0..3
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float v = acc[c][k];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((threadIdx.x & 31) < 8) atomicAdd(&red[(g * 8 + c) * 9 + k], v);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 576; i += 256) atomicAdd(dw + i, red[i]);
}

// ---------------------------------------------------------------------------------------------
// BatchNorm2d (training) statistics over x[rows][C] bf16 -> sums[0][c] = sum x, sums[1][c] = sum x^2 (fp64 atomics)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C, double* __restrict__ sums,
                int rows_per_cta) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sred[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  const int cg = C >> 3;
  const int ry = blockDim.x / cg;
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
  if (ty < ry) {
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = min(rows, r0 + rows_per_cta);
    long long r = r0 + ty;
    for (; r + 3LL * ry < r1; r += 4LL * ry) {  // four independent 16-byte loads in flight per thread
      float v0[8], v1[8], v2[8], v3[8];
      ld8(x + r * C + tx * 8, v0);
      ld8(x + (r + ry) * C + tx * 8, v1);
      ld8(x + (r + 2LL * ry) * C + tx * 8, v2);
      ld8(x + (r + 3LL * ry) * C + tx * 8, v3);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += (v0[i] + v1[i]) + (v2[i] + v3[i]);
        ss[i] += fmaf(v0[i], v0[i], v1[i] * v1[i]) + fmaf(v2[i], v2[i], v3[i] * v3[i]);
      }
    }
    for (; r < r1; r += ry) {
      float v[8];
      ld8(x + r * C + tx * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += v[i];
        ss[i] = fmaf(v[i], v[i], ss[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(&sred[tx * 8 + i], s[i]);
      atomicAdd(&sred[C + tx * 8 + i], ss[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(sums + i, (double)sred[i]);
}

// mean / biased var -> scale, shift (+ saved mean, rstd) and the running-stat update of nn.BatchNorm2d
// (momentum 0.1, unbiased running variance, num_batches_tracked += 1).
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* scale,
                                   float* shift, float* mean_out, float* rstd_out, float* running_mean,
                                   float* running_var, long long* nbt, int C) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const float sc = (float)((double)gamma[c] * rstd);
  scale[c] = sc;
  shift[c] = (float)((double)beta[c] - mean * (double)gamma[c] * rstd);
  mean_out[c] = (float)mean;
  rstd_out[c] = (float)rstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval mode: scale / shift from the running statistics
__global__ void bn_eval_params_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      float eps, float* scale, float* shift, float* mean_out, float* rstd_out, int C) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float rstd = rsqrtf(running_var[c] + eps);
  scale[c] = gamma[c] * rstd;
  shift[c] = beta[c] - running_mean[c] * gamma[c] * rstd;
  if (mean_out) mean_out[c] = running_mean[c];
  if (rstd_out) rstd_out[c] = rstd;
}

// ---------------------------------------------------------------------------------------------
// y = Dropout(MaxPool_(1,k)(LeakyReLU(x * scale + shift)));  scale == NULL => pure max-pool of x.
// x: [rows][W][C];  out: pixel (row, wo) at out + (row*Wo + wo) * ld_out + c_off;  optional second copy in the
// sequence-model layout out_seq[row][c*Wo + wo] (model.py:93,112 permute+view).
// ---------------------------------------------------------------------------------------------
struct PoolGeom {
  long long rows;
  int W, C, k, Wo;
};

__global__ void __launch_bounds__(256)
bn_act_pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, PoolGeom g, const float* __restrict__ scale,
                       const float* __restrict__ shift, float slope, unsigned drop_thresh, float drop_scale,
                       unsigned long long seed, __nv_bfloat16* __restrict__ out, long long ld_out, int c_off,
                       __nv_bfloat16* __restrict__ out_seq, unsigned char* __restrict__ argmax_out) {
  pdl_trigger();
  pdl_wait();
  seed = pe_salted(seed);
  const int cg = g.C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = g.rows * g.Wo * cg;
  if (idx >= total) return;
  const int tx = (int)(idx % cg);
  const int wo = (int)((idx / cg) % g.Wo);
  const long long row = idx / ((long long)cg * g.Wo);
  float sc[8], sh[8];
  if (scale) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sc[i] = __ldg(scale + tx * 8 + i);
      sh[i] = __ldg(shift + tx * 8 + i);
    }
  }
  float best[8];
  int jbest[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    best[i] = -INFINITY;
    jbest[i] = 0;
  }
  const __nv_bfloat16* xp = x + ((row * g.W + (long long)wo * g.k) * g.C + tx * 8);
  for (int j = 0; j < g.k; ++j) {
    float v[8];
    ld8(xp + (long long)j * g.C, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float z = v[i];
      if (scale) {
        z = fmaf(z, sc[i], sh[i]);
        z = z > 0.f ? z : z * slope;
      }
      if (z > best[i]) {  // first maximum wins, as in the backward routing
        best[i] = z;
        jbest[i] = j;
      }
    }
  }
  if (argmax_out) {
    const uint32_t lo = (uint32_t)jbest[0] | ((uint32_t)jbest[1] << 8) | ((uint32_t)jbest[2] << 16) | ((uint32_t)jbest[3] << 24);
    const uint32_t hi = (uint32_t)jbest[4] | ((uint32_t)jbest[5] << 8) | ((uint32_t)jbest[6] << 16) | ((uint32_t)jbest[7] << 24);
    *reinterpret_cast<uint2*>(argmax_out + (row * g.Wo + wo) * g.C + tx * 8) = make_uint2(lo, hi);
  }
  if (drop_thresh) {
    const unsigned long long e0 = (unsigned long long)((row * g.Wo + wo) * g.C + tx * 8);
    const uint32_t km = dropout_keep8(seed, e0 >> 3, drop_thresh);
#pragma unroll
    for (int i = 0; i < 8; ++i) best[i] = ((km >> i) & 1u) ? best[i] * drop_scale : 0.f;
  }
  if (out) st8(out + (row * g.Wo + wo) * ld_out + c_off + tx * 8, best);
  if (out_seq) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      out_seq[row * ((long long)g.C * g.Wo) + (long long)(tx * 8 + i) * g.Wo + wo] = __float2bfloat16(best[i]);
  }
}

// Same pass for K in {1, 2, 4} with memory-level parallelism: a thread owns one 8-channel group (its BN constants stay
// in registers) and U windows spaced one grid sweep apart; all U*K 16-byte loads are issued before any arithmetic, so
// a 256-thread CTA keeps 16 KB in flight instead of 4 KB.  blockDim.x = (C/8) * ry exactly.
template <int K, int U, bool EXTRA>
__global__ void __launch_bounds__(256)
bn_act_pool_fwd_mlp_kernel(const __nv_bfloat16* __restrict__ x, PoolGeom g, const float* __restrict__ scale,
                           const float* __restrict__ shift, float slope, unsigned drop_thresh, float drop_scale,
                           unsigned long long seed, __nv_bfloat16* __restrict__ out, long long ld_out, int c_off,
                           __nv_bfloat16* __restrict__ out_seq) {
  pdl_trigger();
  pdl_wait();
  seed = pe_salted(seed);
  const int cg = g.C >> 3;
  const int ry = blockDim.x / cg;
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  const long long nwin = g.rows * g.Wo;
  const long long stride = (long long)gridDim.x * ry;
  const long long w_first = (long long)blockIdx.x * ry + ty;
  float sc[8], sh[8];
  if (scale) {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(scale + tx * 8)), a1 = __ldg(reinterpret_cast<const float4*>(scale + tx * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + tx * 8)), b1 = __ldg(reinterpret_cast<const float4*>(shift + tx * 8 + 4));
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
  }
  uint4 xv[U][K];
  long long wrow[U];
  int wwo[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long w = w_first + u * stride;
    const long long row = (long long)((unsigned)w / (unsigned)g.Wo);  // the launcher guarantees nwin < 2^31
    wrow[u] = w < nwin ? row : -1;
    wwo[u] = (int)(w - row * g.Wo);
    if (w < nwin) {
      const __nv_bfloat16* xp = x + ((row * g.W + (long long)wwo[u] * K) * g.C + tx * 8);
#pragma unroll
      for (int j = 0; j < K; ++j) xv[u][j] = ld_stream(xp + (long long)j * g.C);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (wrow[u] < 0) continue;
    const long long row = wrow[u];
    const int wo = wwo[u];
    float best[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float b = -INFINITY;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        float z = bf16_at(xv[u][j], i);
        if (scale) {
          z = fmaf(z, sc[i], sh[i]);
          z = z > 0.f ? z : z * slope;
        }
        b = fmaxf(b, z);
      }
      best[i] = b;
    }
    if (EXTRA && drop_thresh) {
      const unsigned long long e0 = (unsigned long long)((row * g.Wo + wo) * g.C + tx * 8);
      const uint32_t km = dropout_keep8(seed, e0 >> 3, drop_thresh);
#pragma unroll
      for (int i = 0; i < 8; ++i) best[i] = ((km >> i) & 1u) ? best[i] * drop_scale : 0.f;
    }
    if (out) st8(out + (row * g.Wo + wo) * ld_out + c_off + tx * 8, best);
    if (EXTRA && out_seq) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        out_seq[row * ((long long)g.C * g.Wo) + (long long)(tx * 8 + i) * g.Wo + wo] = __float2bfloat16(best[i]);
    }
  }
}

// Backward of y = Dropout(MaxPool(1,K)(LeakyReLU(BN_train(x)))) in two coalesced passes.  For one window (row, wo) and 8
// channels a thread recomputes z_j = lrelu(x_j*scale + shift), the first arg-max j*, and g* = dout * dropout * lrelu'.
// Per-channel constants live in shared memory (registers are kept low so that 32+ warps per SM hide HBM latency).
struct BnBwdArgs {
  const __nv_bfloat16* x;
  PoolGeom g;
  const float* scale;
  const float* shift;
  float slope;
  unsigned drop_thresh;
  float drop_scale;
  unsigned long long seed;
  const __nv_bfloat16* dout;      // NHWC consumer gradient (may be NULL)
  long long ld_dout;
  int c_off;
  const __nv_bfloat16* dout_seq;  // sequence-layout consumer gradient (may be NULL)
  // auxiliary MaxPool(1, aux_k) of the same tensor (model.py:45-49): its gradient joins at the saved arg-max positions
  const unsigned char* aux_idx;   // [rows][W / aux_k][C] (may be NULL)
  const __nv_bfloat16* aux_dout;
  long long aux_ld;
  int aux_c_off, aux_k;
};


// loads the K-window and the consumer gradient, returns per channel: jstar, g*, x at jstar
template <int K, bool EXTRA = true>
__device__ __forceinline__ void bn_bwd_route(const BnBwdArgs& a, const float* s_sc, const float* s_sh, long long row,
                                             int wo, int tx, uint4* xv, int* jstar, float* gstar, float* xstar) {
  const PoolGeom& g = a.g;
  const __nv_bfloat16* xp = a.x + ((row * g.W + (long long)wo * K) * g.C + tx * 8);
#pragma unroll
  for (int j = 0; j < K; ++j) xv[j] = __ldg(reinterpret_cast<const uint4*>(xp + (long long)j * g.C));
  float go[8];
  if (a.dout) {
    const uint4 d = __ldg(reinterpret_cast<const uint4*>(a.dout + (row * g.Wo + wo) * a.ld_dout + a.c_off + tx * 8));
#pragma unroll
    for (int i = 0; i < 8; ++i) go[i] = bf16_at(d, i);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) go[i] = 0.f;
  }
  if (EXTRA && a.dout_seq) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      go[i] += __bfloat162float(a.dout_seq[row * ((long long)g.C * g.Wo) + (long long)(tx * 8 + i) * g.Wo + wo]);
  }
  if (EXTRA && a.drop_thresh) {
    const unsigned long long e0 = (unsigned long long)((row * g.Wo + wo) * g.C + tx * 8);
    const uint32_t km = dropout_keep8(a.seed, e0 >> 3, a.drop_thresh);
#pragma unroll
    for (int i = 0; i < 8; ++i) go[i] = ((km >> i) & 1u) ? go[i] * a.drop_scale : 0.f;
  }
  const float4 sc0 = *reinterpret_cast<const float4*>(s_sc + tx * 8), sc1 = *reinterpret_cast<const float4*>(s_sc + tx * 8 + 4);
  const float4 sh0 = *reinterpret_cast<const float4*>(s_sh + tx * 8), sh1 = *reinterpret_cast<const float4*>(s_sh + tx * 8 + 4);
  const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
  const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float best = -INFINITY, pre = 0.f, xs = 0.f;
    int js = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float xj = bf16_at(xv[j], i);
      const float zp = fmaf(xj, sc[i], sh[i]);
      const float z = zp > 0.f ? zp : zp * a.slope;
      if (z > best) {
        best = z;
        js = j;
        pre = zp;
        xs = xj;
      }
    }
    jstar[i] = js;
    xstar[i] = xs;
    gstar[i] = go[i] * (pre > 0.f ? 1.f : a.slope);
  }
}

// pass 1: sums[0][c] += sum g, sums[1][c] += sum g * x   (raw x; centred / scaled in fp64 by the params kernel)
template <int K, bool EXTRA>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(BnBwdArgs a, double* __restrict__ sums, int windows_per_cta) {
  pdl_trigger();
  pdl_wait();
  a.seed = pe_salted(a.seed);
  extern __shared__ __align__(16) float sm[];  // sc[C] | sh[C] | red[2][C]
  const PoolGeom& g = a.g;
  float* s_sc = sm;
  float* s_sh = sm + g.C;
  float* red = sm + 2 * g.C;
  for (int i = threadIdx.x; i < g.C; i += blockDim.x) {
    s_sc[i] = a.scale[i];
    s_sh[i] = a.shift[i];
    red[i] = 0.f;
    red[g.C + i] = 0.f;
  }
  __syncthreads();
  const int cg = g.C >> 3;
  const int ry = blockDim.x / cg;
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  if (ty < ry) {
    float s[8], sx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = sx[i] = 0.f;
    const long long nwin = g.rows * g.Wo;
    const long long w0 = (long long)blockIdx.x * windows_per_cta;
    const long long w1 = min(nwin, w0 + windows_per_cta);
#pragma unroll 2
    for (long long wi = w0 + ty; wi < w1; wi += ry) {
      const long long row = wi / g.Wo;
      const int wo = (int)(wi - row * g.Wo);
      uint4 xv[K];
      int js[8];
      float gs[8], xs[8];
      bn_bwd_route<K, EXTRA>(a, s_sc, s_sh, row, wo, tx, xv, js, gs, xs);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += gs[i];
        sx[i] = fmaf(gs[i], xs[i], sx[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(&red[tx * 8 + i], s[i]);
      atomicAdd(&red[g.C + tx * 8 + i], sx[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * g.C; i += blockDim.x) atomicAdd(sums + i, (double)red[i]);
}

// dbeta += S_g, dgamma += rstd * (S_gx - mean * S_g); coefficients of pass 2: dx = scale*g - A*x + B with
// A = scale*rstd*mgx, B = A*mean - scale*mg, mg = S_g / n, mgx = rstd * (S_gx - mean*S_g) / n
__global__ void bn_bwd_params_kernel(const double* __restrict__ sums, double count, const float* __restrict__ scale,
                                     const float* __restrict__ mean, const float* __restrict__ rstd, float* dgamma,
                                     float* dbeta, float* __restrict__ coef, int C) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double sg = sums[c], sgx = sums[C + c];
  const double mu = mean[c], rs = rstd[c], sc = scale[c];
  const double sgxh = rs * (sgx - mu * sg);
  if (dbeta) dbeta[c] += (float)sg;
  if (dgamma) dgamma[c] += (float)sgxh;
  const double mg = sg / count, mgx = sgxh / count;
  const double A = sc * rs * mgx;
  coef[c] = (float)A;
  coef[C + c] = (float)(A * mu - sc * mg);
}

// pass 2: dx for every input element (pooled-away elements have g = 0; the last window also owns the columns that
// the floor of W / K drops)
template <int K>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(BnBwdArgs a, const float* __restrict__ coef, __nv_bfloat16* __restrict__ dx) {
  pdl_trigger();
  pdl_wait();
  a.seed = pe_salted(a.seed);
  extern __shared__ __align__(16) float sm[];  // sc | sh | A | B
  const PoolGeom& g = a.g;
  float* s_sc = sm;
  float* s_sh = sm + g.C;
  float* s_A = sm + 2 * g.C;
  float* s_B = sm + 3 * g.C;
  for (int i = threadIdx.x; i < g.C; i += blockDim.x) {
    s_sc[i] = a.scale[i];
    s_sh[i] = a.shift[i];
    s_A[i] = coef[i];
    s_B[i] = coef[g.C + i];
  }
  __syncthreads();
  const int cg = g.C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = g.rows * g.Wo * cg;
  if (idx >= total) return;
  const int tx = (int)(idx % cg);
  const int wo = (int)((idx / cg) % g.Wo);
  const long long row = idx / ((long long)cg * g.Wo);
  uint4 xv[K];
  int js[8];
  float gs[8], xs[8];
  bn_bwd_route<K>(a, s_sc, s_sh, row, wo, tx, xv, js, gs, xs);
  const long long base = (row * g.W + (long long)wo * K) * g.C + tx * 8;
  float sg[8], A[8], Bc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sg[i] = s_sc[tx * 8 + i] * gs[i];
    A[i] = s_A[tx * 8 + i];
    Bc[i] = s_B[tx * 8 + i];
  }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(-A[i], bf16_at(xv[j], i), Bc[i]) + (j == js[i] ? sg[i] : 0.f);
    st8(dx + base + (long long)j * g.C, o);
  }
  if (wo == g.Wo - 1) {
    for (int j = K; j < g.W - wo * K; ++j) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(a.x + base + (long long)j * g.C));
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(-A[i], bf16_at(q, i), Bc[i]);
      st8(dx + base + (long long)j * g.C, o);
    }
  }
}

// pass 2 with memory-level parallelism (see bn_act_pool_fwd_mlp_kernel): U windows per thread, every 16-byte load of
// x and of the consumer gradient issued before the arithmetic; per-channel constants in registers.
// EXTRA = false compiles the dropout replay and the sequence-layout gradient gather out (most layers have neither;
// predicated-off Philox rounds would otherwise still take issue slots and make the pass instruction-bound).
template <int K, int U, bool EXTRA, bool AUX>
__global__ void __launch_bounds__(256)
bn_bwd_apply_mlp_kernel(BnBwdArgs a, const float* __restrict__ coef, __nv_bfloat16* __restrict__ dx) {
  pdl_trigger();
  pdl_wait();
  a.seed = pe_salted(a.seed);
  const PoolGeom& g = a.g;
  const int cg = g.C >> 3;
  const int ry = blockDim.x / cg;
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  const long long nwin = g.rows * g.Wo;
  const long long stride = (long long)gridDim.x * ry;
  const long long w_first = (long long)blockIdx.x * ry + ty;
  uint4 xv[U][K], dv[U], adv[AUX ? U : 1];
  uint2 aidx[AUX ? U : 1];
  int aj0[AUX ? U : 1];
  long long wrow[U];
  int wwo[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long w = w_first + u * stride;
    const long long row = (long long)((unsigned)w / (unsigned)g.Wo);  // the launcher guarantees nwin < 2^31
    wrow[u] = w < nwin ? row : -1;
    wwo[u] = (int)(w - row * g.Wo);
    dv[u] = make_uint4(0u, 0u, 0u, 0u);
    if (w < nwin) {
      const __nv_bfloat16* xp = a.x + ((row * g.W + (long long)wwo[u] * K) * g.C + tx * 8);
#pragma unroll
      for (int j = 0; j < K; ++j) xv[u][j] = ld_stream(xp + (long long)j * g.C);
      if (a.dout) dv[u] = ld_stream(a.dout + w * a.ld_dout + a.c_off + tx * 8);
      if (AUX) {  // the K input pixels of a window lie in one auxiliary window (the launcher checks aux_k % K == 0)
        const int w0 = wwo[u] * K;
        const int aw = w0 / a.aux_k;
        aj0[u] = w0 - aw * a.aux_k;
        const long long arow = row * (g.W / a.aux_k) + aw;
        aidx[u] = __ldg(reinterpret_cast<const uint2*>(a.aux_idx + arow * g.C + tx * 8));
        adv[u] = __ldg(reinterpret_cast<const uint4*>(a.aux_dout + arow * a.aux_ld + a.aux_c_off + tx * 8));
      }
    }
  }
  float sc[8], sh[8], A[8], Bc[8];
  {
    const float4* p4 = reinterpret_cast<const float4*>(a.scale + tx * 8);
    const float4 v0 = __ldg(p4), v1 = __ldg(p4 + 1);
    sc[0] = v0.x; sc[1] = v0.y; sc[2] = v0.z; sc[3] = v0.w; sc[4] = v1.x; sc[5] = v1.y; sc[6] = v1.z; sc[7] = v1.w;
    p4 = reinterpret_cast<const float4*>(a.shift + tx * 8);
    const float4 s0 = __ldg(p4), s1 = __ldg(p4 + 1);
    sh[0] = s0.x; sh[1] = s0.y; sh[2] = s0.z; sh[3] = s0.w; sh[4] = s1.x; sh[5] = s1.y; sh[6] = s1.z; sh[7] = s1.w;
    p4 = reinterpret_cast<const float4*>(coef + tx * 8);
    const float4 a0 = __ldg(p4), a1 = __ldg(p4 + 1);
    A[0] = a0.x; A[1] = a0.y; A[2] = a0.z; A[3] = a0.w; A[4] = a1.x; A[5] = a1.y; A[6] = a1.z; A[7] = a1.w;
    p4 = reinterpret_cast<const float4*>(coef + g.C + tx * 8);
    const float4 b0 = __ldg(p4), b1 = __ldg(p4 + 1);
    Bc[0] = b0.x; Bc[1] = b0.y; Bc[2] = b0.z; Bc[3] = b0.w; Bc[4] = b1.x; Bc[5] = b1.y; Bc[6] = b1.z; Bc[7] = b1.w;
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (wrow[u] < 0) continue;
    const long long row = wrow[u];
    const int wo = wwo[u];
    uint32_t km = 0xFFu;
    if (EXTRA && a.drop_thresh) {
      const unsigned long long e0 = (unsigned long long)((row * g.Wo + wo) * g.C + tx * 8);
      km = dropout_keep8(a.seed, e0 >> 3, a.drop_thresh);
    }
    const long long base = (row * g.W + (long long)wo * K) * g.C + tx * 8;
    float o[K][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float go = bf16_at(dv[u], i);
      if (EXTRA && a.dout_seq)
        go += __bfloat162float(a.dout_seq[row * ((long long)g.C * g.Wo) + (long long)(tx * 8 + i) * g.Wo + wo]);
      if (EXTRA && a.drop_thresh) go = ((km >> i) & 1u) ? go * a.drop_scale : 0.f;
      float best = -INFINITY, pre = 0.f;
      int js = 0;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const float zp = fmaf(bf16_at(xv[u][j], i), sc[i], sh[i]);
        const float z = zp > 0.f ? zp : zp * a.slope;
        if (z > best) {
          best = z;
          js = j;
          pre = zp;
        }
      }
      const float sg = sc[i] * go * (pre > 0.f ? 1.f : a.slope);
#pragma unroll
      for (int j = 0; j < K; ++j) o[j][i] = fmaf(-A[i], bf16_at(xv[u][j], i), Bc[i]) + (j == js ? sg : 0.f);
      if (AUX) {
        const int ja = (int)(((i < 4 ? aidx[u].x : aidx[u].y) >> (8 * (i & 3))) & 0xFFu) - aj0[u];
        const float da = bf16_at(adv[u], i);
#pragma unroll
        for (int j = 0; j < K; ++j) o[j][i] += (j == ja) ? da : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < K; ++j) st8(dx + base + (long long)j * g.C, o[j]);
    if (wo == g.Wo - 1) {
      for (int j = K; j < g.W - wo * K; ++j) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(a.x + base + (long long)j * g.C));
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = fmaf(-A[i], bf16_at(q, i), Bc[i]);
        st8(dx + base + (long long)j * g.C, t);
      }
    }
  }
}

// aux max-pool backward (model.py:45-49,103-105): dx[argmax window] += dout
__global__ void __launch_bounds__(256)
maxpool_bwd_add_kernel(const __nv_bfloat16* __restrict__ x, PoolGeom g, const __nv_bfloat16* __restrict__ dout,
                       long long ld_dout, int c_off, __nv_bfloat16* __restrict__ dx) {
  pdl_trigger();
  pdl_wait();
  const int cg = g.C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = g.rows * g.Wo * cg;
  if (idx >= total) return;
  const int tx = (int)(idx % cg);
  const int wo = (int)((idx / cg) % g.Wo);
  const long long row = idx / ((long long)cg * g.Wo);
  float best[8];
  int js[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    best[i] = -INFINITY;
    js[i] = 0;
  }
  const long long base = (row * g.W + (long long)wo * g.k) * g.C + tx * 8;
  for (int j = 0; j < g.k; ++j) {
    float v[8];
    ld8(x + base + (long long)j * g.C, v);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (v[i] > best[i]) {
        best[i] = v[i];
        js[i] = j;
      }
  }
  float d[8];
  ld8(dout + (row * g.Wo + wo) * ld_dout + c_off + tx * 8, d);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat16* p = dx + base + (long long)js[i] * g.C + i;
    *p = __float2bfloat16(__bfloat162float(*p) + d[i]);
  }
}

// same with the arg-max positions saved by the forward pass (one byte per output element): x is not read again
__global__ void __launch_bounds__(256)
maxpool_bwd_idx_kernel(const unsigned char* __restrict__ argmax, PoolGeom g, const __nv_bfloat16* __restrict__ dout,
                       long long ld_dout, int c_off, __nv_bfloat16* __restrict__ dx) {
  pdl_trigger();
  pdl_wait();
  const int cg = g.C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = g.rows * g.Wo * cg;
  if (idx >= total) return;
  const int tx = (int)(idx % cg);
  const int wo = (int)((idx / cg) % g.Wo);
  const long long row = idx / ((long long)cg * g.Wo);
  const uint2 jj = *reinterpret_cast<const uint2*>(argmax + (row * g.Wo + wo) * g.C + tx * 8);
  float d[8];
  ld8(dout + (row * g.Wo + wo) * ld_dout + c_off + tx * 8, d);
  const long long base = (row * g.W + (long long)wo * g.k) * g.C + tx * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = (int)(((i < 4 ? jj.x : jj.y) >> (8 * (i & 3))) & 0xFFu);
    __nv_bfloat16* p = dx + base + (long long)j * g.C + i;
    *p = __float2bfloat16(__bfloat162float(*p) + d[i]);
  }
}

}  // namespace pe

// =================================================================================================
// C-ABI
// =================================================================================================
using namespace pe;
#define PE_ST(s) reinterpret_cast<cudaStream_t>(s)
#define PE_LAUNCH_RC() (cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH)

static int stem_pixels_per_cta(long long P) {
  // about 8 CTAs per SM, whole multiples of the 64-pixel step of the kernels
  long long per = (P + 8LL * pe_host::num_sms() - 1) / (8LL * pe_host::num_sms());
  per = ((per + 63) / 64) * 64;
  if (per < 256) per = 256;
  return (int)per;
}

extern "C" int pe_stem_conv_fwd(const float* x, long long sb, long long st, long long sf, int B, int T, int F,
                                const float* w, void* y, double* stats, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!x || !w || !y || B <= 0 || T <= 0 || F <= 0) return PE_ERR_BAD_SHAPE;
  const long long P = (long long)B * T * F;
  if (P >= (1ll << 31)) return PE_ERR_BAD_SHAPE;
  const int per = stem_pixels_per_cta(P);
  pe_host::launch(stem_conv_fwd_kernel, dim3((unsigned)((P + per - 1) / per)), dim3(256), 0, PE_ST(stream), x, sb, st, sf, B, T, F, w,
                                                                                   (__nv_bfloat16*)y, per, stats);
  return PE_LAUNCH_RC();
}

extern "C" int pe_stem_conv_wgrad(const float* x, long long sb, long long st, long long sf, int B, int T, int F,
                                  const void* dy, float* dw, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!x || !dy || !dw || B <= 0 || T <= 0 || F <= 0) return PE_ERR_BAD_SHAPE;
  const long long P = (long long)B * T * F;
  if (P >= (1ll << 31)) return PE_ERR_BAD_SHAPE;
  const int per = stem_pixels_per_cta(P);
  pe_host::launch(stem_conv_wgrad_kernel, dim3((unsigned)((P + per - 1) / per)), dim3(256), 0, PE_ST(stream), 
      x, sb, st, sf, B, T, F, (const __nv_bfloat16*)dy, dw, per);
  return PE_LAUNCH_RC();
}

static bool chan_ok(int C) { return C > 0 && (C % 8) == 0 && C <= 2048; }

extern "C" int pe_bn_stats(const void* x, long long rows, int C, double* sums, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!x || !sums || rows <= 0 || !chan_ok(C) || C / 8 > 256) return PE_ERR_BAD_SHAPE;
  const int per = 2048;
  pe_host::launch(bn_stats_kernel, dim3((unsigned)((rows + per - 1) / per)), dim3(256), 2 * C * sizeof(float), PE_ST(stream), 
      (const __nv_bfloat16*)x, rows, C, sums, per);
  return PE_LAUNCH_RC();
}

extern "C" int pe_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float eps,
                              float momentum, float* scale, float* shift, float* mean, float* rstd,
                              float* running_mean, float* running_var, long long* num_batches_tracked, int C,
                              pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!sums || !gamma || !beta || !scale || !shift || !mean || !rstd || C <= 0 || count <= 0) return PE_ERR_BAD_SHAPE;
  pe_host::launch(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, PE_ST(stream), sums, count, gamma, beta, eps, momentum, scale, shift,
                                                                 mean, rstd, running_mean, running_var,
                                                                 num_batches_tracked, C);
  return PE_LAUNCH_RC();
}

extern "C" int pe_bn_eval_params(const float* gamma, const float* beta, const float* running_mean,
                                 const float* running_var, float eps, float* scale, float* shift, float* mean,
                                 float* rstd, int C, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!gamma || !beta || !running_mean || !running_var || !scale || !shift || C <= 0) return PE_ERR_BAD_SHAPE;
  pe_host::launch(bn_eval_params_kernel, dim3((C + 127) / 128), dim3(128), 0, PE_ST(stream), gamma, beta, running_mean, running_var, eps, scale,
                                                                    shift, mean, rstd, C);
  return PE_LAUNCH_RC();
}

extern "C" int pe_bn_act_pool_fwd(const void* x, long long rows, int W, int C, int k, const float* scale,
                                  const float* shift, float slope, unsigned drop_thresh, float drop_scale,
                                  unsigned long long seed, void* out, long long ld_out, int c_off, void* out_seq,
                                  void* argmax_out, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!x || rows <= 0 || W <= 0 || !chan_ok(C) || k <= 0 || k > W || (!out && !out_seq) || ((scale == 0) != (shift == 0)))
    return PE_ERR_BAD_SHAPE;
  PoolGeom g{rows, W, C, k, W / k};
  const long long total = rows * g.Wo * (C / 8);
  if (argmax_out && k > 255) return PE_ERR_BAD_SHAPE;
  if (!argmax_out && (k == 1 || k == 2 || k == 4) && rows * g.Wo < (1ll << 31) - (1ll << 24)) {
    const int cg = C / 8, ry = 256 / cg, threads = cg * ry;
    const long long nwin = rows * g.Wo;
    const bool extra = drop_thresh != 0 || out_seq != nullptr;
#define PE_FWD(K, U)                                                                                               \
  do {                                                                                                             \
    const unsigned grid = (unsigned)((nwin + (long long)ry * U - 1) / ((long long)ry * U));                        \
    if (extra)                                                                                                     \
      pe_host::launch(bn_act_pool_fwd_mlp_kernel<K, U, true>, dim3(grid), dim3(threads), 0, PE_ST(stream),                                  \
          (const __nv_bfloat16*)x, g, scale, shift, slope, drop_thresh, drop_scale, seed, (__nv_bfloat16*)out,     \
          ld_out, c_off, (__nv_bfloat16*)out_seq);                                                                 \
    else                                                                                                           \
      pe_host::launch(bn_act_pool_fwd_mlp_kernel<K, U, false>, dim3(grid), dim3(threads), 0, PE_ST(stream),                                 \
          (const __nv_bfloat16*)x, g, scale, shift, slope, drop_thresh, drop_scale, seed, (__nv_bfloat16*)out,     \
          ld_out, c_off, (__nv_bfloat16*)out_seq);                                                                 \
  } while (0)
    if (k == 1) PE_FWD(1, 4);
    else if (k == 2) PE_FWD(2, 2);
    else PE_FWD(4, 1);
#undef PE_FWD
    return PE_LAUNCH_RC();
  }
  pe_host::launch(bn_act_pool_fwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, PE_ST(stream), 
      (const __nv_bfloat16*)x, g, scale, shift, slope, drop_thresh, drop_scale, seed, (__nv_bfloat16*)out, ld_out, c_off,
      (__nv_bfloat16*)out_seq, (unsigned char*)argmax_out);
  return PE_LAUNCH_RC();
}

extern "C" int pe_bn_act_pool_bwd(const void* x, long long rows, int W, int C, int k, const float* scale,
                                  const float* shift, const float* mean, const float* rstd, float slope,
                                  unsigned drop_thresh, float drop_scale, unsigned long long seed, const void* dout,
                                  long long ld_dout, int c_off, const void* dout_seq, double* sums /* [2][C], zeroed */,
                                  int sums_ready, float* coef /* [2][C] scratch */, float* dgamma, float* dbeta,
                                  const void* aux_argmax, const void* aux_dout, long long aux_ld, int aux_c_off,
                                  int aux_k, void* dx, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!x || !scale || !shift || !mean || !rstd || !sums || !coef || !dx || rows <= 0 || W <= 0 || !chan_ok(C) ||
      C / 8 > 256 || (k != 1 && k != 2 && k != 4) || k > W || (!dout && !dout_seq) ||
      rows * (W / k) >= (1ll << 31) - (1ll << 24))
    return PE_ERR_BAD_SHAPE;
  BnBwdArgs a{};
  a.x = (const __nv_bfloat16*)x;
  a.g = PoolGeom{rows, W, C, k, W / k};
  a.scale = scale; a.shift = shift;
  a.slope = slope; a.drop_thresh = drop_thresh; a.drop_scale = drop_scale; a.seed = seed;
  a.dout = (const __nv_bfloat16*)dout; a.ld_dout = ld_dout; a.c_off = c_off;
  a.dout_seq = (const __nv_bfloat16*)dout_seq;
  const bool aux = aux_argmax != nullptr;
  if (aux) {  // joins only the plain (no dropout / sequence gradient) layers, whole pool windows inside an aux window
    if (!aux_dout || aux_k <= 0 || aux_k > 255 || (aux_k % k) || (W % aux_k) || (W % k) || drop_thresh || dout_seq)
      return PE_ERR_BAD_SHAPE;
    a.aux_idx = (const unsigned char*)aux_argmax; a.aux_dout = (const __nv_bfloat16*)aux_dout;
    a.aux_ld = aux_ld; a.aux_c_off = aux_c_off; a.aux_k = aux_k;
  }
  const long long nwin = rows * a.g.Wo;
  // pass-1 CTAs: about 8 per SM, at least 8 windows per thread row (the fp64 atomics of a CTA cost ~2*C operations)
  long long per_ll = (nwin + 8LL * pe_host::num_sms() - 1) / (8LL * pe_host::num_sms());
  const int ry1 = 256 / (C / 8);
  if (per_ll < 8 * ry1) per_ll = 8 * ry1;
  if (per_ll > 4096) per_ll = 4096;
  const int per = (int)per_ll;
  const unsigned g1 = (unsigned)((nwin + per - 1) / per);
  const long long total = nwin * (C / 8);
  const unsigned g2 = (unsigned)((total + 255) / 256);
  const size_t sm1 = 4 * C * sizeof(float), sm2 = 4 * C * sizeof(float);
  cudaStream_t st = PE_ST(stream);
#define PE_BN_BWD(K, U)                                                                                    \
  do {                                                                                                    \
    if (!sums_ready && extra) pe_host::launch(bn_bwd_reduce_kernel<K, true>, dim3(g1), dim3(256), sm1, st, a, sums, per);          \
    else if (!sums_ready) pe_host::launch(bn_bwd_reduce_kernel<K, false>, dim3(g1), dim3(256), sm1, st, a, sums, per);             \
    pe_host::launch(bn_bwd_params_kernel, dim3((C + 127) / 128), dim3(128), 0, st, sums, (double)rows * W, scale, mean, rstd, dgamma, dbeta, \
                                                          coef, C);                                       \
    const unsigned ga = (unsigned)((nwin + (long long)ry * U - 1) / ((long long)ry * U));                 \
    if (aux)                                                                                              \
      pe_host::launch(bn_bwd_apply_mlp_kernel<K, U, false, true>, dim3(ga), dim3(cgs * ry), 0, st, a, coef, (__nv_bfloat16*)dx);   \
    else if (extra)                                                                                       \
      pe_host::launch(bn_bwd_apply_mlp_kernel<K, U, true, false>, dim3(ga), dim3(cgs * ry), 0, st, a, coef, (__nv_bfloat16*)dx);   \
    else                                                                                                  \
      pe_host::launch(bn_bwd_apply_mlp_kernel<K, U, false, false>, dim3(ga), dim3(cgs * ry), 0, st, a, coef, (__nv_bfloat16*)dx);  \
  } while (0)
  const int cgs = C / 8, ry = 256 / cgs;
  const bool extra = drop_thresh != 0 || dout_seq != nullptr;
  (void)g2; (void)sm2;
  if (k == 1) PE_BN_BWD(1, 2);
  else if (k == 2) PE_BN_BWD(2, 2);
  else PE_BN_BWD(4, 1);
#undef PE_BN_BWD
  return PE_LAUNCH_RC();
}

extern "C" int pe_maxpool_bwd_add(const void* x, const void* argmax, long long rows, int W, int C, int k,
                                  const void* dout, long long ld_dout, int c_off, void* dx, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if ((!x && !argmax) || !dout || !dx || rows <= 0 || W <= 0 || !chan_ok(C) || k <= 0 || k > W) return PE_ERR_BAD_SHAPE;
  PoolGeom g{rows, W, C, k, W / k};
  const long long total = rows * g.Wo * (C / 8);
  if (argmax) {
    pe_host::launch(maxpool_bwd_idx_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, PE_ST(stream), 
        (const unsigned char*)argmax, g, (const __nv_bfloat16*)dout, ld_dout, c_off, (__nv_bfloat16*)dx);
    return PE_LAUNCH_RC();
  }
  pe_host::launch(maxpool_bwd_add_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, PE_ST(stream), 
      (const __nv_bfloat16*)x, g, (const __nv_bfloat16*)dout, ld_dout, c_off, (__nv_bfloat16*)dx);
  return PE_LAUNCH_RC();
}
