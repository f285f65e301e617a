// Parameter-side passes: fused AdamW over the flat fp32 parameter arena (optimizers.py:54-64 -> torch.optim.AdamW),
// with the bf16 working copy of the weights written in the same pass, and the small layout transforms that
// produce tensor-core operand layouts of the convolution weights.
#include "common.cuh"
#include "../../include/pitchextractor_b200.h"

namespace pe {

// torch.optim.AdamW (single-tensor formulation): p *= 1 - lr*wd; m, v update; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
             float grad_scale, __nv_bfloat16* __restrict__ p_bf16) {
  pdl_trigger();
  pdl_wait();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    float4 pp = *reinterpret_cast<float4*>(p + i);
    const float4 gg = *reinterpret_cast<const float4*>(g + i);
    float4 mm = *reinterpret_cast<float4*>(m + i);
    float4 vv = *reinterpret_cast<float4*>(v + i);
    float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ga[4] = {gg.x, gg.y, gg.z, gg.w};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ga[j] * grad_scale;
      pa[j] *= (1.f - lr * wd);
      ma[j] = ma[j] + (1.f - beta1) * (gr - ma[j]);  // lerp, as torch does
      va[j] = beta2 * va[j] + (1.f - beta2) * gr * gr;
      const float denom = sqrtf(va[j]) / bc2_sqrt + eps;
      pa[j] -= (lr / bc1) * (ma[j] / denom);
    }
    *reinterpret_cast<float4*>(p + i) = make_float4(pa[0], pa[1], pa[2], pa[3]);
    *reinterpret_cast<float4*>(m + i) = make_float4(ma[0], ma[1], ma[2], ma[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(va[0], va[1], va[2], va[3]);
    if (p_bf16) {
      *reinterpret_cast<uint2*>(p_bf16 + i) = make_uint2(pack_bf16(pa[0], pa[1]), pack_bf16(pa[2], pa[3]));
    }
  } else {
    for (long long k = i; k < n; ++k) {
      const float gr = g[k] * grad_scale;
      float pv = p[k] * (1.f - lr * wd);
      const float mv = m[k] + (1.f - beta1) * (gr - m[k]);
      const float vv = beta2 * v[k] + (1.f - beta2) * gr * gr;
      pv -= (lr / bc1) * (mv / (sqrtf(vv) / bc2_sqrt + eps));
      p[k] = pv; m[k] = mv; v[k] = vv;
      if (p_bf16) p_bf16[k] = __float2bfloat16(pv);
    }
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  pdl_trigger();
  pdl_wait();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    *reinterpret_cast<uint2*>(y + i) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  } else {
    for (long long k = i; k < n; ++k) y[k] = __float2bfloat16(x[k]);
  }
}

// forward operand: out[co][0 : 9*C1] = w[co][tap][ci], out[co][9*C1 : 9*C1+C2] = w2[co][ci2]   (bf16)
__global__ void conv_fwd_weight_kernel(const float* __restrict__ w, int Cout, int K1, const float* __restrict__ w2,
                                       int C2, __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int K = K1 + C2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * K) return;
  const int co = (int)(i / K), k = (int)(i % K);
  out[i] = __float2bfloat16(k < K1 ? w[(long long)co * K1 + k] : w2[(long long)co * C2 + (k - K1)]);
}

// data-gradient operand: out[ci][tap'][co] = w[co][8 - tap'][ci]; out[ci][9*Cout + co2] = w2[co2][ci]   (bf16)
__global__ void conv_dgrad_weight_kernel(const float* __restrict__ w, int Cout, int Cin, const float* __restrict__ w2,
                                         int Cout2, __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int K = 9 * Cout + Cout2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cin * K) return;
  const int ci = (int)(i / K), k = (int)(i % K);
  float v;
  if (k < 9 * Cout) {
    const int tp = k / Cout, co = k % Cout;
    v = w[((long long)co * 9 + (8 - tp)) * Cin + ci];
  } else {
    v = w2[(long long)(k - 9 * Cout) * Cin + ci];
  }
  out[i] = __float2bfloat16(v);
}

}  // namespace pe

using namespace pe;
#define PE_ST(s) reinterpret_cast<cudaStream_t>(s)
#define PE_LAUNCH_RC() (cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH)

extern "C" int pe_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                        float eps, float weight_decay, long long step, float grad_scale, void* p_bf16,
                        pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!p || !g || !m || !v || n <= 0 || step <= 0) return PE_ERR_BAD_SHAPE;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const long long threads = (n + 3) / 4;
  pe_host::launch(adamw_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, PE_ST(stream), 
      p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale,
      (__nv_bfloat16*)p_bf16);
  return PE_LAUNCH_RC();
}

extern "C" int pe_cast_bf16(const float* x, void* y, long long n, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!x || !y || n <= 0) return PE_ERR_BAD_SHAPE;
  const long long threads = (n + 3) / 4;
  pe_host::launch(cast_bf16_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, PE_ST(stream), x, (__nv_bfloat16*)y, n);
  return PE_LAUNCH_RC();
}

extern "C" int pe_conv_weight_prep(const float* w, int Cout, int Cin, const float* w2, int C2, void* w_fwd,
                                   void* w_dgrad, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!w || Cout <= 0 || Cin <= 0 || (C2 > 0 && !w2) || (!w_fwd && !w_dgrad)) return PE_ERR_BAD_SHAPE;
  if (w_fwd) {
    const long long n = (long long)Cout * (9 * Cin + C2);
    pe_host::launch(conv_fwd_weight_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, PE_ST(stream), w, Cout, 9 * Cin, w2, C2,
                                                                                  (__nv_bfloat16*)w_fwd);
  }
  if (w_dgrad) {
    // fused data gradient of (3x3 conv on x) + (1x1 conv on x2) with respect to ... see header
    const long long n = (long long)Cin * (9 * Cout + C2);
    pe_host::launch(conv_dgrad_weight_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, PE_ST(stream), w, Cout, Cin, w2, C2,
                                                                                    (__nv_bfloat16*)w_dgrad);
  }
  return PE_LAUNCH_RC();
}
