// log-mel front-end, fp32 SIMT formulation (v0): reflect pad -> framed, Hann-windowed real DFT as a tiled SGEMM
// against a precomputed windowed basis -> |.|^2 -> mel filterbank -> (log(1e-5 + .) + 4) / 4.
// Mirrors torchaudio.transforms.MelSpectrogram as configured at reference meldataset.py:77 and the
// normalisation at meldataset.py:650.
#include "common.cuh"
#include "../../include/pitchextractor_b200.h"

namespace pe {

constexpr int DFT_TM = 64;  // frames per CTA
constexpr int DFT_TN = 64;  // basis columns per CTA (= 32 bins, re/im interleaved)
constexpr int DFT_TK = 16;

__device__ __forceinline__ int reflect_idx(int i, int L) {
  if (i < 0) i = -i;
  if (i >= L) i = 2 * (L - 1) - i;
  return i;
}

// power[f][bin] for f in [0, B*T), bin in [0, n_bins)
__global__ void __launch_bounds__(256)
dft_power_kernel(const float* __restrict__ wave, int B, int L, int T, int n_fft, int hop,
                 const float* __restrict__ basis, int ld_basis, int n_bins, float* __restrict__ power,
                 const int* __restrict__ lengths) {
  __shared__ __align__(16) float As[DFT_TK][DFT_TM + 4];
  __shared__ __align__(16) float Bs[DFT_TK][DFT_TN];
  const int tid = threadIdx.x;
  const int f0 = blockIdx.x * DFT_TM;
  const int n0 = blockIdx.y * DFT_TN;
  const int total_frames = B * T;
  const int half = n_fft / 2;

  // loader roles
  const int lf = tid >> 2;         // frame within tile 0..63
  const int lk = (tid & 3) * 4;    // k offset 0,4,8,12
  const int gf = f0 + lf;
  const bool f_ok = gf < total_frames;
  const int fb_ = f_ok ? gf / T : 0;
  const int ft = f_ok ? gf - fb_ * T : 0;
  const float* wrow = wave + (size_t)fb_ * L;
  const int len = lengths ? min(max(lengths[fb_], 1), L) : L;  // reflect at the item's own end
  const int start = ft * hop - half;  // sample index of k = 0
  const int bk = tid >> 4;         // basis row within slice 0..15
  const int bc = (tid & 15) * 4;   // basis col 0..60

  // compute roles: 4 frames x 4 columns per thread
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < n_fft; k0 += DFT_TK) {
    float a4[4];
    const int s = start + k0 + lk;
    if (!f_ok) {
      a4[0] = a4[1] = a4[2] = a4[3] = 0.f;
    } else if (s >= 0 && s + 3 < len && ((s & 3) == 0) && ((L & 3) == 0)) {
      const float4 v = *reinterpret_cast<const float4*>(wrow + s);
      a4[0] = v.x; a4[1] = v.y; a4[2] = v.z; a4[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) a4[j] = wrow[min(max(reflect_idx(s + j, len), 0), len - 1)];
    }
    const float4 b4 = *reinterpret_cast<const float4*>(basis + (size_t)(k0 + bk) * ld_basis + n0 + bc);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) As[lk + j][lf] = a4[j];
    *reinterpret_cast<float4*>(&Bs[bk][bc]) = b4;
    __syncthreads();
    // two-level accumulation (slice partial, then running sum) keeps fp32 error growth ~sqrt(16)+sqrt(64)
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int k = 0; k < DFT_TK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = f0 + ty * 4 + i;
    if (f >= total_frames) continue;
#pragma unroll
    for (int j = 0; j < 4; j += 2) {
      const int bin = (n0 + tx * 4 + j) >> 1;
      if (bin < n_bins) power[(size_t)f * n_bins + bin] = acc[i][j] * acc[i][j] + acc[i][j + 1] * acc[i][j + 1];
    }
  }
}

constexpr int MEL_FR = 8;  // frames per CTA

__global__ void __launch_bounds__(128)
mel_log_kernel(const float* __restrict__ power, int B, int T, int n_bins, int n_mels, const float* __restrict__ fb,
               float* __restrict__ out_bmt, float* __restrict__ out_btm, const int* __restrict__ crop, int T_out,
               const int* __restrict__ lengths, int L, int hop, int n_fft) {
  extern __shared__ float ps[];  // [MEL_FR][n_bins]
  const int b = blockIdx.y;
  int Tb = T;
  if (lengths) {  // frames of this item: 1 + len / hop (none if the reflect padding would not fit)
    const int len = min(max(lengths[b], 0), L);
    Tb = len > n_fft / 2 ? min(T, 1 + len / hop) : 0;
  }
  const int t0 = blockIdx.x * MEL_FR;
  const int c = crop ? crop[b] : 0;
  for (int i = threadIdx.x; i < MEL_FR * n_bins; i += blockDim.x) {
    const int fr = i / n_bins, k = i - fr * n_bins;
    const int src = c + t0 + fr;
    ps[i] = (t0 + fr < T_out && src < Tb) ? power[((size_t)b * T + src) * n_bins + k] : 0.f;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    float acc[MEL_FR];
#pragma unroll
    for (int fr = 0; fr < MEL_FR; ++fr) acc[fr] = 0.f;
    for (int k = 0; k < n_bins; ++k) {
      const float w = __ldg(fb + (size_t)k * n_mels + m);
      if (w != 0.f) {
#pragma unroll
        for (int fr = 0; fr < MEL_FR; ++fr) acc[fr] = fmaf(ps[fr * n_bins + k], w, acc[fr]);
      }
    }
#pragma unroll
    for (int fr = 0; fr < MEL_FR; ++fr) {
      const int t = t0 + fr;
      if (t >= T_out) continue;
      const bool pad = (c + t) >= Tb;  // Collater zero padding (meldataset.py:804-816)
      const float y = pad ? 0.f : (logf(1e-5f + acc[fr]) + 4.0f) * 0.25f;
      if (out_bmt) out_bmt[((size_t)b * n_mels + m) * T_out + t] = y;
      if (out_btm) out_btm[((size_t)b * T_out + t) * n_mels + m] = y;
    }
  }
}

}  // namespace pe

extern "C" int pe_logmel_f32(const float* wave, int B, int L, int n_fft, int hop, int n_mels, const float* basis,
                             int ld_basis, const float* fb, float* power, size_t power_bytes, float* out_bmt,
                             float* out_btm, const int* crop, const int* lengths, int T_out, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!wave || !basis || !fb || !power || B <= 0 || L <= 0 || n_fft < 16 || (n_fft % 16) || hop <= 0 || n_mels <= 0)
    return PE_ERR_BAD_SHAPE;
  if (L <= n_fft / 2) return PE_ERR_BAD_SHAPE;  // reflect padding needs pad < length (torch.stft contract)
  const int n_bins = n_fft / 2 + 1;
  const int T = 1 + L / hop;
  if (T_out <= 0) T_out = T;
  const int ncol_tiles = (2 * n_bins + pe::DFT_TN - 1) / pe::DFT_TN;
  if (ld_basis < ncol_tiles * pe::DFT_TN || (ld_basis % 4)) return PE_ERR_BAD_SHAPE;
  if (power_bytes < (size_t)B * T * n_bins * sizeof(float)) return PE_ERR_WORKSPACE;
  if (!out_bmt && !out_btm) return PE_ERR_BAD_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 g1((B * T + pe::DFT_TM - 1) / pe::DFT_TM, ncol_tiles);
  pe::dft_power_kernel<<<g1, 256, 0, st>>>(wave, B, L, T, n_fft, hop, basis, ld_basis, n_bins, power, lengths);
  dim3 g2((T_out + pe::MEL_FR - 1) / pe::MEL_FR, B);
  const size_t smem = (size_t)pe::MEL_FR * n_bins * sizeof(float);
  pe::mel_log_kernel<<<g2, 128, smem, st>>>(power, B, T, n_bins, n_mels, fb, out_bmt, out_btm, crop, T_out, lengths, L, hop,
                                           n_fft);
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}
