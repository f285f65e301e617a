// Polyphase windowed-sinc resampling on the GPU: the arithmetic of torchaudio.functional.resample
// (torchaudio/functional/functional.py: _get_sinc_resample_kernel + _apply_sinc_resample_kernel), which the reference calls
// per file on the CPU at meldataset.py:621-627:
//     y[b][j * up + p] = sum_{k < K} xpad[b][j * down + k] * h[p][k],   xpad = zero-pad(x, width, width + down)
// with up = new_freq / gcd, down = orig_freq / gcd, K = 2 * width + down and the filter bank h [up][K] built on the host
// exactly as torchaudio builds it.  One thread per output sample; a CTA's input span is staged in shared memory once and
// the filter rows stream through the read-only cache (they are shared by every item and CTA).
#include "common.cuh"
#include "../../include/pitchextractor_b200.h"

namespace pe {

constexpr int RS_THREADS = 256;

__global__ void __launch_bounds__(RS_THREADS)
resample_kernel(const float* __restrict__ x, long long ldx, const int* __restrict__ lengths, int L,
                const float* __restrict__ h, int up, int down, int width, int K, float* __restrict__ y, long long ldy,
                int L_out) {
  extern __shared__ float xs[];
  const int b = blockIdx.y;
  const int len = lengths ? min(max(lengths[b], 0), L) : L;
  const long long n0 = (long long)blockIdx.x * RS_THREADS;     // first output sample of this CTA
  const long long j0 = n0 / up;                                 // first polyphase block
  const long long j1 = (n0 + RS_THREADS - 1) / up;              // last one
  const int span = (int)(j1 - j0) * down + K;                   // input samples this CTA touches
  const long long s0 = j0 * down - width;                       // index of xs[0] in the un-padded signal
  const float* xb = x + (long long)b * ldx;
  for (int i = threadIdx.x; i < span; i += RS_THREADS) {
    const long long s = s0 + i;
    xs[i] = (s >= 0 && s < len) ? __ldg(xb + s) : 0.f;
  }
  __syncthreads();
  const long long n = n0 + threadIdx.x;
  if (n >= L_out) return;
  const long long j = n / up;
  const int p = (int)(n - j * up);
  const float* hp = h + (size_t)p * K;
  const float* xp = xs + (int)(j - j0) * down;
  float acc0 = 0.f, acc1 = 0.f;
  int k = 0;
  for (; k + 1 < K; k += 2) {
    acc0 = fmaf(xp[k], __ldg(hp + k), acc0);
    acc1 = fmaf(xp[k + 1], __ldg(hp + k + 1), acc1);
  }
  if (k < K) acc0 = fmaf(xp[k], __ldg(hp + k), acc0);
  y[(long long)b * ldy + n] = acc0 + acc1;
}

}  // namespace pe

extern "C" int pe_resample_f32(const float* x, long long ldx, const int* lengths, int B, int L, const float* h, int up,
                               int down, int width, float* y, long long ldy, int L_out, pe_stream_t stream) {
  using namespace pe;
  if (int rc = pe_host::check_arch()) return rc;
  if (!x || !h || !y || B <= 0 || L <= 0 || up <= 0 || down <= 0 || width <= 0 || L_out <= 0 || ldx < L || ldy < L_out)
    return PE_ERR_BAD_SHAPE;
  const int K = 2 * width + down;
  const int span_max = ((RS_THREADS + up - 1) / up + 1) * down + K;
  const size_t smem = (size_t)span_max * sizeof(float);
  if (smem > 200 * 1024) return PE_ERR_BAD_SHAPE;  // absurd ratio (down > ~150 per output sample block)
  static size_t attr_set = 0;
  if (smem > 48 * 1024 && smem > attr_set) {
    if (cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return PE_ERR_LAUNCH;
    attr_set = smem;
  }
  dim3 grid((unsigned)((L_out + RS_THREADS - 1) / RS_THREADS), (unsigned)B);
  resample_kernel<<<grid, RS_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(x, ldx, lengths, L, h, up, down,
                                                                                      width, K, y, ldy, L_out);
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}
