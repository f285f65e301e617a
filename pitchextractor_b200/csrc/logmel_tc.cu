// log-mel front-end on tcgen05 tensor cores (n_fft = 1024): the 1024-point real DFT of TWO frames is one complex FFT
// (z = x_a + i x_b), factorised 32 x 32 (four-step): two 32-point DFT stages, each a [128 x 64] x [64 x 64] real GEMM
// over 4 frame pairs, with the 1024-th-root twiddles applied between them.  fp32-grade accuracy comes from a two-way
// fp16 split of both operands (hi*hi + hi*lo + lo*hi, the lo parts pre-scaled by 2^11), accumulated in fp32 in TMEM:
// 24 tcgen05.mma (kind::f16, 128x64x16) per 8 frames instead of the 6.3 MFLOP/frame of the dense DFT.
// Frames arrive as 4 KB bulk async copies (cp.async.bulk) straight from the reflect-padded waveform (frame stride =
// hop), double buffered; windowing + the fp16 split write them in the MN-major operand order stage 1 needs, so no
// transposition happens before the first GEMM.  Pair unpacking, |.|^2, the banded mel filterbank, log
// and normalisation run on the TMEM rows in registers / shared memory.  Mirrors torchaudio MelSpectrogram as built at
// reference meldataset.py:77 and the normalisation at meldataset.py:650.
#include "common.cuh"
#include "../../include/pitchextractor_b200.h"
#include <cuda_fp16.h>

namespace pe {

constexpr int LM_NFFT = 1024;
constexpr int LM_FR = 8;              // frames per tile (4 pairs)
constexpr int LM_WORKERS = 256;       // 8 worker warps: two per TMEM lane quarter
constexpr int LM_THREADS = 64 + LM_WORKERS;  // warp 0 TMA, warp 1 MMA, warps 2..9 workers
constexpr int LM_RAW_B = LM_FR * 4096;  // 32 KB of fp32 samples per tile
constexpr int LM_NRAW = 3;            // raw-sample ring depth (tiles in flight)
constexpr int LM_OP_B = 16384;        // one fp16 [128 x 64] operand
constexpr int LM_PSTRIDE = 516;

struct LmParams {
  int B, T, n_mels, T_out, tiles_per_item, num_tiles;
  const float* win;        // [1024]
  const __half* fmat;      // 3 x 8 KB pre-swizzled K-major [64 x 64]: hi, lo, hi * 2^-11
  const float* tw;         // [2][32][32]: cos, sin of 2 pi k1 n2 / 1024, indexed [k1][n2]
  const int* mel_start;    // [n_mels]
  const int* mel_count;    // [n_mels]
  const int* mel_off;      // [n_mels] offset into mel_w
  const float* mel_w;      // banded filter weights
  int mel_nnz;
  const int* crop;         // [B] or NULL
  float* out_bmt;          // [B][n_mels][T_out] or NULL
  float* out_btm;          // [B][T_out][n_mels] or NULL
  long long* dbg;          // optional per-CTA phase cycle counters (tuning)
};

__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  tc_mma_bf16(tmem_d, da, db, idesc, acc);  // same instruction (kind::f16); operand formats live in idesc
}
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// split x into fp16 hi and fp16 lo' = (x - hi) * 2^11
__device__ __forceinline__ void split16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * 2048.0f);
}

__global__ void __launch_bounds__(LM_THREADS, 1)
logmel_tc_kernel(const float* __restrict__ wave, int L, long long ld, int hop, const LmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_raw = smem;                          // LM_NRAW x 32 KB
  uint8_t* s_ahi = smem + LM_NRAW * LM_RAW_B;     // 16 KB
  uint8_t* s_alo = s_ahi + LM_OP_B;               // 16 KB
  uint8_t* s_f = s_alo + LM_OP_B;                 // 24 KB: F_hi | F_lo | F_hi'
  float* s_tw = reinterpret_cast<float*>(s_f + 3 * 8192);   // 8 KB
  float* s_win = s_tw + 2048;                                // 4 KB
  float* s_p = s_win + LM_NFFT;                              // [8][516] power spectra
  float* s_melw = s_p + LM_FR * LM_PSTRIDE;                  // banded mel weights (<= 2048)
  int* s_meli = reinterpret_cast<int*>(s_melw + 2048);       // start | count | off, 3 x 128
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_meli + 384);
  uint64_t* raw_full = bars;                  // [LM_NRAW]
  uint64_t* raw_empty = bars + LM_NRAW;       // [LM_NRAW]
  uint64_t* work_ready = bars + 2 * LM_NRAW;  // workers -> MMA
  uint64_t* mma_done = work_ready + 1;        // MMA -> workers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // constant tables -> smem (all threads)
  for (int i = tid; i < 3 * 8192 / 16; i += LM_THREADS)
    reinterpret_cast<uint4*>(s_f)[i] = __ldg(reinterpret_cast<const uint4*>(p.fmat) + i);
  for (int i = tid; i < 2048; i += LM_THREADS) s_tw[i] = p.tw[i];
  for (int i = tid; i < LM_NFFT; i += LM_THREADS) s_win[i] = p.win[i];
  for (int i = tid; i < p.mel_nnz; i += LM_THREADS) s_melw[i] = p.mel_w[i];
  for (int i = tid; i < p.n_mels; i += LM_THREADS) {
    s_meli[i] = p.mel_start[i];
    s_meli[128 + i] = p.mel_count[i];
    s_meli[256 + i] = p.mel_off[i];
  }
  if (tid == 0) {
    for (int i = 0; i < LM_NRAW; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], LM_WORKERS / 32);
    }
    mbar_init(work_ready, LM_WORKERS / 32);
    mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 128);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer: one box per tile
    if (lane == 0) {
      int buf = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&raw_empty[buf], ph ^ 1u);
        const int b = tile / p.tiles_per_item, t0 = (tile - b * p.tiles_per_item) * LM_FR;
        const int nfr = min(LM_FR, p.T - t0);
        // interior frames (window fully inside the item) are bulk-copied; the few frames that touch the reflect
        // padding are gathered by the workers straight from global memory
        int n_in = 0;
        for (int j = 0; j < nfr; ++j) {
          const long long s0 = (long long)(t0 + j) * hop - LM_NFFT / 2;
          n_in += (s0 >= 0 && s0 + LM_NFFT <= L) ? 1 : 0;
        }
        mbar_arrive_expect_tx(&raw_full[buf], (uint32_t)n_in * 4096u);
        for (int j = 0; j < nfr; ++j) {
          const long long s0 = (long long)(t0 + j) * hop - LM_NFFT / 2;
          if (s0 >= 0 && s0 + LM_NFFT <= L)
            bulk_load_1d(s_raw + buf * LM_RAW_B + j * 4096, wave + (long long)b * ld + s0, 4096, &raw_full[buf]);
        }
        if (++buf == LM_NRAW) {
          buf = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer: 12 + 12 MMAs per tile
    if (lane == 0) {
      constexpr uint32_t IDESC1 = umma_idesc(UMMA_F16, 128, 64, 1, 0);  // A MN-major (frames as landed), B K-major
      constexpr uint32_t IDESC2 = umma_idesc(UMMA_F16, 128, 64, 0, 0);
      const uint32_t ahi = smem_u32(s_ahi), alo = smem_u32(s_alo), f = smem_u32(s_f);
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        // stage 1: D1[(pair,n2)][(c',k1)]   A: 2 MN atoms (8 KB apart) x 64 K rows; 16 K rows per MMA
        mbar_wait(work_ready, ph);
        ph ^= 1u;
        tc_fence_after();
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
          const uint32_t a = prod == 2 ? alo : ahi;
          const uint32_t b = f + (prod == 0 ? 0 : prod == 1 ? 8192 : 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tm, umma_desc_sw128(a + k * 2048, 8192, 1024), umma_desc_sw128(b + k * 32, 16, 1024), IDESC1,
                       (prod > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(mma_done);
        // stage 2: D2[(pair,k1)][(c',k2)]   A K-major
        mbar_wait(work_ready, ph);
        ph ^= 1u;
        tc_fence_after();
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
          const uint32_t a = prod == 2 ? alo : ahi;
          const uint32_t b = f + (prod == 0 ? 0 : prod == 1 ? 8192 : 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tm + 64, umma_desc_sw128(a + k * 32, 16, 1024), umma_desc_sw128(b + k * 32, 16, 1024), IDESC2,
                       (prod > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(mma_done);
      }
    }
  } else {
    // ---------------------------------------------------------------- workers (128 threads)
    const int wt = tid - 64;               // 0..255
    const int q = warp & 3;                // TMEM lane quarter == frame pair handled in the TMEM phases
    const int half = (warp - 2) >> 2;      // the two warps of a quarter split the columns
    const uint32_t trow = tm + ((uint32_t)(q * 32) << 16);
    uint32_t dph = 0, rph = 0;
    int buf = 0;
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tc0 = clock64();
#define LM_TICK(k) do { if (p.dbg) { const long long c_ = clock64(); tph[k] += c_ - tc0; tc0 = c_; } } while (0)
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_item, t0 = (tile - b * p.tiles_per_item) * LM_FR;
      // ---- pre-pass: window, split, lay out as the stage-1 A operand (MN-major fp16)
      mbar_wait(&raw_full[buf], rph);
      LM_TICK(0);
      const uint8_t* raw = s_raw + buf * LM_RAW_B;
#pragma unroll 4
      for (int it = 0; it < 2048 / LM_WORKERS; ++it) {
        const int u = wt + LM_WORKERS * it;              // 16-byte unit of the raw tile: [frame][n1][n2 / 4]
        const int fr = u >> 8, n1 = (u >> 3) & 31, n2 = (u & 7) << 2;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);     // frames past the end of the item contribute zeros
        if (t0 + fr < p.T) {
          const long long s0 = (long long)(t0 + fr) * hop - LM_NFFT / 2;
          if (s0 >= 0 && s0 + LM_NFFT <= L) {
            x = *reinterpret_cast<const float4*>(raw + u * 16);
          } else {  // reflect padding (torch.stft center=True, pad_mode="reflect")
            const float* wb = wave + (long long)b * ld;
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              long long si = s0 + n1 * 32 + n2 + j;
              if (si < 0) si = -si;
              if (si >= L) si = 2LL * (L - 1) - si;
              e[j] = __ldg(wb + si);
            }
            x = make_float4(e[0], e[1], e[2], e[3]);
          }
        }
        const float4 w = *reinterpret_cast<const float4*>(s_win + n1 * 32 + n2);
        __half h[4], l[4];
        split16(x.x * w.x, h[0], l[0]);
        split16(x.y * w.y, h[1], l[1]);
        split16(x.z * w.z, h[2], l[2]);
        split16(x.w * w.w, h[3], l[3]);
        const int pr = fr >> 1, kk = (fr & 1) * 32 + n1;
        const int unit = (pr & 1) * 4 + (n2 >> 3);
        const uint32_t off = (uint32_t)((pr >> 1) * 8192 + kk * 128 + ((unit ^ (kk & 7)) << 4) + (n2 & 7) * 2);
        *reinterpret_cast<uint2*>(s_ahi + off) = make_uint2(*reinterpret_cast<uint32_t*>(&h[0]) , *reinterpret_cast<uint32_t*>(&h[2]));
        *reinterpret_cast<uint2*>(s_alo + off) = make_uint2(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[2]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&raw_empty[buf]);
        mbar_arrive(work_ready);
      }
      if (++buf == LM_NRAW) {
        buf = 0;
        rph ^= 1u;
      }
      LM_TICK(1);
      // ---- stage-1 result: twiddle, split, transpose into the stage-2 A operand (K-major fp16)
      mbar_wait(mma_done, dph);
      dph ^= 1u;
      tc_fence_after();
      LM_TICK(2);
      {
        uint32_t vr[16], vi[16];
        tmem_ld16(trow + 16 * half, vr);
        tmem_ld16(trow + 32 + 16 * half, vi);
        tmem_ld_wait();
        const int n2 = lane;
#pragma unroll
        for (int kk1 = 0; kk1 < 16; ++kk1) {
          const int k1 = 16 * half + kk1;
          const float ct = s_tw[k1 * 32 + n2], st = s_tw[1024 + k1 * 32 + n2];
          const float re = __uint_as_float(vr[kk1]), im = __uint_as_float(vi[kk1]);
          const float yr = fmaf(re, ct, im * st), yi = fmaf(im, ct, -re * st);
          __half hr, lr, hi_, li;
          split16(yr, hr, lr);
          split16(yi, hi_, li);
          const int m2 = q * 32 + k1;
          const uint32_t row = (uint32_t)(m2 * 128);
          const uint32_t o_re = row + ((((n2 >> 3)) ^ (m2 & 7)) << 4) + (n2 & 7) * 2;
          const uint32_t o_im = row + ((((32 + n2) >> 3) ^ (m2 & 7)) << 4) + (n2 & 7) * 2;
          *reinterpret_cast<__half*>(s_ahi + o_re) = hr;
          *reinterpret_cast<__half*>(s_alo + o_re) = lr;
          *reinterpret_cast<__half*>(s_ahi + o_im) = hi_;
          *reinterpret_cast<__half*>(s_alo + o_im) = li;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(work_ready);
      LM_TICK(3);
      // ---- stage-2 result: unpack the two real spectra of the pair, power
      mbar_wait(mma_done, dph);
      dph ^= 1u;
      tc_fence_after();
      LM_TICK(4);
      {
        uint32_t ur[32], ui[32];
        tmem_ld32(trow + 64, ur);
        tmem_ld32(trow + 96, ui);
        tmem_ld_wait();
        const int src = (32 - lane) & 31;
        float* pa = s_p + (2 * q) * LM_PSTRIDE;
        float* pb = pa + LM_PSTRIDE;
#pragma unroll
        for (int k2 = 0; k2 <= 16; ++k2) {
          if ((k2 > 8) != (half == 1)) continue;  // warp-uniform split of the bins between the two warps of a quarter
          const int c_self = (32 - k2) & 31, c_other = 31 - k2;
          const float sr = lane == 0 ? __uint_as_float(ur[c_self]) : __uint_as_float(ur[c_other & 31]);
          const float si = lane == 0 ? __uint_as_float(ui[c_self]) : __uint_as_float(ui[c_other & 31]);
          const float pr_ = __shfl_sync(0xffffffffu, sr, src);
          const float pi_ = __shfl_sync(0xffffffffu, si, src);
          if (k2 < 16 || lane == 0) {
            const float zr = __uint_as_float(ur[k2 & 31]), zi = __uint_as_float(ui[k2 & 31]);
            const float ar = zr + pr_, ai = zi - pi_, br = zi + pi_, bi = pr_ - zr;
            pa[lane + 32 * k2] = 0.25f * fmaf(ar, ar, ai * ai);
            pb[lane + 32 * k2] = 0.25f * fmaf(br, br, bi * bi);
          }
        }
      }
      tc_fence_before();
      bar_workers();
      LM_TICK(5);
      // ---- banded mel filterbank, log, normalise, store
      const int crop = p.crop ? p.crop[b] : 0;
      // one thread per (mel filter, group of 3 frames): the filter weights are read once for the group
      if (wt < 3 * p.n_mels) {
        const int g3 = wt / p.n_mels, m = wt - g3 * p.n_mels;
        const int f0 = 3 * g3, nf = min(3, LM_FR - f0);
        const float* pp = s_p + f0 * LM_PSTRIDE + s_meli[m];
        const float* ww = s_melw + s_meli[256 + m];
        const int cnt = s_meli[128 + m];
        float acc[3] = {0.f, 0.f, 0.f};
        for (int j = 0; j < cnt; ++j) {
          const float w = ww[j];
          acc[0] = fmaf(pp[j], w, acc[0]);
          acc[1] = fmaf(pp[LM_PSTRIDE + j], w, acc[1]);
          if (nf > 2) acc[2] = fmaf(pp[2 * LM_PSTRIDE + j], w, acc[2]);
        }
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          const int t = t0 + f0 + f, t_out = t - crop;
          if (f >= nf || t >= p.T || t_out < 0 || t_out >= p.T_out) continue;
          const float y = (__logf(1e-5f + acc[f]) + 4.0f) * 0.25f;
          if (p.out_bmt) p.out_bmt[((size_t)b * p.n_mels + m) * p.T_out + t_out] = y;
          if (p.out_btm) p.out_btm[((size_t)b * p.T_out + t_out) * p.n_mels + m] = y;
        }
      }
      bar_workers();
      LM_TICK(6);
    }
    if (p.dbg && wt == 0)
      for (int k = 0; k < 7; ++k) p.dbg[blockIdx.x * 8 + k] = tph[k];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tm, 128);
}

}  // namespace pe

static long long* g_lm_dbg = nullptr;
extern "C" int pe_logmel_set_debug(long long* buf) {
  g_lm_dbg = buf;
  return PE_OK;
}

extern "C" int pe_logmel_tc(const float* wave, int B, int L, int n_fft, int hop, int n_mels, const float* win,
                            const void* fmat, const float* tw, const int* mel_start, const int* mel_count,
                            const int* mel_off, const float* mel_w, int mel_nnz, float* xpad, size_t xpad_bytes,
                            float* out_bmt, float* out_btm, const int* crop, int T_out, pe_stream_t stream) {
  using namespace pe;
  if (int rc = pe_host::check_arch()) return rc;
  if (!wave || !win || !fmat || !tw || !mel_start || !mel_count || !mel_off || !mel_w || B <= 0)
    return PE_ERR_BAD_SHAPE;
  if (n_fft != LM_NFFT || hop <= 0 || (hop % 4) || n_mels <= 0 || 3 * n_mels > LM_WORKERS || mel_nnz <= 0 || mel_nnz > 2048)
    return PE_ERR_BAD_SHAPE;
  if (L <= n_fft / 2 || (!out_bmt && !out_btm)) return PE_ERR_BAD_SHAPE;
  const int T = 1 + L / hop;
  if (T_out <= 0) T_out = T;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // bulk copies need 16-byte aligned rows: items whose length is not a multiple of 4 samples go through a re-strided
  // copy in the workspace (same samples, row stride rounded up); the reflect padding itself is done in the kernel
  const float* src = wave;
  long long ld = L;
  if ((L % 4) || (reinterpret_cast<uintptr_t>(wave) & 15)) {
    ld = ((long long)L + 3) / 4 * 4;
    if (!xpad || xpad_bytes < (size_t)B * ld * sizeof(float)) return PE_ERR_WORKSPACE;
    if (cudaMemcpy2DAsync(xpad, ld * sizeof(float), wave, (size_t)L * sizeof(float), (size_t)L * sizeof(float), B,
                          cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return PE_ERR_LAUNCH;
    src = xpad;
  }
  LmParams p{};
  p.B = B; p.T = T; p.n_mels = n_mels; p.T_out = T_out;
  p.tiles_per_item = (T + LM_FR - 1) / LM_FR;
  p.num_tiles = B * p.tiles_per_item;
  p.win = win; p.fmat = (const __half*)fmat; p.tw = tw;
  p.mel_start = mel_start; p.mel_count = mel_count; p.mel_off = mel_off; p.mel_w = mel_w; p.mel_nnz = mel_nnz;
  p.crop = crop; p.out_bmt = out_bmt; p.out_btm = out_btm;
  p.dbg = g_lm_dbg;
  const size_t smem = LM_NRAW * LM_RAW_B + 2 * LM_OP_B + 3 * 8192 + (2048 + LM_NFFT + LM_FR * LM_PSTRIDE + 2048) * 4 + 384 * 4 +
                      8 * 8 + 1024;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(logmel_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return PE_ERR_LAUNCH;
    attr = true;
  }
  const int grid = p.num_tiles < pe_host::num_sms() ? p.num_tiles : pe_host::num_sms();
  logmel_tc_kernel<<<grid, LM_THREADS, smem, st>>>(src, L, ld, hop, p);
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}
