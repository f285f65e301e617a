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
constexpr int LM_WORKERS = 256;       // 8 worker warps = two independent groups of 4 (one tile each, ping-pong)
constexpr int LM_THREADS = 64 + LM_WORKERS;  // warp 0 TMA, warp 1 MMA, warps 2..9 workers
constexpr int LM_RAW_B = LM_FR * 4096;  // 32 KB of fp32 samples per tile
constexpr int LM_NRAW = 2;            // raw-sample ring depth (a slot is free again right after the pre-pass)
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

// split x into fp16 hi and fp16 lo' = (x - hi) * 2^11
__device__ __forceinline__ void split16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * 2048.0f);
}

__global__ void __launch_bounds__(LM_THREADS, 1)
logmel_tc_kernel(const float* __restrict__ wave, int L, long long ld, int hop, const LmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_raw = smem;                                   // LM_NRAW x 32 KB raw samples
  uint8_t* s_a = smem + LM_NRAW * LM_RAW_B;                // per group: A_hi | A_lo (2 x 16 KB)
  uint8_t* s_f = s_a + 2 * 2 * LM_OP_B;                    // 24 KB: F_hi | F_lo | F_hi'
  float* s_tw = reinterpret_cast<float*>(s_f + 3 * 8192);  // 8 KB
  float* s_win = s_tw + 2048;                              // 4 KB
  float* s_p = s_win + LM_NFFT;                            // per group [8][516] power spectra
  float* s_melw = s_p + 2 * LM_FR * LM_PSTRIDE;            // banded mel weights (<= 2048)
  int* s_meli = reinterpret_cast<int*>(s_melw + 2048);     // start | count | off, 3 x 128
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_meli + 384);
  uint64_t* raw_full = bars;                  // [LM_NRAW]
  uint64_t* raw_empty = bars + LM_NRAW;       // [LM_NRAW]
  uint64_t* work_ready = bars + 2 * LM_NRAW;  // [2] workers of group g -> MMA
  uint64_t* mma_done = work_ready + 2;        // [2] MMA -> workers of group g
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
      mbar_init(&raw_empty[i], 4);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&work_ready[g], 4);
      mbar_init(&mma_done[g], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const int n_local = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: 4 KB bulk copies, one per frame
    // (warp-uniform loop, one elected lane waits and issues: see elect_one())
    for (int i = 0; i < n_local; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int buf = i % LM_NRAW;
      const int b = tile / p.tiles_per_item, t0 = (tile - b * p.tiles_per_item) * LM_FR;
      const int nfr = min(LM_FR, p.T - t0);
      // interior frames (window fully inside the item) are bulk-copied; the few frames that touch the reflect
      // padding are gathered by the workers straight from global memory
      int n_in = 0;
      for (int j = 0; j < nfr; ++j) {
        const long long s0 = (long long)(t0 + j) * hop - LM_NFFT / 2;
        n_in += (s0 >= 0 && s0 + LM_NFFT <= L) ? 1 : 0;
      }
      if (elect_one()) {
        mbar_wait(&raw_empty[buf], ((uint32_t)(i / LM_NRAW) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&raw_full[buf], (uint32_t)n_in * 4096u);
        for (int j = 0; j < nfr; ++j) {
          const long long s0 = (long long)(t0 + j) * hop - LM_NFFT / 2;
          if (s0 >= 0 && s0 + LM_NFFT <= L)
            bulk_load_1d(s_raw + buf * LM_RAW_B + j * 4096, wave + (long long)b * ld + s0, 4096, &raw_full[buf]);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer, serving the two worker groups in
    // the order S1(i), S1(i+1), S2(i), S2(i+1): one group's SIMT phases overlap the other group's GEMM stages
    {
      constexpr uint32_t IDESC1 = umma_idesc(UMMA_F16, 128, 64, 1, 0);  // A MN-major (frames as landed), B K-major
      constexpr uint32_t IDESC2 = umma_idesc(UMMA_F16, 128, 64, 0, 0);
      const uint32_t f = smem_u32(s_f);
      uint32_t ph[2] = {0u, 0u};
      auto stage = [&](int g, int which) {
        const uint32_t ahi = smem_u32(s_a) + g * 2 * LM_OP_B, alo = ahi + LM_OP_B;
        const uint32_t d = tm + g * 128 + which * 64;
        if (elect_one()) {  // warp-uniform loop, one elected lane waits and issues
          mbar_wait(&work_ready[g], ph[g]);
          tc_fence_after();
#pragma unroll
          for (int prod = 0; prod < 3; ++prod) {
            const uint32_t a = prod == 2 ? alo : ahi;
            const uint32_t bm = f + (prod == 0 ? 0 : prod == 1 ? 8192 : 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // stage 1: A = 2 MN atoms (8 KB apart) x 64 K rows, 16 K rows per MMA; stage 2: A K-major
              const uint64_t da = which == 0 ? umma_desc_sw128(a + k * 2048, 8192, 1024) : umma_desc_sw128(a + k * 32, 16, 1024);
              tc_mma_f16(d, da, umma_desc_sw128(bm + k * 32, 16, 1024), which == 0 ? IDESC1 : IDESC2,
                         (prod > 0 || k > 0) ? 1u : 0u);
            }
          }
          tc_commit(&mma_done[g]);
        }
        __syncwarp();
        ph[g] ^= 1u;
      };
      for (int i = 0; i < n_local; i += 2) {
        const bool two = i + 1 < n_local;
        stage(0, 0);
        if (two) stage(1, 0);
        stage(0, 1);
        if (two) stage(1, 1);
      }
    }
  } else {
    // ---------------------------------------------------------------- workers: two groups of 4 warps (128 threads)
    const int g = (warp - 2) >> 2;         // group: handles this CTA's tiles g, g + 2, g + 4, ...
    const int q = warp & 3;                // TMEM lane quarter == frame pair handled in the TMEM phases
    const int wt = ((warp - 2) & 3) * 32 + lane;  // 0..127 inside the group
    uint8_t* s_ahi = s_a + g * 2 * LM_OP_B;
    uint8_t* s_alo = s_ahi + LM_OP_B;
    float* s_pg = s_p + g * LM_FR * LM_PSTRIDE;
    const uint32_t trow = tm + ((uint32_t)(q * 32) << 16) + g * 128;
    const uint32_t bar_id = 1 + g;
    uint32_t dph = 0;
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tc0 = clock64();
#define LM_TICK(k) do { if (p.dbg) { const long long c_ = clock64(); tph[k] += c_ - tc0; tc0 = c_; } } while (0)
    for (int i = g; i < n_local; i += 2) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int buf = i % LM_NRAW;
      const int b = tile / p.tiles_per_item, t0 = (tile - b * p.tiles_per_item) * LM_FR;
      // ---- pre-pass: window, split, lay out as the stage-1 A operand (MN-major fp16)
      mbar_wait(&raw_full[buf], (uint32_t)(i / LM_NRAW) & 1u);
      LM_TICK(0);
      const uint8_t* raw = s_raw + buf * LM_RAW_B;
#pragma unroll 4
      for (int it = 0; it < 16; ++it) {
        const int u = wt + 128 * it;              // 16-byte unit of the raw tile: [frame][n1][n2 / 4]
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
        *reinterpret_cast<uint2*>(s_ahi + off) = make_uint2(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[2]));
        *reinterpret_cast<uint2*>(s_alo + off) = make_uint2(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[2]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&raw_empty[buf]);
        mbar_arrive(&work_ready[g]);
      }
      LM_TICK(1);
      // ---- stage-1 result: twiddle, split, transpose into the stage-2 A operand (K-major fp16)
      mbar_wait(&mma_done[g], dph);
      dph ^= 1u;
      tc_fence_after();
      LM_TICK(2);
      {
        uint32_t vr[32], vi[32];
        tmem_ld32(trow, vr);
        tmem_ld32(trow + 32, vi);
        tmem_ld_wait();
        const int n2 = lane;
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) {
          const float ct = s_tw[k1 * 32 + n2], st = s_tw[1024 + k1 * 32 + n2];
          const float re = __uint_as_float(vr[k1]), im = __uint_as_float(vi[k1]);
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
      if (lane == 0) mbar_arrive(&work_ready[g]);
      LM_TICK(3);
      // ---- stage-2 result: unpack the two real spectra of the pair, power
      mbar_wait(&mma_done[g], dph);
      dph ^= 1u;
      tc_fence_after();
      LM_TICK(4);
      {
        uint32_t ur[32], ui[32];
        tmem_ld32(trow + 64, ur);
        tmem_ld32(trow + 96, ui);
        tmem_ld_wait();
        const int src = (32 - lane) & 31;
        float* pa = s_pg + (2 * q) * LM_PSTRIDE;
        float* pb = pa + LM_PSTRIDE;
#pragma unroll
        for (int k2 = 0; k2 <= 16; ++k2) {
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
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      LM_TICK(5);
      // ---- banded mel filterbank, log, normalise, store: one work item per (mel filter, group of 3 frames)
      const int crop = p.crop ? p.crop[b] : 0;
      for (int o = wt; o < 3 * p.n_mels; o += 128) {
        const int g3 = o / p.n_mels, m = o - g3 * p.n_mels;
        const int f0 = 3 * g3, nf = min(3, LM_FR - f0);
        const float* pp = s_pg + f0 * LM_PSTRIDE + s_meli[m];
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
        for (int fi = 0; fi < 3; ++fi) {
          const int t = t0 + f0 + fi, t_out = t - crop;
          if (fi >= nf || t >= p.T || t_out < 0 || t_out >= p.T_out) continue;
          const float y = (__logf(1e-5f + acc[fi]) + 4.0f) * 0.25f;
          if (p.out_bmt) p.out_bmt[((size_t)b * p.n_mels + m) * p.T_out + t_out] = y;
          if (p.out_btm) p.out_btm[((size_t)b * p.T_out + t_out) * p.n_mels + m] = y;
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      LM_TICK(6);
    }
    if (p.dbg && wt == 0 && g == 0)
      for (int k = 0; k < 7; ++k) p.dbg[blockIdx.x * 8 + k] = tph[k];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tm, 256);
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
  if (n_fft != LM_NFFT || hop <= 0 || (hop % 4) || n_mels <= 0 || n_mels > 128 || mel_nnz <= 0 || mel_nnz > 2048)
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
  const size_t smem = LM_NRAW * LM_RAW_B + 4 * LM_OP_B + 3 * 8192 + (2048 + LM_NFFT + 2 * LM_FR * LM_PSTRIDE + 2048) * 4 +
                      384 * 4 + 16 * 8 + 1024;
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
