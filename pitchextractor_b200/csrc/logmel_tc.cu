// log-mel front-end on tcgen05 tensor cores (n_fft = 1024): the 1024-point real DFT of TWO frames is one complex FFT
// (z = x_a + i x_b), factorised 32 x 32 (four-step): two 32-point DFT stages, each a [128 x 64] x [64 x 64] real GEMM
// over 4 frame pairs, with the 1024-th-root twiddles applied between them.  fp32-grade accuracy comes from a two-way
// fp16 split of both operands (hi*hi + hi*lo + lo*hi), accumulated in fp32 in TMEM: 24 tcgen05.mma (kind::f16, 128x64x16)
// per 8 frames instead of the 6.3 MFLOP/frame of the dense DFT.  The data split is conversion-light (the conversion pipe is
// the scarcest one here): hi = the value with its low 13 mantissa bits cleared (exactly representable in fp16), lo = the
// exact fp32 remainder; the data are pre-scaled by powers of two (window x 2^11, twiddles x 2^-5, undone in the mel
// weights) so that the remainders stay in fp16's normal range without a per-element rescale.
//
// One CTA per SM, persistent.  FOUR independent worker groups (4 warps each) work on four 8-frame slots at a time, so
// that the SIMT phases of three groups overlap the tensor-core stage and the waits of the fourth:
//   * every slot's samples are read from HBM ONCE: its 8 overlapping frames span 1024 + 7*hop samples, fetched with one
//     bulk async copy (cp.async.bulk, 12.5 KB at hop 300) into the group's raw buffer; the reflect padding of the few
//     frames at the edges of an item is gathered straight from global memory instead.
//   * pre-pass: window (held in registers) x sample -> fp16 hi / lo' -> MN-major stage-1 operand.
//   * stage-1 result: twiddle, split, stored as the MN-major stage-2 operand: a thread owns one (pair, n2) accumulator row
//     and writes 16-byte runs of 8 consecutive k1 -- no transposition through shared memory.
//   * stage-2 result: pair unpacking by warp shuffles, |.|^2 into the (now free) operand buffer, then the banded mel
//     filterbank with the 80 filters dealt to 16 threads per frame in a load-balanced order, log, normalise, store.
// Only the frames that are asked for are computed (crop .. crop + T_out), frames past an item's own end are written as
// zeros (Collater padding, meldataset.py:804-816).  Mirrors torchaudio MelSpectrogram as built at reference
// meldataset.py:77 and the normalisation at meldataset.py:650.
#include "common.cuh"
#include "../../include/pitchextractor_b200.h"
#include <cuda_fp16.h>

namespace pe {

constexpr int LM_NFFT = 1024;
constexpr int LM_FR = 8;                        // frames per slot (4 pairs)
constexpr int LM_GROUPS = 4;                    // worker groups == slots in flight
constexpr int LM_WORKERS = LM_GROUPS * 128;
constexpr int LM_THREADS = 64 + LM_WORKERS;     // warp 0 producer, warp 1 MMA, warps 2..17 workers
constexpr int LM_MAX_HOP = 320;
constexpr int LM_RAW_B = (LM_NFFT + (LM_FR - 1) * LM_MAX_HOP) * 4;   // 13 056 B of fp32 samples per slot
constexpr int LM_OP_B = 16384;                  // one fp16 [128 x 64] operand
constexpr int LM_SLOT_B = 2 * LM_OP_B;          // A_hi | A_lo; later the slot's power spectra [513 bins][8 frames]
constexpr int LM_PROW_B = 48;                   // bytes per power-spectrum row (8 frames + pad: 3 x 16 B spreads the banks)
constexpr int LM_MEL_W = 1536;                  // banded filter weights held in shared memory
constexpr int LM_ITEMS = 128;                   // mel work items (one per worker thread of a group)

struct LmParams {
  int B, T_out, n_mels, tiles_per_item, num_tiles;
  const float* win;        // [1024], x 2^11
  const __half* fmat;      // 2 x 8 KB pre-swizzled K-major [64 x 64]: hi, lo
  const float* tw;         // [2][32][32]: cos, sin of 2 pi k1 n2 / 1024, indexed [k1][n2], x 2^-5
  const int4* mel_items;   // [128] {filter m (-1: idle), first bin, bins, offset into mel_w | flags << 24}
  const float* mel_w;      // banded filter weights, x 2^-14 (the power spectra carry (2 * 2^6)^2)
  int mel_nnz;
  const int* crop;         // [B] or NULL
  const int* lengths;      // [B] valid samples per item or NULL (= L)
  float* out_bmt;          // [B][n_mels][T_out] or NULL
  float* out_btm;          // [B][T_out][n_mels] or NULL
};
constexpr int LM_ITEM_COMBINE = 1;  // the filter is split over lanes l, l ^ 1: add the partner's partial sums
constexpr int LM_ITEM_WRITER = 2;   // this lane applies the log and stores the filter's 8 frames

__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  tc_mma_bf16(tmem_d, da, db, idesc, acc);  // same instruction (kind::f16); operand formats live in idesc
}

// ---- explicit shared-state-space accesses on 32-bit addresses (the compiler otherwise falls back to generic LD / ST with
// 64-bit address arithmetic for pointers carved out of the dynamic shared-memory block)
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 lds128i(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts64f(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// non-blocking probe (mbarrier.try_wait suspends the thread for a hardware time slice when the phase is not complete:
// useless for a poller that watches several barriers)
__device__ __forceinline__ uint32_t mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// wait with a long hardware suspend per probe: the workers wait for whole pipeline phases, and every probe that returns
// early costs an issue slot the other 17 warps of the SM could use
__device__ __forceinline__ void mbar_wait_long(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if (!ok && ++spins > (1u << 22)) mbar_timeout_trap();
  } while (!ok);
}

// x0, x1 -> packed fp16 hi pair (mantissa truncated to fp16's 10 bits: the conversion is exact) and packed fp16 lo pair
// (the exact remainder x - hi, rounded to fp16): x = hi + lo to ~2^-21 relative
__device__ __forceinline__ void split16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const float h0 = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u);
  const float h1 = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
  const __half2 h = __floats2half2_rn(h0, h1);
  const __half2 l = __floats2half2_rn(x0 - h0, x1 - h1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct LmSlot {        // where a slot's frames lie (identical arithmetic in the producer and the workers)
  int b, t_out0, t0, nfr, len;
  long long s_base;    // sample index of raw[0]
  long long c_lo, c_hi;  // sample range covered by the bulk copy ([c_lo, c_hi), multiples of 4; empty if c_hi <= c_lo)
};

__device__ __forceinline__ LmSlot lm_slot(const LmParams& p, int tile, int L, int hop) {
  LmSlot s;
  s.b = tile / p.tiles_per_item;
  s.t_out0 = (tile - s.b * p.tiles_per_item) * LM_FR;
  s.t0 = (p.crop ? p.crop[s.b] : 0) + s.t_out0;
  s.len = p.lengths ? min(max(p.lengths[s.b], 0), L) : L;
  const int T_b = s.len > LM_NFFT / 2 ? 1 + s.len / hop : 0;   // reflect padding needs pad < length (torch.stft)
  s.nfr = max(0, min(min(LM_FR, p.T_out - s.t_out0), T_b - s.t0));
  s.s_base = (long long)s.t0 * hop - LM_NFFT / 2;
  s.c_lo = s.s_base > 0 ? s.s_base : 0;
  const long long end = s.s_base + LM_NFFT + (long long)(LM_FR - 1) * hop;
  const long long len4 = s.len & ~3;
  s.c_hi = s.nfr > 0 ? (end < len4 ? end : len4) : 0;
  return s;
}

__global__ void __launch_bounds__(LM_THREADS, 1)
logmel_tc_kernel(const float* __restrict__ wave, int L, long long ld, int hop, const LmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // all shared-memory traffic of the workers goes through 32-bit shared-window addresses
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t a_slots = sbase;                                    // LM_GROUPS x 32 KB operand / power buffers
  const uint32_t a_f = a_slots + LM_GROUPS * LM_SLOT_B;              // 16 KB: F_hi | F_lo
  const uint32_t a_raw = a_f + 2 * 8192;                             // LM_GROUPS x raw samples
  const uint32_t a_tw = a_raw + LM_GROUPS * LM_RAW_B;                // 8 KB twiddles
  const uint32_t a_melw = a_tw + 2048 * 4;                           // banded mel weights
  const uint32_t a_items = a_melw + LM_MEL_W * 4;                    // [128] int4
  const uint32_t a_bars = a_items + LM_ITEMS * 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (a_bars - sbase));
  uint64_t* raw_full = bars;                      // [G] producer -> workers
  uint64_t* raw_empty = bars + LM_GROUPS;         // [G] workers -> producer
  uint64_t* work_ready = bars + 2 * LM_GROUPS;    // [G] workers of group g -> MMA
  uint64_t* mma_done = bars + 3 * LM_GROUPS;      // [G] MMA -> workers of group g
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * LM_GROUPS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    uint4* df = reinterpret_cast<uint4*>(smem + (a_f - sbase));
    for (int i = tid; i < 2 * 8192 / 16; i += LM_THREADS) df[i] = __ldg(reinterpret_cast<const uint4*>(p.fmat) + i);
    float* dt = reinterpret_cast<float*>(smem + (a_tw - sbase));
    for (int i = tid; i < 2048; i += LM_THREADS) dt[i] = p.tw[i];
    float* dw = reinterpret_cast<float*>(smem + (a_melw - sbase));
    for (int i = tid; i < p.mel_nnz; i += LM_THREADS) dw[i] = p.mel_w[i];
    int4* di = reinterpret_cast<int4*>(smem + (a_items - sbase));
    for (int i = tid; i < LM_ITEMS; i += LM_THREADS) di[i] = p.mel_items[i];
  }
  if (tid == 0) {
    for (int g = 0; g < LM_GROUPS; ++g) {
      mbar_init(&raw_full[g], 1);
      mbar_init(&raw_empty[g], 4);
      mbar_init(&work_ready[g], 4);
      mbar_init(&mma_done[g], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const int n_local = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // slots of this CTA

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: one bulk copy per slot
    for (int i = 0; i < n_local; ++i) {
      const int g = i & (LM_GROUPS - 1);
      const LmSlot s = lm_slot(p, blockIdx.x + i * gridDim.x, L, hop);
      if (elect_one()) {
        mbar_wait_long(&raw_empty[g], ((uint32_t)(i / LM_GROUPS) & 1u) ^ 1u);
        if (s.c_hi > s.c_lo) {
          const uint32_t bytes = (uint32_t)(s.c_hi - s.c_lo) * 4u;
          mbar_arrive_expect_tx(&raw_full[g], bytes);
          bulk_load_1d(smem + (a_raw - sbase) + g * LM_RAW_B + (s.c_lo - s.s_base) * 4,
                       wave + (long long)s.b * ld + s.c_lo, bytes, &raw_full[g]);
        } else {
          mbar_arrive(&raw_full[g]);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer: serves whichever group has an operand
    // ready (each group alternates stage 1 / stage 2 of its slots): a fixed order would couple the four groups and make
    // each wait for the slowest one twice per slot
    constexpr uint32_t IDESC = umma_idesc(UMMA_F16, 128, 64, 1, 0);  // A MN-major, B K-major, both stages
    if (lane == 0) {
      uint32_t ph = 0;        // bit g: parity of group g's next work_ready phase
      uint32_t which = 0;     // bit g: next stage of group g
      int left[LM_GROUPS];    // stages still to issue for group g
      int total = 0;
#pragma unroll
      for (int g = 0; g < LM_GROUPS; ++g) {
        left[g] = 2 * ((n_local - g + LM_GROUPS - 1) / LM_GROUPS);
        total += left[g];
      }
      uint32_t idle = 0;
      while (total > 0) {
        bool any = false;
#pragma unroll
        for (int g = 0; g < LM_GROUPS; ++g) {
          if (left[g] > 0 && mbar_test(&work_ready[g], (ph >> g) & 1u)) {
            tc_fence_after();
            const uint32_t ahi = a_slots + g * LM_SLOT_B, alo = ahi + LM_OP_B;
            const uint32_t d = tm + g * 128 + ((which >> g) & 1u) * 64;
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
              const uint32_t a = prod == 2 ? alo : ahi;
              const uint32_t bm = a_f + (prod == 1 ? 8192 : 0);   // hi*hi, hi*lo, lo*hi
#pragma unroll
              for (int k = 0; k < 4; ++k)  // A: 2 MN atoms (8 KB apart) x 64 K rows, 16 K rows per MMA
                tc_mma_f16(d, umma_desc_sw128(a + k * 2048, 8192, 1024), umma_desc_sw128(bm + k * 32, 16, 1024), IDESC,
                           (prod > 0 || k > 0) ? 1u : 0u);
            }
            tc_commit(&mma_done[g]);
            ph ^= 1u << g;
            which ^= 1u << g;
            --left[g];
            --total;
            any = true;
          }
        }
        if (any) idle = 0;
        else if (++idle > (1u << 24)) mbar_timeout_trap();
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- workers: LM_GROUPS groups of 4 warps
    const int g = (warp - 2) >> 2;         // group: handles this CTA's slots g, g + 4, g + 8, ...
    const int q = warp & 3;                // TMEM lane quarter == frame pair handled in the TMEM phases
    const int wt = ((warp - 2) & 3) * 32 + lane;  // 0..127 inside the group
    const uint32_t a_hi = a_slots + g * LM_SLOT_B, a_lo = a_hi + LM_OP_B;
    const uint32_t a_p = a_hi;             // power spectra [513][8] reuse the operand buffer after stage 2
    const uint32_t a_rawg = a_raw + g * LM_RAW_B;
    const uint32_t trow = tm + ((uint32_t)(q * 32) << 16) + g * 128;
    const uint32_t bar_id = 1 + g;
    uint32_t dph = 0;
    // this thread's two window positions (two 16-byte units of a frame): fixed for every frame
    float4 wreg[2];
    uint32_t spos[2];      // byte offset of the position inside a frame
    uint32_t soff[2][2];   // operand byte offset for (frame parity, position) without the pair terms
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // (n1, n2) of the thread's 16-byte unit.  The four n1 rows of a warp are {b, b+1, b+4, b+5}: their operand rows
      // then split 2 + 2 over the two 64-byte halves of the swizzled 128-byte row, the minimum for a 256-byte store
      // (four consecutive rows would all land in the same half: 4-way bank conflicts)
      const int r4 = (wt >> 3) & 3, w4 = wt >> 5;
      const int n1 = (r4 & 1) + 4 * (r4 >> 1) + 2 * (w4 & 1) + 8 * ((w4 >> 1) + 2 * h), n2 = (wt & 7) << 2;
      wreg[h] = __ldg(reinterpret_cast<const float4*>(p.win + n1 * 32 + n2));
      spos[h] = (uint32_t)(n1 * 32 + n2) * 4u;
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int kk = par * 32 + n1;
        soff[par][h] = (uint32_t)(kk * 128 + (n2 & 7) * 2) | ((uint32_t)((n2 >> 3) ^ (kk & 7)) << 16);  // row | unit-xor
      }
    }
    const int4 item = lds128i(a_items + wt * 16);
    for (int i = g; i < n_local; i += LM_GROUPS) {
      const LmSlot s = lm_slot(p, blockIdx.x + i * gridDim.x, L, hop);
      const long long span_end = s.s_base + LM_NFFT + (long long)(LM_FR - 1) * hop;
      const bool all_fast = s.nfr == LM_FR && s.s_base >= s.c_lo && span_end <= s.c_hi;   // slot interior to the item
      // ---- pre-pass: window, split, lay out as the stage-1 A operand (MN-major fp16)
      mbar_wait_long(&raw_full[g], (uint32_t)(i / LM_GROUPS) & 1u);
      auto emit = [&](int fr, int h, const float4 x) {
        uint32_t h01, l01, h23, l23;
        split16x2(x.x * wreg[h].x, x.y * wreg[h].y, h01, l01);
        split16x2(x.z * wreg[h].z, x.w * wreg[h].w, h23, l23);
        const int pr = fr >> 1;
        const uint32_t so = soff[fr & 1][h];
        // unit = (pr & 1) * 4 + (n2 >> 3), xor (kk & 7): the pair bit only flips bit 2 of the 16-byte unit index
        const uint32_t off = (uint32_t)((pr >> 1) * 8192) + (so & 0xFFFFu) + ((((so >> 16) ^ (uint32_t)((pr & 1) * 4))) << 4);
        sts64(a_hi + off, h01, h23);
        sts64(a_lo + off, l01, l23);
      };
      if (all_fast) {
        uint32_t fo = 0;
#pragma unroll
        for (int fr = 0; fr < LM_FR; ++fr) {
          const float4 x0 = lds128(a_rawg + fo + spos[0]);
          const float4 x1 = lds128(a_rawg + fo + spos[1]);
          emit(fr, 0, x0);
          emit(fr, 1, x1);
          fo += (uint32_t)hop * 4u;
        }
      } else {
#pragma unroll 1
        for (int fr = 0; fr < LM_FR; ++fr) {
          const long long s0 = s.s_base + (long long)fr * hop;
          const bool live = fr < s.nfr;
          const bool fast = live && s0 >= s.c_lo && s0 + LM_NFFT <= s.c_hi;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);     // frames past the end of the item contribute zeros
            if (fast) {
              x = lds128(a_rawg + (uint32_t)(fr * hop) * 4u + spos[h]);
            } else if (live) {  // reflect padding (torch.stft center=True, pad_mode="reflect") at the item's own ends
              const float* wb = wave + (long long)s.b * ld;
              float e[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                long long si = s0 + (spos[h] >> 2) + j;
                if (si < 0) si = -si;
                if (si >= s.len) si = 2LL * (s.len - 1) - si;
                si = si < 0 ? 0 : (si >= s.len ? s.len - 1 : si);
                e[j] = __ldg(wb + si);
              }
              x = make_float4(e[0], e[1], e[2], e[3]);
            }
            emit(fr, h, x);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&raw_empty[g]);
        mbar_arrive(&work_ready[g]);
      }
      // ---- stage-1 result: twiddle, split, store as the stage-2 A operand (MN-major: K row = (re/im, n2), this
      // thread's 32 k1 values of a row are 64 contiguous bytes)
      mbar_wait(&mma_done[g], dph);
      dph ^= 1u;
      tc_fence_after();
      {
        const int n2 = lane;
        const uint32_t atom = (uint32_t)(q >> 1) * 8192u;
        const uint32_t row_re = atom + (uint32_t)n2 * 128u, row_im = row_re + 32u * 128u;
        const uint32_t sw = (uint32_t)(n2 & 7);      // (32 + n2) & 7 == n2 & 7
        const uint32_t a_twl = a_tw + (uint32_t)n2 * 4u;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint32_t vr[8], vi[8];
          tmem_ld8(trow + c8 * 8, vr);
          tmem_ld8(trow + 32 + c8 * 8, vi);
          tmem_ld_wait();
          uint32_t rh[4], rl[4], ih[4], il[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float yr[2], yi[2];
#pragma unroll
            for (int d = 0; d < 2; ++d) {
              const int k1 = c8 * 8 + e * 2 + d;
              const float ct = lds32(a_twl + k1 * 128), st = lds32(a_twl + 4096 + k1 * 128);
              const float re = __uint_as_float(vr[e * 2 + d]), im = __uint_as_float(vi[e * 2 + d]);
              yr[d] = fmaf(re, ct, im * st);
              yi[d] = fmaf(im, ct, -re * st);
            }
            split16x2(yr[0], yr[1], rh[e], rl[e]);
            split16x2(yi[0], yi[1], ih[e], il[e]);
          }
          const uint32_t o = (((uint32_t)((q & 1) * 4 + c8)) ^ sw) << 4;
          sts128(a_hi + row_re + o, rh[0], rh[1], rh[2], rh[3]);
          sts128(a_lo + row_re + o, rl[0], rl[1], rl[2], rl[3]);
          sts128(a_hi + row_im + o, ih[0], ih[1], ih[2], ih[3]);
          sts128(a_lo + row_im + o, il[0], il[1], il[2], il[3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&work_ready[g]);
      // ---- stage-2 result: unpack the two real spectra of the pair, power -> P[bin][frame] in the free operand buffer.
      // Bin k = k1 + 32 k2 (k1 = lane) pairs with N - k = (32 - k1) % 32 + 32 (31 - k2 [+1 if k1 == 0]): own columns ascend,
      // the partner's descend, so the accumulator row is read in two halves of 16 + 16 columns
      mbar_wait(&mma_done[g], dph);
      dph ^= 1u;
      tc_fence_after();
      {
        const int src = (32 - lane) & 31;
        const uint32_t a_pk = a_p + (uint32_t)lane * LM_PROW_B + (uint32_t)q * 8u;   // + k2 * 32 rows
        auto unpack = [&](int k2, float zr, float zi, float sr, float si) {
          const float pr_ = __shfl_sync(0xffffffffu, sr, src);
          const float pi_ = __shfl_sync(0xffffffffu, si, src);
          if (k2 < 16 || lane == 0) {
            const float ar = zr + pr_, ai = zi - pi_, br = zi + pi_, bi = pr_ - zr;
            sts64f(a_pk + (uint32_t)k2 * (32u * LM_PROW_B), fmaf(ar, ar, ai * ai), fmaf(br, br, bi * bi));
          }
        };
        {  // k2 = 0..7: own columns 0..7; partner columns 31..24 (lane 0: 0, 31..25)
          uint32_t ur[8], ui[8], tr[8], ti[8];
          tmem_ld8(trow + 64, ur);
          tmem_ld8(trow + 96, ui);
          tmem_ld8(trow + 64 + 24, tr);
          tmem_ld8(trow + 96 + 24, ti);
          tmem_ld_wait();
#pragma unroll
          for (int k2 = 0; k2 < 8; ++k2) {
            // lane 0 sends column (32 - k2) & 31, the others 31 - k2
            const uint32_t o_r = tr[7 - k2], o_i = ti[7 - k2];
            const uint32_t s_r = k2 == 0 ? ur[0] : tr[8 - k2 > 7 ? 7 : 8 - k2], s_i = k2 == 0 ? ui[0] : ti[8 - k2 > 7 ? 7 : 8 - k2];
            unpack(k2, __uint_as_float(ur[k2]), __uint_as_float(ui[k2]),
                   __uint_as_float(lane == 0 ? s_r : o_r), __uint_as_float(lane == 0 ? s_i : o_i));
          }
        }
        {  // k2 = 8..16: own columns 8..16; partner columns 23..15 (lane 0: 24..16)
          uint32_t ur[16], ui[16], t24r[8], t24i[8];
          tmem_ld16(trow + 64 + 8, ur);    // columns 8..23
          tmem_ld16(trow + 96 + 8, ui);
          tmem_ld8(trow + 64 + 24, t24r);  // column 24 (lane 0, k2 = 8)
          tmem_ld8(trow + 96 + 24, t24i);
          tmem_ld_wait();
#pragma unroll
          for (int k2 = 8; k2 <= 16; ++k2) {
            const int co = 31 - k2 - 8, cs = 32 - k2 - 8;  // indices into ur / ui (columns 8..23)
            const uint32_t o_r = ur[co < 0 ? 0 : co], o_i = ui[co < 0 ? 0 : co];
            const uint32_t s_r = cs > 15 ? t24r[0] : ur[cs], s_i = cs > 15 ? t24i[0] : ui[cs];
            unpack(k2, __uint_as_float(ur[k2 - 8]), __uint_as_float(ui[k2 - 8]),
                   __uint_as_float(lane == 0 ? s_r : o_r), __uint_as_float(lane == 0 ? s_i : o_i));
          }
        }
      }
      tc_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      // ---- banded mel filterbank: thread = one work item (a filter or half of a long one) x the slot's 8 frames;
      // log, normalise, store
      {
        const int m = item.x, cnt = item.z;
        const uint32_t flags = (uint32_t)item.w >> 24;
        float acc[LM_FR];
#pragma unroll
        for (int f = 0; f < LM_FR; ++f) acc[f] = 0.f;
        uint32_t ap = a_p + (uint32_t)item.y * LM_PROW_B;
        uint32_t aw = a_melw + ((uint32_t)item.w & 0xFFFFFFu) * 4u;
#pragma unroll 2
        for (int c = 0; c < cnt; ++c) {
          const float w = lds32(aw);
          const float4 p0 = lds128(ap), p1 = lds128(ap + 16);
          acc[0] = fmaf(p0.x, w, acc[0]); acc[1] = fmaf(p0.y, w, acc[1]);
          acc[2] = fmaf(p0.z, w, acc[2]); acc[3] = fmaf(p0.w, w, acc[3]);
          acc[4] = fmaf(p1.x, w, acc[4]); acc[5] = fmaf(p1.y, w, acc[5]);
          acc[6] = fmaf(p1.z, w, acc[6]); acc[7] = fmaf(p1.w, w, acc[7]);
          ap += LM_PROW_B;
          aw += 4;
        }
#pragma unroll
        for (int f = 0; f < LM_FR; ++f) {
          const float o = __shfl_xor_sync(0xffffffffu, acc[f], 1);
          if (flags & LM_ITEM_COMBINE) acc[f] += o;
        }
        if ((flags & LM_ITEM_WRITER) && m >= 0) {
          float y[LM_FR];
#pragma unroll
          for (int f = 0; f < LM_FR; ++f) y[f] = f < s.nfr ? fmaf(__logf(1e-5f + acc[f]), 0.25f, 1.0f) : 0.f;
          const int nw = min(LM_FR, p.T_out - s.t_out0);   // frames of this slot that exist in the output
          if (p.out_bmt) {
            float* o = p.out_bmt + ((size_t)s.b * p.n_mels + m) * p.T_out + s.t_out0;
            if (nw == LM_FR && (p.T_out & 3) == 0) {       // 8 consecutive frames of one filter: two 16-byte stores
              reinterpret_cast<float4*>(o)[0] = make_float4(y[0], y[1], y[2], y[3]);
              reinterpret_cast<float4*>(o)[1] = make_float4(y[4], y[5], y[6], y[7]);
            } else {
#pragma unroll
              for (int f = 0; f < LM_FR; ++f)
                if (f < nw) o[f] = y[f];
            }
          }
          if (p.out_btm) {
            float* o = p.out_btm + ((size_t)s.b * p.T_out + s.t_out0) * p.n_mels + m;
#pragma unroll
            for (int f = 0; f < LM_FR; ++f)
              if (f < nw) o[(size_t)f * p.n_mels] = y[f];
          }
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tm, 512);
}

constexpr size_t LM_SMEM = (size_t)LM_GROUPS * LM_SLOT_B + 2 * 8192 + (size_t)LM_GROUPS * LM_RAW_B +
                           (2048 + LM_MEL_W) * 4 + LM_ITEMS * 16 + (4 * LM_GROUPS + 2) * 8 + 1024;

}  // namespace pe

extern "C" int pe_logmel_tc(const float* wave, int B, int L, int n_fft, int hop, int n_mels, const float* win,
                            const void* fmat, const float* tw, const int* mel_items, const float* mel_w, int mel_nnz,
                            float* xpad, size_t xpad_bytes, float* out_bmt, float* out_btm, const int* crop,
                            const int* lengths, int T_out, pe_stream_t stream) {
  using namespace pe;
  if (int rc = pe_host::check_arch()) return rc;
  if (!wave || !win || !fmat || !tw || !mel_items || !mel_w || B <= 0)
    return PE_ERR_BAD_SHAPE;
  if (n_fft != LM_NFFT || hop <= 0 || (hop % 4) || hop > LM_MAX_HOP || n_mels <= 0 || n_mels > 128 || mel_nnz <= 0 ||
      mel_nnz > LM_MEL_W)
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
  p.B = B; p.n_mels = n_mels; p.T_out = T_out;
  p.tiles_per_item = (T_out + LM_FR - 1) / LM_FR;
  p.num_tiles = B * p.tiles_per_item;
  p.win = win; p.fmat = (const __half*)fmat; p.tw = tw;
  p.mel_items = reinterpret_cast<const int4*>(mel_items); p.mel_w = mel_w; p.mel_nnz = mel_nnz;
  p.crop = crop; p.lengths = lengths; p.out_bmt = out_bmt; p.out_btm = out_btm;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(logmel_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LM_SMEM) != cudaSuccess)
      return PE_ERR_LAUNCH;
    attr = true;
  }
  const int want = (p.num_tiles + LM_GROUPS - 1) / LM_GROUPS;   // every CTA should have a slot for each of its groups
  const int grid = want < pe_host::num_sms() ? want : pe_host::num_sms();
  logmel_tc_kernel<<<grid, LM_THREADS, LM_SMEM, st>>>(src, L, ld, hop, p);
  return cudaGetLastError() == cudaSuccess ? PE_OK : PE_ERR_LAUNCH;
}
