// tcgen05 / TMEM / TMA tile engine: one persistent warp-specialised kernel (compiled per mode and per k-blocks-per-stage)
//   mode 0  plain GEMM          D[M,N] = sum_k A(m,k) B(n,k)            (linear fwd / dgrad / wgrad)
//   mode 1  3x3 conv (+1x1)     implicit GEMM over NHWC pixels, taps walked by shifted 4-D TMA boxes
//   mode 2  conv weight grad    contraction over pixels, both operands MN-major straight from NHWC
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue (TMEM -> registers ->
// fused bias / GELU / dropout / residual / column statistics -> swizzled staging block -> TMA store or reduce-add).
// Producer and MMA warps run warp-uniform loops in which one elected lane issues.  Accumulators live in TMEM (two
// stages), operands in a 192 KB shared-memory ring shared by consecutive tiles.
#include "common.cuh"
#include "../../include/pitchextractor_b200.h"
#include <cstdlib>

PE_USES_STEP_SALT()

namespace pe {

constexpr int kBlockM = 128;
constexpr int kRowBytes = 128;                    // one swizzle-128B row: 64 bf16 (or 32 tf32) along the contiguous dim
constexpr int kATileBytes = kBlockM * kRowBytes;  // 16 KB
constexpr int kNumEpiWarps = 8;
constexpr int kStagingBytes = 2 * 32 * 64;  // per epilogue warp: two 32-row x 64-byte blocks (64-byte swizzled)
constexpr int kNumThreads = 64 + 32 * kNumEpiWarps;  // warp 0 = TMA, warp 1 = MMA, warps 2..9 = epilogue
constexpr int kNumThreadsWgrad = kNumThreads + 32;   // mode 2: warp 10 = second TMA producer (operand B boxes)

struct TcParams {
  int mode;  // 0 gemm, 1 conv fwd, 2 conv wgrad
  int kind;  // 0 bf16, 1 tf32
  int a_mn, b_mn;
  int M, N;        // output extents used for masking
  int block_n;     // columns per CTA tile (multiple of 16, <= 256)
  int tmem_cols;   // power of two >= 2 * block_n (two accumulator stages)
  int acc_stride;  // TMEM column offset of accumulator stage 1
  int stages;
  int kb_total, kb_per_split;
  int a_boxes, b_boxes;  // number of 64-wide TMA boxes for MN-major operands
  int tx_bytes;          // bytes the TMA loads of one k-block deliver (mbarrier expect_tx)
  int tiles_x, tiles_y, tiles_z, num_tiles;  // persistent tile space (x fastest)
  // conv geometry
  int H, W, tw, th, tiles_w, tiles_h;
  int c1_chunks, c2_chunks;
  int taps;  // wgrad: 9 or 1
  unsigned a_inc, b_inc;  // descriptor start-address step (>> 4) between the four MMAs of a 64-deep k-block (set by launch_tc)
  int kps;           // 64-wide k-blocks per smem stage (1 or 2): narrow tiles amortise the barrier round trip over two
  int commit_group;  // 1, 2 or 4: k-blocks per tcgen05.commit on the smem ring (p.stages is a multiple of it)
  int out_tma;  // bit 0 / 1: ep.out / ep.out2 are written with TMA tensor stores from the staging blocks
  long long* dbg;  // optional [gridDim.x][8] counters (tuning): mma loop cycles, mma wait-full, mma wait-tmem, tma wait-empty,
                   // entry globaltimer ns, cycles entry->setup done, entry->mma loop end, entry->CTA end
  pe_epilogue ep;
};

__device__ __forceinline__ void decode_conv_tile(const TcParams& p, int tile, int& b, int& h0, int& w0) {
  const int per_img = p.tiles_h * p.tiles_w;
  b = tile / per_img;
  const int r = tile - b * per_img;
  h0 = (r / p.tiles_w) * p.th;
  w0 = (r % p.tiles_w) * p.tw;
}

struct TileCoord {
  int tx, ty, tz;
};
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int tile) {
  TileCoord c;
  c.tx = tile % p.tiles_x;
  const int r = tile / p.tiles_x;
  c.ty = r % p.tiles_y;
  c.tz = r / p.tiles_y;
  return c;
}

// Persistent, warp-specialised tile engine: every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the smem
// operand ring and the two TMEM accumulator stages are shared by consecutive tiles, so the epilogue of tile i overlaps
// the TMA + MMA main loop of tile i + 1.
// MODE and KPS are compile-time copies of p.mode and p.kps: the producer / MMA warps are single instruction streams whose per-k-block
// latency bounds narrow tiles, so their loops must not carry the other modes' branches.
// PAIR = 1: the kernel runs as clusters of two CTAs and every tile is 256 rows tall (tcgen05 cta_group::2): CTA r of the
// pair loads its own 128 rows of A and half of the B tile and keeps its 128 accumulator rows in its own TMEM; only the
// leader (rank 0) issues MMAs.  Each SM then fills / reads half of B per MMA, which is what the shared-memory port of a
// single SM cannot sustain at full tensor rate (operand reads + TMA fills of a 128 x 256 x 64 step: 96 KB per 512 cycles).
// FEAT selects what the epilogue is compiled with: bit 0 = column statistics (BatchNorm batch statistics / BatchNorm-backward
// sums), bit 1 = the "heavy" element-wise paths (GELU and its derivative, Philox dropout, second output).  An epilogue that
// carries code it does not run is measurably slower (registers, instruction cache): with the statistics paths compiled in,
// FFN1's epilogue took 60 us instead of 52.
template <int MODE, int KPS, int PAIR, int FEAT>
__global__ void __launch_bounds__(kNumThreadsWgrad, 1)
tc_tile_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_a2,
               const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_out,
               const __grid_constant__ CUtensorMap tma_out2, const TcParams p) {
  // No static shared memory in this kernel: the dynamic window then starts 1024-byte aligned (which the 128-byte
  // swizzle of the operand tiles needs), and all 227 KB are usable.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  constexpr bool STATS = (FEAT & 1) != 0, HEAVY = (FEAT & 2) != 0;
  constexpr int NC = PAIR ? 2 : 1;  // CTAs per tile
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int first_tile = PAIR ? (int)cluster_id_x() : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)cluster_count_x() : (int)gridDim.x;
  const int b_tile_bytes = (p.block_n / NC) * kRowBytes;  // this CTA's part of the B tile
  const int sub_bytes = kATileBytes + b_tile_bytes;  // one 64-wide k-block of both operands
  const int stage_bytes = KPS * sub_bytes;
  // layout: operand ring | epilogue staging (8 warps x 2 blocks of 32 rows x 64 B) | barriers | TMEM slot | column statistics
  uint8_t* staging = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kNumEpiWarps * kStagingBytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_stats = reinterpret_cast<float*>(tmem_slot + 4);  // [2][256] per-CTA column partial sums (ep.stats)
  uint32_t* s_sign = reinterpret_cast<uint32_t*>(s_stats + 512);  // [128] stats mode 3: per column pair, 0xffff where scale >= 0

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();  // the next kernel's CTAs may take this SM as soon as this CTA leaves it (they wait in their own prologue)
  const long long t_entry = p.dbg ? clock64() : 0;
  if (p.dbg && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    p.dbg[blockIdx.x * 16 + 4] = (long long)ns;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (MODE == 1 && p.c2_chunks > 0) tma_prefetch_desc(&tma_a2);
    if (p.out_tma & 1) tma_prefetch_desc(&tma_out);
    if (p.out_tma & 2) tma_prefetch_desc(&tma_out2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], MODE == 2 ? 2 : 1);  // mode 2: two producer warps arrive
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kNumEpiWarps * NC);  // pair: the peer's epilogue warps arrive on the leader's barrier
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols);
    else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  }
  if (STATS && p.ep.stats_mode)
    for (int i = threadIdx.x; i < 512; i += (int)blockDim.x) s_stats[i] = 0.f;

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // (the peer's barriers are signalled remotely: their initialisation must be visible)
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the tail of the previous kernel; from here on its results are read
  if (STATS && p.ep.stats_mode == 3) {
    for (int i = threadIdx.x; i < p.N / 2; i += (int)blockDim.x)
      s_sign[i] = (p.ep.stats_scale[2 * i] >= 0.f ? 0x0000ffffu : 0u) | (p.ep.stats_scale[2 * i + 1] >= 0.f ? 0xffff0000u : 0u);
    __syncthreads();
  }
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 16 + 5] = clock64() - t_entry;

  if (warp == 0 || (MODE == 2 && warp == 10)) {
    // ------------------------------------------------------------------ TMA producer
    // (whole warp in uniform control flow, one elected lane issues: see elect_one()).  The weight-gradient mode needs
    // 4-6 boxes per k-block and was bound by this warp's issue rate: there warp 0 loads operand A, warp 10 operand B.
    const int part = MODE == 2 ? (warp == 0 ? 1 : 2) : 0;  // 0: both operands, 1: A only, 2: B only
    {
      constexpr int elems_per_row = 64;  // bf16 per 128-byte row
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0;
      // pair: both CTAs' loads complete on the leader's full barriers (the leader alone announces the byte count)
      const uint32_t full_leader = PAIR ? mapa_shared(smem_u32(full_bar), 0) : 0u;
      for (int tile = first_tile; tile < p.num_tiles; tile += tile_step) {
        const TileCoord tc = decode_tile(p, tile);
        const int tx = PAIR ? tc.tx * 2 + (int)rank : tc.tx;  // (an odd tile count leaves the last peer an empty tile)
        const int kb_begin = tc.tz * p.kb_per_split;
        const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
        int m0 = 0, n0 = 0, img = 0, h0 = 0, w0 = 0;
        if (MODE == 0) {
          m0 = tx * kBlockM;
          n0 = tc.ty * p.block_n + (int)rank * (p.block_n / NC);
        } else if (MODE == 1) {
          decode_conv_tile(p, tx, img, h0, w0);
          n0 = tc.ty * p.block_n + (int)rank * (p.block_n / NC);
        } else {
          m0 = tc.tx * kBlockM;  // Cout tile; tc.ty = tap
        }
        // Per-tile constants and incremental k-block state: the producer is a single thread, so its address arithmetic
        // (integer divisions) must stay out of the per-k-block path or it starves the tensor pipe.
        int c_tap = 0, c_chunk = 0;                    // conv fwd: current tap and 64-channel chunk
        int pb = 0, ph0 = 0, pw0 = 0;                  // conv wgrad: current 64-pixel patch
        int box_c[4], box_dh[4], box_dw[4];            // conv wgrad: channel / tap shift of each B box of this tile
        if (MODE == 1) {
          c_tap = kb_begin / p.c1_chunks;              // kb_begin is 0 for convolutions (no split-K), kept general
          c_chunk = kb_begin - c_tap * p.c1_chunks;
        } else if (MODE == 2) {
          decode_conv_tile(p, kb_begin, pb, ph0, pw0);
          const int cin = p.c1_chunks * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = tc.ty * p.block_n + j * 64;
            const int tap = col / cin;
            box_c[j] = tap < p.taps ? col - tap * cin : cin;  // past the last tap: out-of-range channel -> zero fill
            box_dh[j] = p.taps == 9 ? (tap / 3) - 1 : 0;
            box_dw[j] = p.taps == 9 ? (tap % 3) - 1 : 0;
          }
        }
        const int main_kb = 9 * p.c1_chunks;
        for (int kb0 = kb_begin; kb0 < kb_end; kb0 += KPS) {
          const int s = stage;
          const int nsub = min(KPS, kb_end - kb0);
          const long long c0 = p.dbg ? clock64() : 0;
          mbar_wait(&empty_bar[s | (p.commit_group - 1)], phase ^ 1u);  // slots are released per commit group
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
          if (p.dbg) w_empty += clock64() - c0;
          const bool leader = elect_one();
          if (leader) {
            const int bytes = part == 0 ? p.tx_bytes
                                        : (part == 1 ? p.a_boxes : p.b_boxes) * 64 * kRowBytes;  // (mode 2: KPS == 1)
            if (!PAIR || rank == 0) mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(nsub * bytes * NC));
          }
          const uint32_t fb = full_leader + 8u * (uint32_t)s;
#define PE_LOAD_2D(map, dst, c0, c1)                                  \
  do {                                                                \
    if (PAIR) tma_load_2d_pair(map, fb, dst, c0, c1);                 \
    else tma_load_2d(map, &full_bar[s], dst, c0, c1);                 \
  } while (0)
#define PE_LOAD_4D(map, dst, c0, c1, c2, c3)                          \
  do {                                                                \
    if (PAIR) tma_load_4d_pair(map, fb, dst, c0, c1, c2, c3);         \
    else tma_load_4d(map, &full_bar[s], dst, c0, c1, c2, c3);         \
  } while (0)
          for (int sub = 0; sub < nsub; ++sub) {
          const int kb = kb0 + sub;
          uint8_t* sa = smem + (size_t)s * stage_bytes + (size_t)sub * sub_bytes;
          uint8_t* sb = sa + kATileBytes;
          if (!leader) {
            // only the elected lane issues; the others just keep the per-k-block state below in step
          } else if (MODE == 0) {
            const int k0 = kb * elems_per_row;
            if (!p.a_mn) {
              PE_LOAD_2D(&tma_a, sa, k0, m0);
            } else {
              for (int j = 0; j < p.a_boxes; ++j)
                PE_LOAD_2D(&tma_a, sa + j * (elems_per_row * kRowBytes), m0 + j * elems_per_row, k0);
            }
            if (!p.b_mn) {
              PE_LOAD_2D(&tma_b, sb, k0, n0);
            } else {
              for (int j = 0; j < p.b_boxes; ++j)
                PE_LOAD_2D(&tma_b, sb + j * (elems_per_row * kRowBytes), n0 + j * elems_per_row, k0);
            }
          } else if (MODE == 1) {
            if (kb < main_kb) {
              const int kh = c_tap >= 6 ? 2 : (c_tap >= 3 ? 1 : 0);
              PE_LOAD_4D(&tma_a, sa, c_chunk * 64, w0 + (c_tap - 3 * kh) - 1, h0 + kh - 1, img);
            } else {
              PE_LOAD_4D(&tma_a2, sa, (kb - main_kb) * 64, w0, h0, img);
            }
            PE_LOAD_2D(&tma_b, sb, kb * 64, n0);
          } else {
            // k-block = one 64-pixel patch; A = dy (Cout-major); B = x shifted by the tap(s) this tile owns: output
            // column n = tap * Cin + ci, 64-wide boxes, several taps per tile when Cin is small (A is loaded once)
            if (part != 2)
              for (int j = 0; j < p.a_boxes; ++j)
                tma_load_4d(&tma_a, &full_bar[s], sa + j * (64 * kRowBytes), m0 + j * 64, pw0, ph0, pb);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (part != 1 && j < p.b_boxes)
                tma_load_4d(&tma_b, &full_bar[s], sb + j * (64 * kRowBytes), box_c[j], pw0 + box_dw[j], ph0 + box_dh[j], pb);
          }
          // per-k-block state, advanced by every lane
          if (MODE == 1) {
            if (kb < main_kb && ++c_chunk == p.c1_chunks) {
              c_chunk = 0;
              ++c_tap;
            }
          } else if (MODE == 2) {
            pw0 += p.tw;  // next patch: row-major over (image, patch row, patch column)
            if (pw0 >= p.tiles_w * p.tw) {
              pw0 = 0;
              ph0 += p.th;
              if (ph0 >= p.tiles_h * p.th) {
                ph0 = 0;
                ++pb;
              }
            }
          }
          }  // sub
#undef PE_LOAD_2D
#undef PE_LOAD_4D
          __syncwarp();
        }
      }
      if (p.dbg && lane == 0 && warp == 0) p.dbg[blockIdx.x * 16 + 3] = w_empty;
    }
  } else if (warp == 1 && (!PAIR || rank == 0)) {
    // ------------------------------------------------------------------ MMA issuer (pair: the leader CTA's)
    // The whole warp walks the loop (uniform control flow), one elected lane issues the tcgen05 instructions.
    {
      const uint32_t idesc = umma_idesc(UMMA_BF16, kBlockM * NC, p.block_n, p.a_mn, p.b_mn);
      constexpr int k_rows = 16;  // UMMA_K (bf16)
      // start-address step (>> 4) between the four MMAs of a 64-deep k-block: 32 B along a K-major row, 16 rows of an
      // MN-major tile; compile-time for the convolution modes, so the issue loop adds immediates
      const uint32_t a_inc = MODE == 1 ? 2u : (MODE == 2 ? (uint32_t)(k_rows * kRowBytes) >> 4 : p.a_inc);
      const uint32_t b_inc = MODE == 1 ? 2u : (MODE == 2 ? (uint32_t)(k_rows * kRowBytes) >> 4 : p.b_inc);
      constexpr int block_k_rows = 64;
      // MN-major: 64-wide (bf16) MN blocks are separate TMA boxes, block_k_rows * 128 B apart
      const uint32_t a_lbo = p.a_mn ? (uint32_t)(block_k_rows * kRowBytes) : 16u;
      const uint32_t b_lbo = p.b_mn ? (uint32_t)(block_k_rows * kRowBytes) : 16u;
      uint32_t local = 0;
      int stage = 0;
      uint32_t phase = 0;
      long long w_full = 0, w_tmem = 0;
      const long long t_begin = p.dbg ? clock64() : 0;
      // descriptor templates: everything but the 14-bit start-address field is loop invariant
      const uint64_t da0 = umma_desc_sw128(0, a_lbo, 1024), db0 = umma_desc_sw128(0, b_lbo, 1024);
      const uint32_t da_lo = (uint32_t)da0, da_hi = (uint32_t)(da0 >> 32), db_lo = (uint32_t)db0, db_hi = (uint32_t)(db0 >> 32);
      const uint32_t smem_base = smem_u32(smem);
      const int gmask = p.commit_group - 1;  // smem slots are handed back in groups of 1, 2 or 4 k-blocks
      for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++local) {
        const TileCoord tc = decode_tile(p, tile);
        const int kb_begin = tc.tz * p.kb_per_split;
        const int num_kb = min(p.kb_total, kb_begin + p.kb_per_split) - kb_begin;
        const uint32_t acc = local & 1u;
        const long long c1 = p.dbg ? clock64() : 0;
        mbar_wait(&tmem_empty_bar[acc], ((local >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator stage
        if (p.dbg) w_tmem += clock64() - c1;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.acc_stride;
        for (int i0 = 0; i0 < num_kb; i0 += KPS) {
          const int nsub = min(KPS, num_kb - i0);
          const uint32_t s0 = (smem_base + (uint32_t)stage * (uint32_t)stage_bytes) >> 4;
          // Only the elected lane waits: the warp-level loop then has no divergent exit, which lets the compiler keep
          // stage / phase / descriptor words in uniform registers instead of moving them there (R2UR) every k-block.
          if (elect_one()) {
            const long long c2 = p.dbg ? clock64() : 0;
            mbar_wait(&full_bar[stage], phase);
            if (p.dbg) w_full += clock64() - c2;
            tc_fence_after();
            // only the low descriptor word (14-bit start address >> 4, no carry: shared memory ends below 256 KB) moves
            // from MMA to MMA: one 32-bit add per operand, which the compiler keeps on the uniform datapath
            uint32_t a_sub = da_lo + s0, b_sub = db_lo + s0 + (kATileBytes >> 4);
            for (int sub = 0; sub < nsub; ++sub) {
              uint32_t a_lo = a_sub, b_lo = b_sub;
              const uint32_t first = (i0 + sub) > 0 ? 1u : 0u;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (PAIR) tc_mma_bf16_pair_lh(d_tmem, a_lo, da_hi, b_lo, db_hi, idesc, k > 0 ? 1u : first);
                else tc_mma_bf16_lh(d_tmem, a_lo, da_hi, b_lo, db_hi, idesc, k > 0 ? 1u : first);
                a_lo += a_inc;
                b_lo += b_inc;
              }
              a_sub += (uint32_t)sub_bytes >> 4;
              b_sub += (uint32_t)sub_bytes >> 4;
            }
            // tcgen05.commit costs the issuing thread ~200 cycles: narrow tiles (short MMAs) release their slots in
            // groups -- the commit on the group's last slot covers every earlier MMA, the producer waits on that slot
            if (PAIR) {  // multicast commits: the slot is free / the accumulator is ready in both CTAs
              if ((stage & gmask) == gmask) tc_commit_pair(&empty_bar[stage], 3);
              if (i0 + nsub >= num_kb) tc_commit_pair(&tmem_full_bar[acc], 3);
            } else {
              if ((stage & gmask) == gmask) tc_commit(&empty_bar[stage]);
              if (i0 + nsub >= num_kb) tc_commit(&tmem_full_bar[acc]);
            }
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (p.dbg && lane == 0) {
        p.dbg[blockIdx.x * 16 + 0] = clock64() - t_begin;
        p.dbg[blockIdx.x * 16 + 1] = w_full;
        p.dbg[blockIdx.x * 16 + 2] = w_tmem;
        p.dbg[blockIdx.x * 16 + 6] = clock64() - t_entry;
      }
    }
  } else if (warp >= 2 && warp < 2 + kNumEpiWarps) {
    // -------------------------------------------------------------------- epilogue (8 warps, 2 per TMEM lane quarter)
    const pe_epilogue& ep = p.ep;
    const int stats_mode = STATS ? ep.stats_mode : 0;  // (compiled out of the instantiations without statistics)
    const int act = HEAVY ? ep.act : PE_ACT_NONE;
    const unsigned drop_thresh = HEAVY ? ep.drop_thresh : 0u;
    const int aux_mode = (!HEAVY && ep.aux_mode == PE_AUX_GELU_GRAD) ? PE_AUX_NONE : ep.aux_mode;
    const unsigned long long drop_seed = pe_salted(ep.drop_seed);
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int pair = (warp - 2) >> 2;    // which of the two warps of the quarter: takes chunks c with (c & 1) == pair
    const int r = q * 32 + lane;
    const int nchunks = (p.block_n + 31) / 32;
    uint8_t* stg = staging + (warp - 2) * kStagingBytes;
    uint32_t stg_flip = 0;
    uint32_t local = 0;
    long long e_wait = 0, e_ld = 0, e_work = 0, e_last = 0;
    const uint32_t tmem_empty_leader = PAIR ? mapa_shared(smem_u32(tmem_empty_bar), 0) : 0u;
    // BatchNorm-backward sums read column-wise out of the staging blocks (every 32-column chunk of the launch is whole)
    const bool stats_x_block = stats_mode >= 2 && (p.out_tma & 1) && ep.out_mode == PE_OUT_BF16 && (p.N % 32) == 0 &&
                               (p.block_n % 32) == 0;
    for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++local) {
      const TileCoord tc = decode_tile(p, tile);
      const int tx = PAIR ? tc.tx * 2 + (int)rank : tc.tx;
      int m0 = 0, n0 = 0, img = 0, h0 = 0, w0 = 0;
      long long grow;
      bool row_ok;
      long long out_col_off = 0;
      if (MODE == 1) {
        decode_conv_tile(p, tx, img, h0, w0);
        n0 = tc.ty * p.block_n;
        const int h = h0 + r / p.tw, w = w0 + r % p.tw;
        grow = ((long long)img * p.H + h) * p.W + w;
        row_ok = (h < p.H) && (w < p.W) && (grow < (long long)p.M);  // (pair: the last peer tile may lie past the batch)
      } else {
        m0 = tx * kBlockM;
        n0 = tc.ty * p.block_n;  // wgrad: columns are tap-major (n = tap * Cin + ci), several taps per tile
        grow = m0 + r;
        row_ok = grow < p.M;
      }
      // 64-byte row blocks -> one of this warp's two staging blocks (each lane its own row; 16-byte units XOR-swizzled
      // as SWIZZLE_64B expects, which also makes the writes bank-conflict free) -> one TMA tensor store / reduction.
      // Rows and columns outside the tensor are clipped by the TMA unit, so every lane takes part regardless of row_ok.
      // (stats_x_block: the first staging block holds the BatchNorm input of this chunk, see below; outputs then leave
      // through the second block alone, which is free again once the previous chunk's store has been read)
      auto stage_block = [&](const CUtensorMap* map, const uint4 (&w)[4], int col, bool reduce) {
        uint8_t* buf = stg + (stats_x_block ? 1u : stg_flip) * (kStagingBytes / 2);
        stg_flip ^= 1u;
        if (lane == 0) {
          if (stats_x_block) bulk_wait_read_all();
          else bulk_wait_read_1();  // the store issued two blocks ago (same buffer) has been read
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(buf + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) = w[u];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (MODE == 1) {
            if (reduce) tma_reduce_add_4d(map, buf, col, w0, h0 + (q * 32) / p.tw, img);
            else tma_store_4d(map, buf, col, w0, h0 + (q * 32) / p.tw, img);
          } else {
            if (reduce) tma_reduce_add_2d(map, buf, col, m0 + q * 32);
            else tma_store_2d(map, buf, col, m0 + q * 32);
          }
          bulk_commit_group();
        }
      };
      auto stage_and_store = [&](const CUtensorMap* map, const float (&vals)[32], int col) {  // bf16 chunk
        uint4 w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          w[u] = make_uint4(pack_bf16(vals[u * 8], vals[u * 8 + 1]), pack_bf16(vals[u * 8 + 2], vals[u * 8 + 3]),
                            pack_bf16(vals[u * 8 + 4], vals[u * 8 + 5]), pack_bf16(vals[u * 8 + 6], vals[u * 8 + 7]));
        stage_block(map, w, col, false);
      };
      auto stage_and_store_f32 = [&](const CUtensorMap* map, const float (&vals)[32], int col, bool reduce) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // two 16-column halves of 64 bytes per row
          uint4 w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            w[u] = make_uint4(__float_as_uint(vals[h * 16 + u * 4]), __float_as_uint(vals[h * 16 + u * 4 + 1]),
                              __float_as_uint(vals[h * 16 + u * 4 + 2]), __float_as_uint(vals[h * 16 + u * 4 + 3]));
          stage_block(map, w, col + h * 16, reduce);
        }
      };
      const uint32_t acc = local & 1u;
      const long long e_t0 = p.dbg ? clock64() : 0;
      mbar_wait_relaxed(&tmem_full_bar[acc], (local >> 1) & 1u);
      tc_fence_after();
      const long long e_t1 = p.dbg ? clock64() : 0;
      e_wait += e_t1 - e_t0;
      const uint32_t t_row = tmem_base + acc * (uint32_t)p.acc_stride + ((uint32_t)(q * 32) << 16);
      for (int c = pair; c < nchunks; c += 2) {
        uint32_t v[32];
        const long long e_t2 = p.dbg ? clock64() : 0;
        tmem_ld32(t_row + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        e_ld += p.dbg ? clock64() - e_t2 : 0;
        if (!row_ok && !stats_mode && !p.out_tma) continue;  // (warp-uniform paths below need every lane)
        const int col0 = n0 + c * 32;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * ep.alpha;
        const bool full = (col0 + 32 <= p.N) && (c * 32 + 32 <= p.block_n);
        // rows outside the tensor only take part in the column-statistics shuffles and (with junk values that the TMA
        // unit clips) in the staged stores; their global loads / direct stores stay predicated on row_ok
        if (row_ok || p.out_tma) {
        if (ep.bias) {
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
              f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) f[j] += __ldg(ep.bias + col0 + j);
          }
        }
        if (act == PE_ACT_GELU_SAVE_GRAD) {
          // out = dropout(gelu(v)); out2 = d out / d v = gelu'(v) * dropout factor (the backward GEMM just multiplies)
          float gp[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) gelu_erf_both(f[j], f[j], gp[j]);
          if (drop_thresh) {
            const unsigned long long e0 = (unsigned long long)grow * (unsigned long long)p.N + (unsigned long long)col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint32_t km = dropout_keep8(drop_seed, (e0 + j) >> 3, drop_thresh);
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const float sc = ((km >> t) & 1u) ? ep.drop_scale : 0.f;
                f[j + t] *= sc;
                gp[j + t] *= sc;
              }
            }
          }
          __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(ep.out2) + grow * ep.ld2 + col0;
          if (p.out_tma & 2) {
            stage_and_store(&tma_out2, gp, col0);
          } else if (!row_ok) {
          } else if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              *reinterpret_cast<uint4*>(o2 + j) = make_uint4(pack_bf16(gp[j], gp[j + 1]), pack_bf16(gp[j + 2], gp[j + 3]),
                                                             pack_bf16(gp[j + 4], gp[j + 5]), pack_bf16(gp[j + 6], gp[j + 7]));
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N && c * 32 + j < p.block_n) o2[j] = __float2bfloat16(gp[j]);
          }
        } else {
        if (act == PE_ACT_GELU) {
          if (ep.out2) {
            __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(ep.out2) + grow * ep.ld2 + col0;
            if (p.out_tma & 2) {
              stage_and_store(&tma_out2, f, col0);
            } else if (!row_ok) {
            } else if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                *reinterpret_cast<uint4*>(o2 + j) = make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]),
                                                               pack_bf16(f[j + 4], f[j + 5]), pack_bf16(f[j + 6], f[j + 7]));
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N && c * 32 + j < p.block_n) o2[j] = __float2bfloat16(f[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
        }
        if (drop_thresh) {  // element index row * N + col; N % 8 == 0 is enforced by the launcher
          const unsigned long long e0 = (unsigned long long)grow * (unsigned long long)p.N + (unsigned long long)col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint32_t km = dropout_keep8(drop_seed, (e0 + j) >> 3, drop_thresh);
#pragma unroll
            for (int t = 0; t < 8; ++t) f[j + t] = ((km >> t) & 1u) ? f[j + t] * ep.drop_scale : 0.f;
          }
        }
        }
        if (aux_mode != PE_AUX_NONE) {
          const __nv_bfloat16* ax = reinterpret_cast<const __nv_bfloat16*>(ep.aux) + grow * ep.ld_aux + col0;
          float a[32];
          if (!row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] = 0.f;
          } else if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 u = *reinterpret_cast<const uint4*>(ax + j);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 x2 = __bfloat1622float2(h2[t]);
                a[j + 2 * t] = x2.x;
                a[j + 2 * t + 1] = x2.y;
              }
            }
          } else {
            for (int j = 0; j < 32; ++j)
              a[j] = (col0 + j < p.N && c * 32 + j < p.block_n) ? __bfloat162float(ax[j]) : 0.f;
          }
          if (aux_mode == PE_AUX_ADD) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += a[j];
          } else if (aux_mode == PE_AUX_MUL) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= a[j];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= gelu_erf_grad(a[j]);
          }
        }
        }
        // BatchNorm batch statistics of a bf16 output that leaves through the staging block: summed column-wise out of
        // the staged block after the store below (32 two-byte loads per lane instead of the 62-shuffle butterfly)
        // The first BatchNorm-backward pass (modes 2 / 3) goes the same way: the lanes first put the BatchNorm input of
        // their rows (mode 3: of the pooled pair, the element MaxPool selected) into the other staging block.
        const bool stats_from_block = stats_x_block || (stats_mode == 1 && (p.out_tma & 1) && ep.out_mode == PE_OUT_BF16);
        if (stats_x_block) {
          __syncwarp();  // (the previous chunk's column pass has finished reading the block)
          uint4 xs[4];
          if (stats_mode == 2) {
            const uint4* xp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(ep.stats_x) +
                                                             grow * (long long)p.N + col0);
#pragma unroll
            for (int u = 0; u < 4; ++u) xs[u] = row_ok ? xp[u] : make_uint4(0u, 0u, 0u, 0u);
          } else {
            // the BatchNorm input is twice as wide as this gradient (MaxPool (1,2) in between): output pixel `grow` owns
            // input pixels 2*grow and 2*grow+1 and the gradient goes to the maximum of the activated pair.  LeakyReLU of
            // an affine map is monotonic, so that is the larger input where the BatchNorm scale is positive and the
            // smaller one where it is negative: packed bf16 max / min and a per-column sign mask, no arithmetic.
            const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(ep.stats_x) + grow * 2LL * p.N + col0;
            const uint4* sg = reinterpret_cast<const uint4*>(s_sign + (col0 >> 1));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint4 u0 = make_uint4(0u, 0u, 0u, 0u), u1 = u0;
              if (row_ok) {
                u0 = *reinterpret_cast<const uint4*>(xp + 8 * u);
                u1 = *reinterpret_cast<const uint4*>(xp + p.N + 8 * u);
              }
              const uint4 m = sg[u];
              const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&u0);
              const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u1);
              const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
              uint32_t sel[4];
#pragma unroll
              for (int t2 = 0; t2 < 4; ++t2) {
                const __nv_bfloat162 mx = __hmax2(a2[t2], b2[t2]), mn = __hmin2(a2[t2], b2[t2]);
                sel[t2] = (*reinterpret_cast<const uint32_t*>(&mx) & mw[t2]) | (*reinterpret_cast<const uint32_t*>(&mn) & ~mw[t2]);
              }
              xs[u] = make_uint4(sel[0], sel[1], sel[2], sel[3]);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(stg + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) = xs[u];
        }
        if (p.out_tma & 1) {
          if (ep.out_mode == PE_OUT_BF16) stage_and_store(&tma_out, f, col0);
          else stage_and_store_f32(&tma_out, f, col0, ep.out_mode == PE_OUT_F32_ATOMIC);
          if (stats_x_block) {
            // lane = column: g = v * lrelu'(x * scale + shift) over the 32 rows; sums of g and g * x
            const uint8_t* vb = stg + kStagingBytes / 2 + (lane & 7) * 2;
            const uint8_t* xb = stg + (lane & 7) * 2;
            const uint32_t rowmask = __ballot_sync(0xffffffffu, row_ok);
            const int u0 = lane >> 3;
            const int uo[4] = {(u0 ^ 0) << 4, (u0 ^ 1) << 4, (u0 ^ 2) << 4, (u0 ^ 3) << 4};
            const float sc = __ldg(ep.stats_scale + col0 + lane), sh = __ldg(ep.stats_shift + col0 + lane);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int off = uo[(j >> 1) & 3] + j * 64;
              const float v = __uint_as_float((uint32_t)(*reinterpret_cast<const unsigned short*>(vb + off)) << 16);
              const float x = __uint_as_float((uint32_t)(*reinterpret_cast<const unsigned short*>(xb + off)) << 16);
              float g = v * (fmaf(x, sc, sh) > 0.f ? 1.f : ep.stats_slope);
              if (rowmask != 0xffffffffu && !((rowmask >> j) & 1u)) g = 0.f;
              s1 += g;
              s2 = fmaf(g, x, s2);
            }
            atomicAdd(&s_stats[col0 + lane], s1);
            atomicAdd(&s_stats[256 + col0 + lane], s2);
          } else if (stats_from_block) {
            // lane = column: element (row j, column lane) of the block just staged (SWIZZLE_64B: 16-byte unit index
            // XOR bits 1-2 of the row); rows outside the tensor hold junk and are skipped
            const uint8_t* blk = stg + (stg_flip ^ 1u) * (kStagingBytes / 2) + (lane & 7) * 2;
            const uint32_t rowmask = __ballot_sync(0xffffffffu, row_ok);
            const int u0 = lane >> 3;
            const uint8_t* bk[4] = {blk + ((u0 ^ 0) << 4), blk + ((u0 ^ 1) << 4), blk + ((u0 ^ 2) << 4), blk + ((u0 ^ 3) << 4)};
            float s1 = 0.f, s2 = 0.f;
            if (rowmask == 0xffffffffu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float((uint32_t)(*reinterpret_cast<const unsigned short*>(bk[(j >> 1) & 3] + j * 64)) << 16);
                s1 += v;
                s2 = fmaf(v, v, s2);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float((uint32_t)(*reinterpret_cast<const unsigned short*>(bk[(j >> 1) & 3] + j * 64)) << 16);
                if ((rowmask >> j) & 1u) {
                  s1 += v;
                  s2 = fmaf(v, v, s2);
                }
              }
            }
            if (full) {
              atomicAdd(&s_stats[col0 + lane], s1);
              atomicAdd(&s_stats[256 + col0 + lane], s2);
            }
          }
        } else if (!row_ok) {
          // nothing to write for rows outside the tensor
        } else if (ep.out_mode == PE_OUT_BF16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + grow * ep.ldc + out_col_off + col0;
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]),
                                                            pack_bf16(f[j + 4], f[j + 5]), pack_bf16(f[j + 6], f[j + 7]));
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N && c * 32 + j < p.block_n) o[j] = __float2bfloat16(f[j]);
          }
        } else if (ep.out_mode == PE_OUT_F32) {
          float* o = reinterpret_cast<float*>(ep.out) + grow * ep.ldc + out_col_off + col0;
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N && c * 32 + j < p.block_n) o[j] = f[j];
          }
        } else {
          float* o = reinterpret_cast<float*>(ep.out) + grow * ep.ldc + out_col_off + col0;
          if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)  // 16-byte vector reductions: 4x fewer L2 atomic operations
              atomicAdd(reinterpret_cast<float4*>(o + j), make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]));
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N && c * 32 + j < p.block_n) atomicAdd(o + j, f[j]);
          }
        }
      }
      // hand the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(tmem_empty_leader + 8u * acc);
        else mbar_arrive(&tmem_empty_bar[acc]);
      }
      if (p.dbg) { e_last = clock64() - e_t1; e_work += e_last; }
    }
    if (p.out_tma && lane == 0) bulk_wait_all();  // staged stores have landed before the CTA may exit
    if (p.dbg && warp == 2 && lane == 0) {
      p.dbg[blockIdx.x * 16 + 8] = e_wait;
      p.dbg[blockIdx.x * 16 + 9] = e_ld;
      p.dbg[blockIdx.x * 16 + 10] = e_work;
      p.dbg[blockIdx.x * 16 + 11] = e_last;
    }
    if (stats_mode) {
      asm volatile("bar.sync 2, 256;" ::: "memory");  // the 8 epilogue warps
      for (int i = threadIdx.x - 64; i < 2 * p.N; i += 32 * kNumEpiWarps) {
        const int which = i / p.N, col = i - which * p.N;
        atomicAdd(ep.stats + (long long)which * p.N + col, (double)s_stats[which * 256 + col]);
      }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the leader's MMAs read the peer's shared memory and write its TMEM until the very end
  else __syncthreads();
  if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 16 + 7] = clock64() - t_entry;
  if (warp == 2) {
    if (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace pe

// =================================================================================================
// host launchers
// =================================================================================================
using pe::TcParams;

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

// Tensor maps for the epilogue's staged bf16 stores (32-column x 32-row boxes, SWIZZLE_64B).  Returns false when the
// output cannot be addressed by TMA (pitch / base not 16-byte aligned): the kernel then stores from the lanes directly.
static bool out_tmap(CUtensorMap* m, const TcParams& p, void* base, long long ld, int conv_B, int elem_bytes) {
  if (!base || (ld * elem_bytes) % 16 || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const uint32_t bc = 64 / elem_bytes;  // box columns: 64 bytes per row
  const uint64_t eb = (uint64_t)elem_bytes;
  if (p.mode == 1) {
    if (32 % p.tw) return false;
    uint64_t dims[4] = {(uint64_t)p.N, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)conv_B};
    uint64_t str[3] = {(uint64_t)ld * eb, (uint64_t)p.W * ld * eb, (uint64_t)p.H * p.W * ld * eb};
    uint32_t box[4] = {bc, (uint32_t)p.tw, (uint32_t)(32 / p.tw), 1};
    return pe_host::encode_tmap(m, dt, 4, base, dims, str, box, 64) == PE_OK;
  }
  uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.M};
  uint64_t str[1] = {(uint64_t)ld * eb};
  uint32_t box[2] = {bc, 32};
  return pe_host::encode_tmap(m, dt, 2, base, dims, str, box, 64) == PE_OK;
}

// CTA-pair (cta_group::2) execution of a launch: 256-row tiles, each CTA of a cluster of two holds half of the B tile.
// K-major B is split by rows of the N x 64 box, MN-major B by its 64-wide column boxes.
static bool want_pair(int mode, int block_n, int b_mn, int m_tiles) {
  // Measured (profiles/r02_tile_engine_notes.md): the main loop of 256-wide tiles gets 9 % shorter, narrower tiles do not
  // change (they are bound by the ~85-cycle issue interval of SS-operand MMAs, not by operand traffic), and the training
  // step is unchanged within noise -- so pairs are opt-in: PE_TC_PAIR=1 every eligible launch, 2 only N = 256.
  static const int knob = getenv("PE_TC_PAIR") ? atoi(getenv("PE_TC_PAIR")) : 0;
  if (!knob || mode == 2 || m_tiles < 2) return false;
  if (knob == 2 && block_n != 256) return false;
  if (block_n % 32 || block_n < 32) return false;   // N / 2 per CTA, multiples of 16 (the MMA's N is block_n)
  if (b_mn && (block_n % 128)) return false;        // whole 64-wide boxes per CTA
  return pe_host::num_sms() % 2 == 0;
}

static int launch_tc(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, TcParams& p, dim3 tiles,
                     cudaStream_t stream, int conv_B = 0, bool pair = false) {
  const int nc = pair ? 2 : 1;
  const int sub_bytes = pe::kATileBytes + (p.block_n / nc) * pe::kRowBytes;  // per CTA
  // narrow tiles (short MMAs) are bound by the per-stage barrier round trip of the producer / MMA threads: two k-blocks
  // per stage halve it (wide tiles keep one: two 96 KB stages would not cover the TMA latency)
  p.kps = (p.block_n <= 128 && p.mode != 2) ? 2 : 1;
  if (const char* env = getenv("PE_TC_KPS")) {  // tuning knob
    const int v = atoi(env);
    if (v == 1 || (v == 2 && 2 * 2 * sub_bytes <= 192 * 1024)) p.kps = v;
  }
  const int stage_bytes = p.kps * sub_bytes;
  int stages = (192 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (const char* env = getenv("PE_TC_STAGES")) {  // tuning knob
    const int v = atoi(env);
    if (v >= 2 && v < stages) stages = v;
  }
  if (stages < 2) return PE_ERR_BAD_SHAPE;
  // narrow tiles: MMAs are short, so the per-k-block commit would dominate the issue thread
  p.commit_group = 1;
  if (p.kps == 1 && p.block_n <= 128 && stages >= 4) p.commit_group = 2;
  if (p.kps == 1 && p.block_n <= 64 && stages >= 8) p.commit_group = 4;
  if (p.kps == 2 && p.block_n <= 64 && stages >= 4) p.commit_group = 2;
  if (const char* env = getenv("PE_TC_COMMIT_GROUP")) {  // tuning knob
    const int v = atoi(env);
    if ((v == 1 || v == 2 || v == 4) && stages >= 2 * v) p.commit_group = v;
  }
  stages -= stages % p.commit_group;
  p.stages = stages;
  p.a_inc = p.a_mn ? (16 * pe::kRowBytes) >> 4 : 2;  // 16 rows of an MN-major tile / 32 bytes along a K-major row
  p.b_inc = p.b_mn ? (16 * pe::kRowBytes) >> 4 : 2;
  p.tmem_cols = pow2_cols(2 * p.block_n);
  p.tx_bytes = p.mode == 2 ? (p.a_boxes + p.b_boxes) * 64 * pe::kRowBytes : sub_bytes;  // per k-block
  p.acc_stride = p.tmem_cols / 2;
  p.tiles_x = ((int)tiles.x + nc - 1) / nc;  // pair: columns of 256-row tiles
  p.tiles_y = (int)tiles.y;
  p.tiles_z = (int)tiles.z;
  p.num_tiles = p.tiles_x * p.tiles_y * p.tiles_z;
  p.dbg = p.ep.debug;
  if (p.ep.drop_thresh && (p.N % 8)) return PE_ERR_BAD_SHAPE;
  if (p.ep.act == PE_ACT_GELU_SAVE_GRAD && !p.ep.out2) return PE_ERR_BAD_SHAPE;
  if (p.ep.stats_mode) {
    if (!p.ep.stats || p.N > 256 || (p.N % 32) || p.tiles_y != 1 || p.mode == 2) return PE_ERR_BAD_SHAPE;
    if (p.ep.stats_mode < 1 || p.ep.stats_mode > 3) return PE_ERR_BAD_SHAPE;
    if (p.ep.stats_mode >= 2 && (!p.ep.stats_x || !p.ep.stats_scale || !p.ep.stats_shift)) return PE_ERR_BAD_SHAPE;
    // the sums are read out of the staged bf16 output blocks (whole 32-column chunks, TMA-addressable output)
    if (p.ep.out_mode != PE_OUT_BF16 || (p.block_n % 32)) return PE_ERR_BAD_SHAPE;
  }
  // staged TMA stores for bf16 outputs of the GEMM / conv modes (whole 32-column chunks only)
  CUtensorMap tout = ta, tout2 = ta;
  p.out_tma = 0;
  static const bool tma_store_off = getenv("PE_TC_DIRECT_STORE") != nullptr;  // tuning knob
  if (!tma_store_off && (p.block_n % 32 == 0 || p.tiles_y == 1)) {
    if (out_tmap(&tout, p, p.ep.out, p.ep.ldc, conv_B, p.ep.out_mode == PE_OUT_BF16 ? 2 : 4)) p.out_tma |= 1;
    if (p.ep.out2 && p.ep.act != PE_ACT_NONE && out_tmap(&tout2, p, p.ep.out2, p.ep.ld2, conv_B, 2)) p.out_tma |= 2;
  }
  if (p.ep.stats_mode && !(p.out_tma & 1)) return PE_ERR_BAD_SHAPE;
  const size_t smem = (size_t)stages * stage_bytes + pe::kNumEpiWarps * pe::kStagingBytes +
                      (2 * stages + 4) * sizeof(uint64_t) + 16 + 2048 + 512;
  const bool heavy = p.ep.act != PE_ACT_NONE || p.ep.drop_thresh != 0 || p.ep.aux_mode == PE_AUX_GELU_GRAD;
  int feat = (p.ep.stats_mode ? 1 : 0) | (heavy ? 2 : 0);
  if (p.mode == 1 && (feat & 2)) feat = 3;  // (convolutions: the element-wise paths only exist in the full instantiation)
  if (p.mode == 2 && feat) return PE_ERR_BAD_SHAPE;
  if (pair && ((feat & 1) || (p.mode == 1 && feat))) return PE_ERR_BAD_SHAPE;  // (callers switch pairs off for these)
  static bool attr_set = false;
  if (!attr_set) {
    const int sz = 227 * 1024;
#define PE_TC_ATTR(M, K, P, F) \
  (cudaFuncSetAttribute(pe::tc_tile_kernel<M, K, P, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, sz) == cudaSuccess)
#define PE_TC_ATTR_K(M, P, F) (PE_TC_ATTR(M, 1, P, F) && PE_TC_ATTR(M, 2, P, F))
    if (!(PE_TC_ATTR_K(0, 0, 0) && PE_TC_ATTR_K(0, 0, 1) && PE_TC_ATTR_K(0, 0, 2) && PE_TC_ATTR_K(0, 0, 3) &&
          PE_TC_ATTR_K(0, 1, 0) && PE_TC_ATTR_K(0, 1, 2) && PE_TC_ATTR_K(1, 0, 0) && PE_TC_ATTR_K(1, 0, 1) &&
          PE_TC_ATTR_K(1, 0, 3) && PE_TC_ATTR_K(1, 1, 0) && PE_TC_ATTR(2, 1, 0, 0)))
      return PE_ERR_LAUNCH;
#undef PE_TC_ATTR_K
#undef PE_TC_ATTR
    attr_set = true;
  }
  const int units = pe_host::num_sms() / nc;  // CTAs or CTA pairs
  const int grid = (p.num_tiles < units ? p.num_tiles : units) * nc;
  cudaError_t lerr = cudaSuccess;
#define PE_TC_LAUNCH(M, K, P, F)                                                                                     \
  lerr = pe_host::launch_cluster(pe::tc_tile_kernel<M, K, P, F>, dim3(grid),                                          \
                                 dim3((M) == 2 ? pe::kNumThreadsWgrad : pe::kNumThreads), smem, stream, (P) ? 2 : 1, ta, \
                                 ta2, tb, tout, tout2, p)
#define PE_TC_LAUNCH_K(M, P, F)          \
  do {                                   \
    if (p.kps == 1) PE_TC_LAUNCH(M, 1, P, F); \
    else PE_TC_LAUNCH(M, 2, P, F);       \
  } while (0)
  if (p.mode == 2) PE_TC_LAUNCH(2, 1, 0, 0);
  else if (p.mode == 0 && pair) {
    if (feat == 0) PE_TC_LAUNCH_K(0, 1, 0);
    else PE_TC_LAUNCH_K(0, 1, 2);
  } else if (p.mode == 0) {
    if (feat == 0) PE_TC_LAUNCH_K(0, 0, 0);
    else if (feat == 1) PE_TC_LAUNCH_K(0, 0, 1);
    else if (feat == 2) PE_TC_LAUNCH_K(0, 0, 2);
    else PE_TC_LAUNCH_K(0, 0, 3);
  } else if (pair) PE_TC_LAUNCH_K(1, 1, 0);
  else if (feat == 0) PE_TC_LAUNCH_K(1, 0, 0);
  else if (feat == 1) PE_TC_LAUNCH_K(1, 0, 1);
  else PE_TC_LAUNCH_K(1, 0, 3);
#undef PE_TC_LAUNCH_K
#undef PE_TC_LAUNCH
  return (lerr == cudaSuccess && cudaGetLastError() == cudaSuccess) ? PE_OK : PE_ERR_LAUNCH;
}

static void default_ep(pe_epilogue& e) {
  if (e.alpha == 0.f) e.alpha = 1.f;
}

extern "C" int pe_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, int M,
                            int N, int K, const pe_epilogue* ep, int splits, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!A || !B || !ep || !ep->out || M <= 0 || N <= 0 || K <= 0) return PE_ERR_BAD_SHAPE;
  if ((lda % 8) || (ldb % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
    return PE_ERR_BAD_SHAPE;
  if (splits < 1) splits = 1;
  if (splits > 1 && ep->out_mode != PE_OUT_F32_ATOMIC) return PE_ERR_BAD_SHAPE;
  TcParams p{};
  p.mode = 0;
  p.kind = 0;
  p.a_mn = a_mn ? 1 : 0;
  p.b_mn = b_mn ? 1 : 0;
  p.M = M;
  p.N = N;
  // tile width: as wide as N allows (<= 256): one 128x256 tile reads 87 flops per operand byte;
  // MN-major B needs whole 64-wide boxes
  int bn = N >= 256 ? 256 : ((N + 15) / 16) * 16;
  if (p.b_mn) bn = ((bn + 63) / 64) * 64;
  if (bn > 256) bn = 256;
  // Wave quantisation: with 256-wide tiles a [12288 x 512] output is 192 tiles = 1.3 waves of 148 CTAs (the second wave
  // 30 % full); 128-wide tiles make it 384 half-size tiles = 2.6 waves.  Narrow tiles when the last wave would otherwise
  // be less than two-thirds full and the narrower tiling fills it better.
  if (bn == 256 && splits == 1 && (N % 128) == 0 && !ep->stats_mode) {  // (column statistics need one tile per row block)
    const long long sms = pe_host::num_sms();
    const long long t256 = (long long)((M + 127) / 128) * ((N + 255) / 256), t128 = (long long)((M + 127) / 128) * (N / 128);
    auto waste = [&](long long tiles) { return (double)(((tiles + sms - 1) / sms) * sms - tiles) / (double)(((tiles + sms - 1) / sms) * sms); };
    static const int force = getenv("PE_TC_BN") ? atoi(getenv("PE_TC_BN")) : 0;  // tuning knob: 128 / 256
    if (force == 128 || (force == 0 && waste(t256) > 0.33 && waste(t128) < waste(t256) - 0.1)) bn = 128;
  }
  p.block_n = bn;
  const bool pair = !ep->stats_mode && want_pair(0, bn, p.b_mn, (M + 127) / 128);
  const int nc = pair ? 2 : 1;
  p.a_boxes = 2;
  p.b_boxes = bn / 64 / nc;  // (MN-major B: 64-wide boxes per CTA)
  p.kb_total = (K + 63) / 64;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.ep = *ep;
  default_ep(p.ep);

  CUtensorMap ta, tb;
  {
    uint64_t dims[2], str[1];
    uint32_t box[2];
    if (!p.a_mn) { dims[0] = (uint64_t)K; dims[1] = (uint64_t)M; box[0] = 64; box[1] = 128; }
    else         { dims[0] = (uint64_t)M; dims[1] = (uint64_t)K; box[0] = 64; box[1] = 64; }
    str[0] = (uint64_t)lda * 2;
    if (int rc = pe_host::encode_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, str, box)) return rc;
    if (!p.b_mn) { dims[0] = (uint64_t)K; dims[1] = (uint64_t)N; box[0] = 64; box[1] = (uint32_t)(bn / nc); }
    else         { dims[0] = (uint64_t)N; dims[1] = (uint64_t)K; box[0] = 64; box[1] = 64; }
    str[0] = (uint64_t)ldb * 2;
    if (int rc = pe_host::encode_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, dims, str, box)) return rc;
  }
  dim3 grid((M + 127) / 128, (N + bn - 1) / bn, splits);
  return launch_tc(ta, ta, tb, p, grid, reinterpret_cast<cudaStream_t>(stream), 0, pair);
}

// patch shapes: 128-pixel output tiles (fwd) and 64-pixel contraction blocks (wgrad) that tile W in 5 columns
static bool conv_patch(int W, int pixels, int& tw, int& th) {
  // choose the widest tw <= 16 dividing `pixels` such that tiles cover W with little waste
  int best = 0;
  for (int cand = 16; cand >= 1; cand >>= 1) {
    if (pixels % cand) continue;
    if (W % cand == 0) { best = cand; break; }
  }
  if (!best) best = (W >= 16) ? 16 : (W >= 8 ? 8 : (W >= 4 ? 4 : 2));
  tw = best;
  th = pixels / tw;
  return th <= 256;
}

static int nhwc_tmap(CUtensorMap* m, const void* base, int B, int H, int W, int C, int tw, int th) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, 1};
  return pe_host::encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, str, box);
}

extern "C" int pe_conv3x3_nhwc(const void* x, const void* x2, const void* w, int B, int H, int W, int C1, int C2,
                               int Cout, const pe_epilogue* ep, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!x || !w || !ep || !ep->out || B <= 0 || H <= 0 || W <= 0) return PE_ERR_BAD_SHAPE;
  if (C1 % 64 || C2 % 64 || C1 <= 0 || Cout % 16 || Cout <= 0 || (C2 > 0 && !x2)) return PE_ERR_BAD_SHAPE;
  TcParams p{};
  p.mode = 1;
  p.kind = 0;
  p.H = H;
  p.W = W;
  if (!conv_patch(W, 128, p.tw, p.th)) return PE_ERR_BAD_SHAPE;
  p.tiles_w = (W + p.tw - 1) / p.tw;
  p.tiles_h = (H + p.th - 1) / p.th;
  p.c1_chunks = C1 / 64;
  p.c2_chunks = C2 / 64;
  p.M = B * H * W;
  p.N = Cout;
  p.block_n = Cout > 256 ? 256 : Cout;
  const bool conv_plain = !ep->stats_mode && ep->act == PE_ACT_NONE && !ep->drop_thresh && ep->aux_mode != PE_AUX_GELU_GRAD;
  const bool pair = conv_plain && want_pair(1, p.block_n, 0, B * p.tiles_h * p.tiles_w);
  p.kb_total = 9 * p.c1_chunks + p.c2_chunks;
  p.kb_per_split = p.kb_total;
  p.ep = *ep;
  default_ep(p.ep);
  CUtensorMap ta, ta2, tb;
  if (int rc = nhwc_tmap(&ta, x, B, H, W, C1, p.tw, p.th)) return rc;
  ta2 = ta;
  if (C2 > 0)
    if (int rc = nhwc_tmap(&ta2, x2, B, H, W, C2, p.tw, p.th)) return rc;
  {
    const int Ktot = 9 * C1 + C2;
    uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)Cout};
    uint64_t str[1] = {(uint64_t)Ktot * 2};
    uint32_t box[2] = {64, (uint32_t)(p.block_n / (pair ? 2 : 1))};
    if (int rc = pe_host::encode_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, str, box)) return rc;
  }
  dim3 grid(B * p.tiles_h * p.tiles_w, (Cout + p.block_n - 1) / p.block_n, 1);
  return launch_tc(ta, ta2, tb, p, grid, reinterpret_cast<cudaStream_t>(stream), B, pair);
}

extern "C" int pe_conv_wgrad_nhwc(const void* dy, const void* x, float* dw, long long ldw, int B, int H, int W, int C,
                                  int Cout, int taps, int splits, pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!dy || !x || !dw || B <= 0 || H <= 0 || W <= 0) return PE_ERR_BAD_SHAPE;
  if (C % 64 || C <= 0 || C > 256 || Cout % 8 || Cout <= 0 || (taps != 9 && taps != 1)) return PE_ERR_BAD_SHAPE;
  TcParams p{};
  p.mode = 2;
  p.kind = 0;
  p.a_mn = 1;
  p.b_mn = 1;
  p.H = H;
  p.W = W;
  if (!conv_patch(W, 64, p.tw, p.th)) return PE_ERR_BAD_SHAPE;
  p.tiles_w = (W + p.tw - 1) / p.tw;
  p.tiles_h = (H + p.th - 1) / p.th;
  p.taps = taps;
  p.c1_chunks = C / 64;
  p.M = Cout;
  p.N = taps * C;
  // several taps per tile when that keeps the tile <= 256 columns: dy is then loaded once for all of them
  int group = 1;
  if (taps == 9) group = C == 64 ? 3 : (C == 128 ? 2 : 1);
  p.block_n = group * C;
  p.a_boxes = Cout > 64 ? 2 : 1;  // rows >= Cout of the accumulator are never stored: skip their operand box
  p.b_boxes = p.block_n / 64;
  p.kb_total = B * p.tiles_h * p.tiles_w;
  const int n_tiles = (p.N + p.block_n - 1) / p.block_n;
  const int m_tiles = (Cout + 127) / 128;
  if (splits < 1) {
    splits = pe_host::num_sms() / (m_tiles * n_tiles);
    if (splits < 1) splits = 1;
  }
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  pe_epilogue e{};
  e.out = dw;
  e.ldc = ldw;
  e.out_mode = PE_OUT_F32_ATOMIC;
  e.alpha = 1.f;
  p.ep = e;
  CUtensorMap ta, tb;
  if (int rc = nhwc_tmap(&ta, dy, B, H, W, Cout, p.tw, p.th)) return rc;
  if (int rc = nhwc_tmap(&tb, x, B, H, W, C, p.tw, p.th)) return rc;
  dim3 grid(m_tiles, n_tiles, splits);
  return launch_tc(ta, ta, tb, p, grid, reinterpret_cast<cudaStream_t>(stream));
}

// Weight gradient of a per-token linear map with strided token views: dw[Cout][ldw] += sum_{b,t} dy[b][t][co] * x[b][t][ci]
// for t in [0, T), where dy / x are [B][T][.] views with arbitrary row and image strides (elements).  Used for the
// recurrent LSTM weights, where dy and x are the same activations shifted by one time step.  C <= 256 per call.
extern "C" int pe_wgrad_tokens(const void* dy, long long dy_ld, long long dy_img, const void* x, long long x_ld,
                               long long x_img, float* dw, long long ldw, int B, int T, int C, int Cout, int splits,
                               pe_stream_t stream) {
  if (int rc = pe_host::check_arch()) return rc;
  if (!dy || !x || !dw || B <= 0 || T <= 0 || C % 64 || C <= 0 || C > 256 || Cout % 8 || Cout <= 0)
    return PE_ERR_BAD_SHAPE;
  if ((dy_ld % 8) || (x_ld % 8) || (dy_img % 8) || (x_img % 8)) return PE_ERR_BAD_SHAPE;
  TcParams p{};
  p.mode = 2;
  p.kind = 0;
  p.a_mn = 1;
  p.b_mn = 1;
  p.H = T;
  p.W = 1;
  p.tw = 1;
  p.th = 64;
  p.tiles_w = 1;
  p.tiles_h = (T + 63) / 64;
  p.taps = 1;
  p.c1_chunks = C / 64;
  p.M = Cout;
  p.N = C;
  p.block_n = C;
  p.a_boxes = Cout > 64 ? 2 : 1;
  p.b_boxes = C / 64;
  p.kb_total = B * p.tiles_h;
  const int m_tiles = (Cout + 127) / 128;
  if (splits < 1) {
    splits = pe_host::num_sms() / m_tiles;
    if (splits < 1) splits = 1;
  }
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  pe_epilogue e{};
  e.out = dw;
  e.ldc = ldw;
  e.out_mode = PE_OUT_F32_ATOMIC;
  e.alpha = 1.f;
  p.ep = e;
  CUtensorMap ta, tb;
  {
    uint64_t dims[4] = {(uint64_t)Cout, 1, (uint64_t)T, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)dy_ld * 2, (uint64_t)dy_ld * 2, (uint64_t)dy_img * 2};
    uint32_t box[4] = {64, 1, 64, 1};
    if (int rc = pe_host::encode_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dy, dims, str, box)) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)C, 1, (uint64_t)T, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)x_ld * 2, (uint64_t)x_ld * 2, (uint64_t)x_img * 2};
    uint32_t box[4] = {64, 1, 64, 1};
    if (int rc = pe_host::encode_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, str, box)) return rc;
  }
  dim3 grid(m_tiles, 1, splits);
  return launch_tc(ta, ta, tb, p, grid, reinterpret_cast<cudaStream_t>(stream));
}
