/* pitchextractor_b200 -- C-ABI of the B200 (sm_100a) hot-path kernels.
 *
 * The reference (martinambrus/PitchExtractor) is pure Python on top of torch / torchaudio and has no FFI of
 * its own; its "operator interface" for this path is the set of library calls made by
 *   meldataset.py:77,644,650   torchaudio MelSpectrogram + log/normalise           -> pe_logmel_*
 *   model.py:23-57,143-175     Conv2d / BatchNorm2d / LeakyReLU / MaxPool2d / Dropout -> pe_conv_*, pe_bn_*, pe_stem_*
 *   model.py:196-256           nn.LSTM / nn.TransformerEncoder / LayerNorm / Linear -> pe_gemm_bf16, pe_attn_*, pe_ln_*, pe_lstm_*
 *   trainer.py:237-239         SmoothL1Loss / BCEWithLogitsLoss on the two heads    -> pe_heads_loss_*
 *   optimizers.py:54-64        AdamW                                                -> pe_adamw
 * Each entry point below replaces one of those call sites (cited again next to the declaration).
 *
 * Conventions: plain pointers and sizes only (no torch types); every pointer is DEVICE memory owned by the
 * caller; functions never allocate, free or synchronise; work is enqueued on `stream`.  Return value: 0 on
 * success, negative PE_ERR_* otherwise (bad shape, wrong architecture, driver entry point missing, launch error).
 * There is no CPU fallback: on a machine without an sm_100 device every compute call returns PE_ERR_ARCH.
 */
#ifndef PITCHEXTRACTOR_B200_H
#define PITCHEXTRACTOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pe_stream_t; /* == cudaStream_t */

#define PE_OK 0
#define PE_ERR_BAD_SHAPE (-1)
#define PE_ERR_WORKSPACE (-2)
#define PE_ERR_ARCH (-3)
#define PE_ERR_DRIVER (-4)
#define PE_ERR_LAUNCH (-5)

int pe_version(void);
int pe_check_device(void); /* PE_OK iff the current device is sm_100 */
/* Bytes of caller-provided workspace entry point `op` (its name, e.g. "pe_lstm_seq_fwd") needs for a batch of B items:
 * pe_lstm_seq_fwd / _bwd (B, T): arrival counters;  pe_logmel_tc (B, L samples): the re-strided waveform copy;
 * pe_logmel_f32 (B, T frames, L = n_fft): the power spectrogram.  0 for entry points that need none, -1 on bad arguments. */
long long pe_workspace_bytes(const char* op, int B, int T, int L);

/* Per-step dropout salt: every dropout site (reference model.py:40,56 nn.Dropout, the Transformer / LSTM dropouts of
 * model.py:306-341) draws its mask from (seed argument + salt[slot]), slot = the top 8 bits of the seed argument.  The
 * salt lives in device memory so that a training step captured once into a CUDA graph -- whose seed arguments are
 * frozen -- still sees fresh masks on every replay: the host bumps the salt of ITS slot (stream-ordered, one 1-block
 * kernel) before each replay.  Every model instance uses its own slot (0..255), so the library holds no state that two
 * callers share; eager launches keep their slot at 0. */
int pe_set_step_salt(int slot, unsigned long long salt, pe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused epilogue description shared by the tensor-core GEMM / implicit-GEMM convolution.
 *   v = acc (+ bias[col]);  if act==GELU {out2 = bf16(v); v = gelu(v)};  if drop_thresh {v = keep ? v*drop_scale : 0};
 *   (act==GELU_SAVE_GRAD: out2 = bf16(gelu'(v) * dropout factor) instead, i.e. d out / d v, so that the backward GEMM
 *   only multiplies: aux_mode==MUL)
 *   if aux_mode==ADD v += aux; if aux_mode==GELU_GRAD v *= gelu'(aux); if aux_mode==MUL v *= aux;
 *   store (fp32 | bf16 | atomic fp32 add).
 * ------------------------------------------------------------------------------------------------ */
enum { PE_OUT_F32 = 0, PE_OUT_BF16 = 1, PE_OUT_F32_ATOMIC = 2 };
enum { PE_ACT_NONE = 0, PE_ACT_GELU = 1, PE_ACT_GELU_SAVE_GRAD = 2 };
enum { PE_AUX_NONE = 0, PE_AUX_ADD = 1, PE_AUX_GELU_GRAD = 2, PE_AUX_MUL = 3 };

typedef struct pe_epilogue {
  void* out;           /* [M][ldc] */
  long long ldc;       /* elements */
  int out_mode;        /* PE_OUT_* */
  int act;             /* PE_ACT_* */
  void* out2;          /* bf16 [M][ld2], pre-activation copy when act == GELU (may be NULL) */
  long long ld2;
  const float* bias;   /* [N] fp32 or NULL */
  const void* aux;     /* bf16 [M][ld_aux] or NULL */
  long long ld_aux;
  int aux_mode;        /* PE_AUX_* */
  unsigned int drop_thresh;      /* keep iff philox(seed, row*N+col) < drop_thresh; 0 disables dropout */
  float drop_scale;              /* 1/keep_prob */
  unsigned long long drop_seed;
  float alpha;                   /* scale applied to the accumulator first (1.0f for none) */
  /* optional per-column reductions of the STORED value v over all rows (N <= 256), accumulated atomically in fp64:
   *   stats_mode 1: stats[0][n] += sum v,  stats[1][n] += sum v^2            (BatchNorm batch statistics)
   *   stats_mode 2: g = v * lrelu'(x*scale + shift); stats[0][n] += sum g, stats[1][n] += sum g*x
   *                 (first pass of the BatchNorm backward; x = stats_x bf16 [M][N], the BN input)
   *   stats_mode 3: the same through a MaxPool2d((1,2)) between the BN and this gradient (model.py:149-153):
   *                 x = stats_x bf16 [2M][N]; row m owns x rows 2m, 2m+1 and routes g to the first maximum of
   *                 lrelu(x*scale + shift) */
  double* stats;
  int stats_mode;
  const void* stats_x;
  const float* stats_scale;
  const float* stats_shift;
  float stats_slope;
  /* tuning aid (NULL in production): per-CTA counters [grid][16] = {main-loop cycles, MMA thread waiting for operands,
   * MMA thread waiting for a free accumulator stage, TMA thread waiting for a free smem slot, CTA entry time
   * (globaltimer ns), cycles from entry to: set-up done, MMA loop end, CTA exit, ...} */
  long long* debug;
} pe_epilogue;

/* D[M,N] = sum_k A(m,k) * B(n,k), bf16 operands, fp32 accumulation in TMEM (tcgen05.mma kind::f16).
 *   a_mn == 0: A is [M][lda] (K contiguous)      a_mn == 1: A is [K][lda] (M contiguous)
 *   b_mn == 0: B is [N][ldb] (K contiguous)      b_mn == 1: B is [K][ldb] (N contiguous)
 * splits > 1 partitions K over gridDim.z and requires out_mode == PE_OUT_F32_ATOMIC.
 * Replaces torch.nn.functional.linear and its two backward GEMMs (model.py:67,70,220-240 via torch). */
int pe_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, int M, int N, int K,
                 const pe_epilogue* ep, int splits, pe_stream_t stream);

/* 3x3 / pad 1 / stride 1 convolution as an implicit GEMM over NHWC bf16 activations (model.py:27,157-161),
 * optionally fused with a 1x1 shortcut convolution on a second input accumulated into the same tile
 * (model.py:167,172).  x: [B][H][W][C1], x2: [B][H][W][C2] or NULL, w: [Cout][9*C1 + C2] bf16 (tap-major,
 * channel-minor; shortcut columns last), y: [B][H][W][Cout].  C1, C2 multiples of 64.  The same entry point
 * computes the data gradient when given the flipped / transposed weights. */
int pe_conv3x3_nhwc(const void* x, const void* x2, const void* w, int B, int H, int W, int C1, int C2, int Cout,
                    const pe_epilogue* ep, pe_stream_t stream);

/* Weight gradient of the convolution above: dw[Cout][taps*C + ...] += sum_pixels dy[p][co] * x[p+tap][ci].
 * dy: [B][H][W][Cout] bf16, x: [B][H][W][C] bf16, dw: fp32 [Cout][ldw] at column offset tap*C (3x3, taps=9) or
 * a single centre tap (taps=1, the 1x1 shortcut).  Accumulates atomically (caller zeroes dw). */
int pe_conv_wgrad_nhwc(const void* dy, const void* x, float* dw, long long ldw, int B, int H, int W, int C, int Cout,
                       int taps, int splits, pe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * log-mel front-end (meldataset.py:77,644,650 / torchaudio MelSpectrogram):
 * wave [B][L] fp32 -> y[b][m][t] = (log(1e-5 + mel[b][m][t]) + 4) / 4.
 *   basis  : fp32 [n_fft][ld_basis] windowed DFT basis (w[n]cos, w[n]sin interleaved per bin), zero padded to
 *            ld_basis = round_up(2*(n_fft/2+1), 64) columns
 *   fb     : fp32 [n_fft/2+1][n_mels] mel filterbank
 *   power  : fp32 workspace [B*T][n_fft/2+1]
 *   out_bmt: [B][n_mels][T_out] (reference layout) or NULL;  out_btm: [B][T_out][n_mels] (model layout) or NULL
 *   crop   : int32 [B] first frame kept per item (NULL = 0); T_out frames are written (zero-padded past T).
 *   lengths: int32 [B] valid samples per item (NULL = L): the reflect padding happens at the item's OWN end and frames
 *            t >= 1 + lengths[b] / hop are written as 0 (what Collater pads short items with, meldataset.py:804-816).
 * ------------------------------------------------------------------------------------------------ */
int pe_logmel_f32(const float* wave, int B, int L, int n_fft, int hop, int n_mels, const float* basis,
                  int ld_basis, const float* fb, float* power, size_t power_bytes, float* out_bmt, float* out_btm,
                  const int* crop, const int* lengths, int T_out, pe_stream_t stream);
/* Same transform on tcgen05 tensor cores for n_fft == 1024, hop % 4 == 0, hop <= 320: two frames per complex FFT,
 * 32 x 32 four-step factorisation as two [128 x 64] x [64 x 64] fp16-split GEMM stages per 8 frames (fp32-grade
 * accuracy); every 8-frame slot reads its 1024 + 7*hop samples once.
 *   win [1024] fp32 window x 2^11;  fmat: 2 x [64][64] fp16 operand images (hi, lo) of the 32-point complex DFT matrix
 *   in the 128-byte-swizzled K-major layout;  tw [2][32][32] fp32 cos / sin of 2 pi k1 n2 / 1024, x 2^-5;
 *   mel_w [mel_nnz <= 1536]: the non-zero filterbank weights x 2^-14, band after band (the power-of-two scales keep
 *   the fp16 split of the data in range and cancel exactly);  mel_items int32 [128][4]: the work items
 *   of the banded mel product, one per worker lane: {filter (-1: idle), first bin, number of bins, offset into mel_w |
 *   flags << 24}, flags 1 = add the partial sums of lane ^ 1 (a long filter split over two adjacent lanes), 2 = this lane
 *   writes the filter (n_mels <= 128);
 *   xpad: fp32 workspace [B][round_up(L, 4)], used only when L % 4 != 0 (rows are re-strided to 16 bytes); the reflect
 *   padding happens inside the kernel.  Only frames crop[b] .. crop[b] + T_out - 1 are computed; ALL T_out output rows of
 *   every item are written (zeros past the item's end, see `lengths` above). */
int pe_logmel_tc(const float* wave, int B, int L, int n_fft, int hop, int n_mels, const float* win, const void* fmat,
                 const float* tw, const int* mel_items, const float* mel_w, int mel_nnz, float* xpad, size_t xpad_bytes,
                 float* out_bmt, float* out_btm, const int* crop, const int* lengths, int T_out, pe_stream_t stream);

/* Polyphase windowed-sinc resampling (torchaudio.functional.resample, called per file at meldataset.py:621-627):
 * y[b][j*up + p] = sum_k xpad[b][j*down + k] * h[p][k], xpad = x zero-padded by (width, width + down); h fp32
 * [up][2*width + down] is the filter bank of torchaudio's _get_sinc_resample_kernel for the reduced rates
 * up = new/gcd, down = orig/gcd.  x [B][ldx], y [B][ldy] with L_out = ceil(L*up/down) samples written per item;
 * lengths (optional int32 [B]): samples at or past an item's length count as zeros. */
int pe_resample_f32(const float* x, long long ldx, const int* lengths, int B, int L, const float* h, int up, int down,
                    int width, float* y, long long ldy, int L_out, pe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Conv trunk, memory-bound passes over NHWC bf16 activations (model.py:23-57,143-175).
 * ------------------------------------------------------------------------------------------------ */
/* Conv2d(1->64, 3x3, pad 1, no bias) (model.py:24): x[b][t][f] fp32 with element strides (sb, st, sf) -> y bf16
 * [B][T][F][64]; w fp32 [64][9].  stats (optional, fp64 [2][64], caller zeroes): the batch statistics of the following
 * BatchNorm2d (model.py:25) -- sum y, sum y^2 of the bf16-rounded output -- accumulated in the same pass. */
int pe_stem_conv_fwd(const float* x, long long sb, long long st, long long sf, int B, int T, int F, const float* w,
                     void* y, double* stats, pe_stream_t stream);
/* its weight gradient: dw[64][9] += sum_p dy[p][c] x[p+tap] */
int pe_stem_conv_wgrad(const float* x, long long sb, long long st, long long sf, int B, int T, int F, const void* dy,
                       float* dw, pe_stream_t stream);
/* nn.BatchNorm2d training statistics (model.py:25,37,54,150,159): sums[0][c] += sum x, sums[1][c] += sum x^2 over
 * x bf16 [rows][C]; caller zeroes sums (fp64 [2][C]). */
int pe_bn_stats(const void* x, long long rows, int C, double* sums, pe_stream_t stream);
/* batch mean / biased variance -> scale = gamma*rstd, shift = beta - mean*scale, saved mean / rstd; updates
 * running_mean / running_var (momentum, unbiased variance) and num_batches_tracked as torch does (may be NULL). */
int pe_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float eps, float momentum,
                   float* scale, float* shift, float* mean, float* rstd, float* running_mean, float* running_var,
                   long long* num_batches_tracked, int C, pe_stream_t stream);
/* eval mode: scale / shift from the running statistics */
int pe_bn_eval_params(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float eps, float* scale, float* shift, float* mean, float* rstd, int C, pe_stream_t stream);
/* y = Dropout(MaxPool2d((1,k))(LeakyReLU(x*scale + shift))) (model.py:36-41,149-153; scale == NULL: pure max-pool,
 * model.py:45-49).  x bf16 [rows][W][C]; output pixel (row, wo) written at out + (row*Wo+wo)*ld_out + c_off, and/or
 * in the sequence layout out_seq[row][c*Wo + wo] (model.py:93,112).  argmax_out (optional, uint8 [rows][Wo][C], k <= 255):
 * position of the (first) maximum inside each window, for pe_maxpool_bwd_add. */
int pe_bn_act_pool_fwd(const void* x, long long rows, int W, int C, int k, const float* scale, const float* shift,
                       float slope, unsigned drop_thresh, float drop_scale, unsigned long long seed, void* out,
                       long long ld_out, int c_off, void* out_seq, void* argmax_out, pe_stream_t stream);
/* backward of the block above through dropout, max-pool, LeakyReLU and the BatchNorm batch statistics:
 * dx bf16 [rows][W][C]; dgamma / dbeta accumulated (+=); sums is a zeroed fp64 [2][C] scratch, coef an fp32 [2][C]
 * scratch; k in {1, 2, 4}.  sums_ready != 0: the reduction pass already ran (fused in the producing convolution's
 * epilogue, pe_epilogue.stats_mode 2 / 3) and is skipped.  aux_argmax != NULL: x also feeds an auxiliary
 * MaxPool2d((1,aux_k)) (model.py:45-49,103-105) whose arg-max positions pe_bn_act_pool_fwd saved; its gradient
 * aux_dout (pixel (row, w/aux_k) at aux_dout + (row*(W/aux_k) + w/aux_k)*aux_ld + aux_c_off) is added to dx here. */
int pe_bn_act_pool_bwd(const void* x, long long rows, int W, int C, int k, const float* scale, const float* shift,
                       const float* mean, const float* rstd, float slope, unsigned drop_thresh, float drop_scale,
                       unsigned long long seed, const void* dout, long long ld_dout, int c_off, const void* dout_seq,
                       double* sums, int sums_ready, float* coef, float* dgamma, float* dbeta, const void* aux_argmax,
                       const void* aux_dout, long long aux_ld, int aux_c_off, int aux_k, void* dx, pe_stream_t stream);
/* backward of the auxiliary max-pools: dx[argmax of each window] += dout.  With argmax (saved by the forward pass) x is
 * not read; otherwise the arg-max is recomputed from x. */
int pe_maxpool_bwd_add(const void* x, const void* argmax, long long rows, int W, int C, int k, const void* dout,
                       long long ld_dout, int c_off, void* dx, pe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Sequence model passes (model.py:178-256; torch/nn/modules/transformer.py:961-982).
 * ------------------------------------------------------------------------------------------------ */
/* out = LayerNorm(x (+ pe[m % T])) over D in {256,512,768}; x is fp32 (x_f32) or bf16 (x_bf16); out bf16;
 * mean / rstd fp32 [M] saved for the backward pass. */
int pe_layernorm_fwd(const float* x_f32, const void* x_bf16, const float* pe_table, int T, int D, const float* gamma,
                     const float* beta, float eps, long long M, void* out, float* mean, float* rstd,
                     pe_stream_t stream);
/* dx (bf16) and, optionally, dx_masked = dropout-masked dx (gradient entering the preceding Linear);
 * dgamma, dbeta, dbias (= column sums of dx_masked) accumulated (+=). */
int pe_layernorm_bwd(const void* dy, const float* x_f32, const void* x_bf16, const float* pe_table, int T, int D,
                     const float* gamma, const float* mean, const float* rstd, long long M, void* dx, void* dx_masked,
                     unsigned drop_thresh, float drop_scale, unsigned long long seed, float* dgamma, float* dbeta,
                     float* dbias, pe_stream_t stream);
/* out[n] += sum_m x[m][n], x bf16 [M][ld] (bias gradients) */
int pe_colsum_bf16(const void* x, long long M, int N, long long ld, float* out, pe_stream_t stream);
/* multi-head softmax attention with dropout on the probabilities (nn.MultiheadAttention inside
 * nn.TransformerEncoderLayer, model.py:231-239): qkv bf16 [B*T][3*H*64] -> ctx bf16 [B*T][H*64], lse fp32 [B][H][T] */
/* T == 192 runs on the tcgen05 kernels, other lengths on the fp32 SIMT kernels; force_simt != 0 selects the SIMT kernels
 * at T == 192 too (cross-checks). */
int pe_attn_fwd(const void* qkv, int B, int T, int H, int head_dim, unsigned drop_thresh, float drop_scale,
                unsigned long long seed, void* ctx, float* lse, int force_simt, pe_stream_t stream);
int pe_attn_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse, int B, int T, int H,
                int head_dim, unsigned drop_thresh, float drop_scale, unsigned long long seed, void* dqkv, float* delta,
                int force_simt, pe_stream_t stream);
/* heads (num_class == 1) + losses (model.py:96-98,115-117; train.py:104-106; trainer.py:237-239):
 * f0 = hc.wc + bc, logit = hd.(wd[0]+wd[1]) + bd[0]+bd[1]; loss_out = {total, lambda*SmoothL1, BCE}.
 * When dhc != NULL also writes dL/dhc, dL/dhd (bf16) and accumulates the head parameter gradients; with gc_ext /
 * gd_ext (fp32 [M]) the output gradients are supplied by the caller instead of derived from the losses. */
int pe_heads_loss(const void* hc, const void* hd, long long M, int D, const float* wc, const float* bc, const float* wd,
                  const float* bd, const float* f0_target, const float* sil_target, float lambda_f0, float grad_scale,
                  float* f0_pred, float* sil_logit, double* loss_acc, float* loss_out, const float* gc_ext,
                  const float* gd_ext, void* dhc, void* dhd, float* dwc, float* dbc, float* dwd, float* dbd,
                  pe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * BiLSTM sequence model (model.py:218-228 -> torch.nn.LSTM; gates i,f,g,o, torch/nn/modules/rnn.py:842-847).
 * One call advances the four recurrences of a layer (2 sequence models x 2 directions) through time steps
 * [step_begin, step_end) (one dependent launch per step); hidden size 384.  Arrays indexed [model] (2 entries) or
 * [model*2 + direction] (4 entries); all pointers device memory.  Token tensors are TIME-MAJOR (row = t * B + b): a
 * time step then touches one contiguous slab of B rows.
 *   gx : fp32 [T][B][2*1536]  input projection x W_ih^T on entry, activated gates (kept for the backward) on exit
 *   c  : fp32 [T][B][2*384]   cell state;   y : bf16 [T][B][2*384] hidden state (direction d at columns d*384)
 * ------------------------------------------------------------------------------------------------ */
int pe_lstm_steps_fwd(int B, int T, int hidden, int step_begin, int step_end, float* const* gx, float* const* c,
                      void* const* y, const void* const* w_hh /* bf16 [1536][384] */, const float* const* b_ih,
                      const float* const* b_hh, pe_stream_t stream);
/* backward steps (step 0 = the last forward step of each direction): dg bf16 [T][B][2*1536] receives the
 * pre-activation gate gradients, dc fp32 [B][2*384] carries dL/dc between calls, dy bf16 [T][B][2*384] is dL/dy. */
int pe_lstm_steps_bwd(int B, int T, int hidden, int step_begin, int step_end, const float* const* gates,
                      const float* const* c, const void* const* dy, void* const* dg, float* const* dc,
                      const void* const* w_hh, pe_stream_t stream);
/* The same recurrences as ONE persistent launch per layer over all T steps (north_star: "persistent kernel with its
 * weights resident in shared memory"): every CTA keeps its 192 KB slice of W_hh in shared memory, h_t / dgates_t are
 * exchanged through L2 with a per-step arrival counter.  Same tensors as above; the cell-state gradient is carried in
 * registers.  workspace: pe_workspace_bytes("pe_lstm_seq_fwd" / "pe_lstm_seq_bwd", B, T, 0) bytes of device memory: the
 * arrival counters (zeroed by the call) and, for the backward, room for a transposed copy of the four W_hh (written by
 * the call; used by the gate-stacked kernel of narrow batch tiles -- with a workspace that only holds the counters the
 * call falls back to the plain kernel). */
int pe_lstm_seq_fwd(int B, int T, int hidden, float* const* gx, float* const* c, void* const* y,
                    const void* const* w_hh, const float* const* b_ih, const float* const* b_hh, void* workspace,
                    size_t workspace_bytes, pe_stream_t stream);
int pe_lstm_seq_bwd(int B, int T, int hidden, const float* const* gates, const float* const* c, const void* const* dy,
                    void* const* dg, const void* const* w_hh, void* workspace, size_t workspace_bytes,
                    pe_stream_t stream);
/* y = dropout(x) on n bf16 elements (n % 8 == 0); the same call with the same seed back-propagates */
int pe_dropout_bf16(const void* x, void* y, long long n, unsigned drop_thresh, float drop_scale,
                    unsigned long long seed, pe_stream_t stream);
/* dw[Cout][ldw] += sum_{b,t} dy[b][t][co] x[b][t][ci] over strided [B][T][.] token views (row / image strides in
 * elements); the recurrent weight gradient is this with dy and x shifted by one time step.  C <= 256. */
int pe_wgrad_tokens(const void* dy, long long dy_ld, long long dy_img, const void* x, long long x_ld, long long x_img,
                    float* dw, long long ldw, int B, int T, int C, int Cout, int splits, pe_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Parameter passes (optimizers.py:54-64).
 * ------------------------------------------------------------------------------------------------ */
/* torch.optim.AdamW step over a flat fp32 arena; also writes the bf16 working copy when p_bf16 != NULL. */
int pe_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
             float weight_decay, long long step, float grad_scale, void* p_bf16, pe_stream_t stream);
int pe_cast_bf16(const float* x, void* y, long long n, pe_stream_t stream);
/* tensor-core operand layouts of a conv weight w fp32 [Cout][9][Cin] (+ 1x1 shortcut w2 [Cout][C2] or, for the data
 * gradient, [C2][Cin]): w_fwd bf16 [Cout][9*Cin + C2]; w_dgrad bf16 [Cin][9*Cout + C2] with flipped taps. */
int pe_conv_weight_prep(const float* w, int Cout, int Cin, const float* w2, int C2, void* w_fwd, void* w_dgrad,
                        pe_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PITCHEXTRACTOR_B200_H */
