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

/* ------------------------------------------------------------------------------------------------
 * Fused epilogue description shared by the tensor-core GEMM / implicit-GEMM convolution.
 *   v = acc (+ bias[col]);  if act==GELU {out2 = bf16(v); v = gelu(v)};  if drop_thresh {v = keep ? v*drop_scale : 0};
 *   if aux_mode==ADD v += aux; if aux_mode==GELU_GRAD v *= gelu'(aux);  store (fp32 | bf16 | atomic fp32 add).
 * ------------------------------------------------------------------------------------------------ */
enum { PE_OUT_F32 = 0, PE_OUT_BF16 = 1, PE_OUT_F32_ATOMIC = 2 };
enum { PE_ACT_NONE = 0, PE_ACT_GELU = 1 };
enum { PE_AUX_NONE = 0, PE_AUX_ADD = 1, PE_AUX_GELU_GRAD = 2 };

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
  float* stats;        /* optional [2][N] fp32: per-column sum and sum of squares of the STORED value (atomic) */
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
 * ------------------------------------------------------------------------------------------------ */
int pe_logmel_f32(const float* wave, int B, int L, int n_fft, int hop, int n_mels, const float* basis,
                  int ld_basis, const float* fb, float* power, size_t power_bytes, float* out_bmt, float* out_btm,
                  const int* crop, int T_out, pe_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PITCHEXTRACTOR_B200_H */
