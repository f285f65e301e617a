"""Thin Python wrappers over the C-ABI kernels: argument marshalling only (no arithmetic on this side)."""
import ctypes

import torch

from . import _lib as L
from ._lib import Epilogue, call, ptr, stream

c_int, c_ll, c_size = ctypes.c_int, ctypes.c_longlong, ctypes.c_size_t

# When set to a list, every tensor-core tile-engine launch appends (tag, start_event, end_event, algorithmic_flops):
# bench.py uses it to time the dominant kernel with CUDA events on the launching stream.
PROFILE = None


class _Timed:
    def __init__(self, tag, flops):
        self.tag, self.flops = tag, flops

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.tag, self.e0, self.e1, self.flops))


DEBUG_BUFFER = None  # tuning aid: int64 [grid][16] tensor the next tile-engine launches fill with per-CTA counters


def make_epilogue(out, out_mode=None, bias=None, act=L.PE_ACT_NONE, out2=None, aux=None, aux_mode=L.PE_AUX_NONE,
                  p_drop=0.0, seed=0, alpha=1.0, ldc=None, stats=None, stats_mode=0, stats_x=None, stats_scale=None,
                  stats_shift=None, stats_slope=0.01):
    ep = Epilogue()
    if DEBUG_BUFFER is not None:
        ep.debug = DEBUG_BUFFER.data_ptr()
    ep.out = out.data_ptr()
    ep.ldc = out.stride(-2) if ldc is None else ldc
    if out_mode is None:
        out_mode = L.PE_OUT_BF16 if out.dtype == torch.bfloat16 else L.PE_OUT_F32
    ep.out_mode = out_mode
    ep.act = act
    if out2 is not None:
        ep.out2, ep.ld2 = out2.data_ptr(), out2.stride(-2)
    if bias is not None:
        assert bias.dtype == torch.float32
        ep.bias = bias.data_ptr()
    if aux is not None:
        assert aux.dtype == torch.bfloat16
        ep.aux, ep.ld_aux, ep.aux_mode = aux.data_ptr(), aux.stride(-2), aux_mode
    ep.drop_thresh, ep.drop_scale = L.drop_thresh(p_drop)
    ep.drop_seed = seed
    ep.alpha = alpha
    if stats is not None:
        assert stats.dtype == torch.float64
        ep.stats, ep.stats_mode = stats.data_ptr(), (stats_mode or 1)
        if ep.stats_mode >= 2:
            ep.stats_x, ep.stats_scale, ep.stats_shift = stats_x.data_ptr(), stats_scale.data_ptr(), stats_shift.data_ptr()
            ep.stats_slope = stats_slope
    return ep


def gemm(a, b, out, M, N, K, a_mn=False, b_mn=False, splits=1, **epkw):
    """out[M,N] (op)= sum_k A(m,k) B(n,k).  a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N] if b_mn), bf16."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    ep = make_epilogue(out, **epkw)
    with _Timed("gemm", 2.0 * M * N * K):
        call("pe_gemm_bf16", ptr(a), c_ll(a.stride(0)), c_int(int(a_mn)), ptr(b), c_ll(b.stride(0)), c_int(int(b_mn)),
             c_int(M), c_int(N), c_int(K), ctypes.byref(ep), c_int(splits), stream())
    return out


def conv3x3(x, w, out, x2=None, **epkw):
    """x [B,H,W,C1] bf16 NHWC, x2 [B,H,W,C2] or None, w [Cout, 9*C1+C2] bf16, out [B,H,W,Cout]."""
    B, H, W, C1 = x.shape
    C2 = 0 if x2 is None else x2.shape[-1]
    Cout = w.shape[0]
    assert w.shape[1] == 9 * C1 + C2 and x.is_contiguous() and w.is_contiguous()
    ep = make_epilogue(out, ldc=Cout, **epkw)
    with _Timed("conv", 2.0 * B * H * W * Cout * (9 * C1 + C2)):
        call("pe_conv3x3_nhwc", ptr(x), ptr(x2), ptr(w), c_int(B), c_int(H), c_int(W), c_int(C1), c_int(C2),
             c_int(Cout), ctypes.byref(ep), stream())
    return out


def conv_wgrad(dy, x, dw, taps=9, splits=0, col_offset=0):
    """dw[Cout, col_offset + tap*C + ci] += sum_p dy[p,co] x[p+tap,ci]; dw fp32 [Cout, ldw]."""
    B, H, W, Cout = dy.shape
    C = x.shape[-1]
    base = ctypes.c_void_p(dw.data_ptr() + 4 * col_offset)
    with _Timed("wgrad", 2.0 * B * H * W * Cout * C * taps):
        call("pe_conv_wgrad_nhwc", ptr(dy), ptr(x), base, c_ll(dw.stride(0)), c_int(B), c_int(H), c_int(W), c_int(C),
             c_int(Cout), c_int(taps), c_int(splits), stream())
    return dw


def logmel(wave, tables, out_bmt=None, out_btm=None, crop=None, T_out=0, power_ws=None, lengths=None):
    """wave [B,L] fp32 cuda -> normalised log-mel; tables from ``logmel_tables``."""
    B, Lw = wave.shape
    n_fft, hop, n_mels = tables["n_fft"], tables["hop"], tables["n_mels"]
    T = 1 + Lw // hop
    n_bins = n_fft // 2 + 1
    if power_ws is None or power_ws.numel() < B * T * n_bins:
        power_ws = torch.empty(B * T * n_bins, device=wave.device, dtype=torch.float32)
    call("pe_logmel_f32", ptr(wave), c_int(B), c_int(Lw), c_int(n_fft), c_int(hop), c_int(n_mels),
         ptr(tables["basis"]), c_int(tables["basis"].stride(0)), ptr(tables["fb"]), ptr(power_ws),
         c_size(power_ws.numel() * 4), ptr(out_bmt), ptr(out_btm), ptr(crop), ptr(lengths), c_int(T_out), stream())
    return power_ws


def attn_fwd(qkv, B, T, H, ctx, lse, p_drop=0.0, seed=0, force_simt=False):
    th, sc = L.attn_drop_thresh(p_drop)
    call("pe_attn_fwd", ptr(qkv), c_int(B), c_int(T), c_int(H), c_int(64), ctypes.c_uint(th), ctypes.c_float(sc),
         ctypes.c_ulonglong(seed), ptr(ctx), ptr(lse), c_int(int(force_simt)), stream())


def attn_bwd(qkv, ctx, dctx, lse, B, T, H, dqkv, delta, p_drop=0.0, seed=0, force_simt=False):
    th, sc = L.attn_drop_thresh(p_drop)
    call("pe_attn_bwd", ptr(qkv), ptr(ctx), ptr(dctx), ptr(lse), c_int(B), c_int(T), c_int(H), c_int(64),
         ctypes.c_uint(th), ctypes.c_float(sc), ctypes.c_ulonglong(seed), ptr(dqkv), ptr(delta), c_int(int(force_simt)),
         stream())


def colsum(x, out):
    M, N = x.shape
    call("pe_colsum_bf16", ptr(x), c_ll(M), c_int(N), c_ll(x.stride(0)), ptr(out), stream())


def logmel_tc(wave, tables, ws, out_bmt=None, out_btm=None, crop=None, T_out=0, lengths=None):
    """tcgen05 four-step log-mel (n_fft 1024); ws: fp32 workspace for the reflect-padded waveform."""
    B, Lw = wave.shape
    t = tables
    call("pe_logmel_tc", ptr(wave), c_int(B), c_int(Lw), c_int(t["n_fft"]), c_int(t["hop"]), c_int(t["n_mels"]),
         ptr(t["win"]), ptr(t["fmat"]), ptr(t["tw"]), ptr(t["mel_items"]), ptr(t["mel_w"]), c_int(t["mel_nnz"]), ptr(ws),
         c_size(ws.numel() * 4), ptr(out_bmt), ptr(out_btm), ptr(crop), ptr(lengths), c_int(T_out), stream())
