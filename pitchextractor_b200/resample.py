"""GPU resampling with the arithmetic of ``torchaudio.functional.resample``'s default
``sinc_interp_hann`` method, which the reference applies per file on the CPU (meldataset.py:621-627).

The polyphase filter bank is built on the host following torchaudio's published construction
(torchaudio/functional/functional.py ``_get_sinc_resample_kernel``: Hann-windowed sinc, ``lowpass_filter_width`` 6,
``rolloff`` 0.99, evaluated in float64 and rounded to float32); the convolution runs in ``pe_resample_f32``.
"""
import ctypes
import math

import numpy as np
import torch

from ._lib import call, ptr, stream

_BANKS = {}


def sinc_filter_bank(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """-> (h float32 [up][2*width + down], up, down, width) for orig_freq -> new_freq."""
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq or orig_freq <= 0 or new_freq <= 0:
        raise ValueError("resampling needs positive integer sample rates")
    g = math.gcd(int(orig_freq), int(new_freq))
    down, up = int(orig_freq) // g, int(new_freq) // g
    base = min(down, up) * rolloff
    width = math.ceil(lowpass_filter_width * down / base)
    idx = np.arange(-width, width + down, dtype=np.float64)[None, :] / down
    # torchaudio divides the (integer) phase offsets in float32 before adding the float64 tap grid; follow it bit for bit
    phase = (np.arange(0, -up, -1).astype(np.float32) / np.float32(up)).astype(np.float64)
    t = (phase[:, None] + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2.0) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(t == 0, 1.0, np.sin(t) / t)
    h = sinc * window * (base / down)
    return torch.from_numpy(h.astype(np.float32)), up, down, width


def resample(wave, orig_freq, new_freq, lengths=None):
    """wave [B, L] (or [L]) fp32 -> [B, ceil(L * new / orig)] on the GPU; identity when the rates agree.
    lengths: optional int32 [B] valid samples per (zero-padded) item: samples past an item's end are treated as zeros."""
    if int(orig_freq) == int(new_freq):
        return wave
    squeeze = wave.dim() == 1
    if squeeze:
        wave = wave[None]
    if not wave.is_cuda:
        raise RuntimeError("pitchextractor_b200.resample runs on CUDA (sm_100) only; move the waveform to the GPU")
    wave = wave.to(torch.float32).contiguous()
    key = (int(orig_freq), int(new_freq), str(wave.device))
    if key not in _BANKS:
        h, up, down, width = sinc_filter_bank(orig_freq, new_freq)
        _BANKS[key] = (h.to(wave.device), up, down, width)
    h, up, down, width = _BANKS[key]
    B, L = wave.shape
    L_out = int(math.ceil(up * L / down))
    out = torch.empty(B, L_out, device=wave.device, dtype=torch.float32)
    if lengths is not None:
        lengths = lengths.to(wave.device, torch.int32).contiguous()
    call("pe_resample_f32", ptr(wave), ctypes.c_longlong(wave.stride(0)), ptr(lengths), ctypes.c_int(B), ctypes.c_int(L),
         ptr(h), ctypes.c_int(up), ctypes.c_int(down), ctypes.c_int(width), ptr(out), ctypes.c_longlong(out.stride(0)),
         ctypes.c_int(L_out), stream())
    return out[0] if squeeze else out
