"""fp64 numpy restatement of torchaudio.functional.resample (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99),
the call the reference makes at meldataset.py:621-627 (oracle; test infrastructure only).

Follows torchaudio/functional/functional.py: ``_get_sinc_resample_kernel`` (filter bank) and
``_apply_sinc_resample_kernel`` (zero pad by (width, width + orig), strided correlation, crop to ceil(new * L / orig)).
Pinned against torchaudio itself in tests/test_resample_cache.py.
"""
import math

import numpy as np


def filter_bank(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)  # torchaudio: float32 here
    t = np.clip((phase[:, None] + idx) * base, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2.0) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        kern = np.where(t == 0, 1.0, np.sin(t) / t) * window * (base / orig)
    return kern.astype(np.float32).astype(np.float64), orig, new, width


def resample(x, orig_freq, new_freq):
    """x [L] -> [ceil(L * new / orig)] (float64 accumulation of the float32 filter taps)."""
    if int(orig_freq) == int(new_freq):
        return np.asarray(x, dtype=np.float64)
    kern, orig, new, width = filter_bank(orig_freq, new_freq)
    x = np.asarray(x, dtype=np.float64)
    L = x.shape[0]
    xp = np.concatenate([np.zeros(width), x, np.zeros(width + orig)])
    K = kern.shape[1]
    n_blocks = (xp.shape[0] - K) // orig + 1
    out = np.empty((n_blocks, new))
    for j in range(n_blocks):
        out[j] = kern @ xp[j * orig:j * orig + K]
    return out.reshape(-1)[:int(math.ceil(new * L / orig))]
