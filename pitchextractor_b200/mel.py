"""Host-side constant tables and the batched GPU log-mel transform (mirror of the reference's
``MelDataset.to_melspec`` + normalisation, meldataset.py:77,644,650)."""
import math

import numpy as np
import torch

from . import ops

DEFAULT_MEL_PARAMS = {  # meldataset.py:34-40
    "sample_rate": 24000,
    "n_mels": 80,
    "n_fft": 1024,
    "win_length": 1024,
    "hop_length": 300,
}


def mel_filterbank(sample_rate, n_fft, n_mels, f_min=0.0, f_max=None):
    """HTK triangular filterbank [n_fft//2+1, n_mels], computed with the same fp32 torch ops (and in the same
    order) as torchaudio.functional.melscale_fbanks (norm=None, mel_scale='htk') so the table is bit-identical
    to the one the reference builds at meldataset.py:77."""
    f_max = float(sample_rate // 2) if f_max is None else f_max
    n_freqs = n_fft // 2 + 1
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).contiguous()


def windowed_dft_basis(n_fft, win_length=None):
    """fp32 [n_fft, ld] with basis[n, 2k] = w[n] cos(2 pi k n / N), basis[n, 2k+1] = w[n] sin(2 pi k n / N)
    (periodic Hann, torch.hann_window default), computed in fp64; ld = 2*(N/2+1) rounded up to 64."""
    win_length = n_fft if win_length is None else win_length
    n_bins = n_fft // 2 + 1
    w = np.zeros(n_fft)
    left = (n_fft - win_length) // 2  # torch.stft centres a short window inside n_fft
    w[left:left + win_length] = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)
    n = np.arange(n_fft)[:, None]
    k = np.arange(n_bins)[None, :]
    # reduce the phase index mod N in integers so the fp64 angle stays small and exact
    ang = 2.0 * np.pi * ((n * k) % n_fft) / n_fft
    ld = ((2 * n_bins + 63) // 64) * 64
    basis = np.zeros((n_fft, ld), dtype=np.float32)
    basis[:, 0:2 * n_bins:2] = (w[:, None] * np.cos(ang)).astype(np.float32)
    basis[:, 1:2 * n_bins:2] = (w[:, None] * np.sin(ang)).astype(np.float32)
    return torch.from_numpy(basis)


_TABLE_CACHE = {}


def logmel_tables(device, sample_rate=24000, n_fft=1024, win_length=1024, hop_length=300, n_mels=80, **_):
    """Immutable constant tables, built once per (device, params)."""
    key = (str(device), sample_rate, n_fft, win_length, hop_length, n_mels)
    if key not in _TABLE_CACHE:
        _TABLE_CACHE[key] = {
            "n_fft": n_fft, "hop": hop_length, "n_mels": n_mels, "sr": sample_rate,
            "basis": windowed_dft_basis(n_fft, win_length).to(device),
            "fb": mel_filterbank(sample_rate, n_fft, n_mels).to(device),
        }
    return _TABLE_CACHE[key]


class LogMel:
    """Batched replacement of ``(log(1e-5 + MelSpectrogram(wave)) + 4) / 4``: wave [B, L] fp32 (cuda) ->
    [B, n_mels, T] with T = 1 + L // hop.  CUDA-only."""

    def __init__(self, device="cuda", **mel_params):
        params = dict(DEFAULT_MEL_PARAMS)
        params.update(mel_params)
        if "win_len" in params and "win_length" not in mel_params:
            params["win_length"] = params.pop("win_len")
        params.pop("win_len", None)
        self.params = params
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pitchextractor_b200.LogMel runs on CUDA (sm_100) only; there is no CPU path")
        self.tables = logmel_tables(self.device, **params)
        self._ws = None

    def num_frames(self, num_samples):
        return 1 + num_samples // self.params["hop_length"]

    def __call__(self, wave, crop=None, T_out=0, layout="bmt"):
        if wave.dim() == 1:
            wave = wave[None]
        wave = wave.to(self.device, torch.float32).contiguous()
        B, Lw = wave.shape
        T = self.num_frames(Lw)
        To = T_out if T_out > 0 else T
        n_mels = self.params["n_mels"]
        shape = (B, n_mels, To) if layout == "bmt" else (B, To, n_mels)
        out = torch.empty(shape, device=self.device, dtype=torch.float32)
        need = B * T * (self.params["n_fft"] // 2 + 1)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, device=self.device, dtype=torch.float32)
        if crop is not None:
            crop = crop.to(self.device, torch.int32).contiguous()
        ops.logmel(wave, self.tables, out_bmt=out if layout == "bmt" else None,
                   out_btm=out if layout == "btm" else None, crop=crop, T_out=To, power_ws=self._ws)
        return out
