"""fp64 numpy restatement of the reference log-mel front-end (oracle; test infrastructure only).

Follows, line by line:
  * ``torchaudio.transforms.MelSpectrogram(sample_rate=24000, n_fft=1024, win_length=1024,
    hop_length=300, n_mels=80)`` as constructed at reference ``meldataset.py:77`` (defaults
    f_min=0, f_max=sr/2, power=2, center=True, pad_mode="reflect", norm=None, mel_scale="htk",
    periodic Hann) -> ``torchaudio/functional/functional.py:122-144`` (stft, |.|^2) and
    ``torchaudio/functional/functional.py:492-572`` (melscale_fbanks).
  * ``(log(1e-5 + mel) - mean) / std`` with mean, std = -4, 4 -- ``meldataset.py:111,650``.
  * ``F0Extractor.align_length`` -- ``f0_backends.py:788-806``.
  * the label / crop logic of ``MelDataset._build_training_example`` -- ``meldataset.py:652-677``.
  * ``Collater.__call__`` -- ``meldataset.py:804-826``.
"""
import numpy as np

SR, N_FFT, HOP, N_MELS = 24000, 1024, 300, 80
LOG_EPS, MEAN, STD = 1e-5, -4.0, 4.0
MAX_MEL_LENGTH = 192


def hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_to_hz_htk(m):
    return 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def melscale_fbanks(n_freqs=N_FFT // 2 + 1, f_min=0.0, f_max=SR / 2.0, n_mels=N_MELS, sr=SR):
    """torchaudio/functional/functional.py:544-572 (norm=None, mel_scale='htk'), in fp64."""
    all_freqs = np.linspace(0.0, sr // 2, n_freqs)
    m_pts = np.linspace(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = mel_to_hz_htk(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))  # [n_freqs, n_mels]


def hann_periodic(n=N_FFT):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def num_frames(num_samples, hop=HOP):
    return 1 + num_samples // hop


def power_spectrogram(wave, n_fft=N_FFT, hop=HOP):
    """wave [L] -> power [T, n_fft//2+1] in fp64 (center=True, reflect pad)."""
    x = np.asarray(wave, dtype=np.float64)
    xp = np.pad(x, (n_fft // 2, n_fft // 2), mode="reflect")
    T = num_frames(x.shape[0], hop)
    idx = hop * np.arange(T)[:, None] + np.arange(n_fft)[None, :]
    frames = xp[idx] * hann_periodic(n_fft)[None, :]
    spec = np.fft.rfft(frames, axis=-1)
    return spec.real ** 2 + spec.imag ** 2


def mel_spectrogram(wave, **kw):
    """wave [L] -> mel power [n_mels, T] (the layout MelSpectrogram returns)."""
    return (power_spectrogram(wave) @ melscale_fbanks(**kw)).T


def log_mel(wave):
    """wave [L] -> normalised log-mel [80, T] fp64 -- meldataset.py:644,650."""
    return (np.log(LOG_EPS + mel_spectrogram(wave)) - MEAN) / STD


def align_length(values, target_frames):
    """f0_backends.py:788-806."""
    values = np.asarray(values, dtype=np.float64)
    if target_frames <= 0:
        return np.zeros((0,), dtype=np.float32)
    if values.size == target_frames:
        return values.astype(np.float32)
    if values.size == 0:
        return np.zeros((target_frames,), dtype=np.float32)
    src = np.linspace(0.0, values.size - 1, num=values.size)
    dst = np.linspace(0.0, values.size - 1, num=target_frames)
    out = np.interp(dst, src, values)
    zero_mask = values == 0.0
    if np.any(zero_mask):
        nearest = np.clip(np.round(dst).astype(int), 0, values.size - 1)
        out[zero_mask[nearest]] = 0.0
    return out.astype(np.float32)


def build_training_example(wave, f0, crop_start=None, zero_value=0.0):
    """meldataset.py:629-677 without caches; ``crop_start`` replaces np.random.randint(0, T-192)."""
    wave = np.asarray(wave)
    if wave.ndim > 1:
        wave = wave.mean(axis=-1)
    wave = wave.astype(np.float32)
    mel = log_mel(wave)
    T = mel.shape[1]
    f0 = np.zeros((T,), np.float32) if f0 is None else align_length(f0, T)
    sil = (f0 == 0).astype(np.float32)
    if T > MAX_MEL_LENGTH:
        s = 0 if crop_start is None else int(crop_start)
        mel, f0, sil = mel[:, s:s + MAX_MEL_LENGTH], f0[s:s + MAX_MEL_LENGTH], sil[s:s + MAX_MEL_LENGTH]
    f0 = np.where(np.isnan(f0), np.float32(zero_value), f0)
    return mel, f0, sil


def collate(batch):
    """meldataset.py:804-826: zero-pad to 192 frames, stack, unsqueeze(1)."""
    B = len(batch)
    n_mels = batch[0][0].shape[0]
    mels = np.zeros((B, n_mels, MAX_MEL_LENGTH), np.float32)
    f0s = np.zeros((B, MAX_MEL_LENGTH), np.float32)
    sils = np.zeros((B, MAX_MEL_LENGTH), np.float32)
    for i, (mel, f0, sil) in enumerate(batch):
        n = mel.shape[1]
        mels[i, :, :n], f0s[i, :n], sils[i, :n] = mel, f0, sil
    return mels[:, None], f0s, sils
