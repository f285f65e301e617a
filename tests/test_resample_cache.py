"""the on-disk F0 / mel cache formats and the GPU resampler + real-audio ingestion."""
import json
import os
import sys
import wave

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RATES = [(44100, 24000), (48000, 24000), (16000, 24000), (22050, 24000), (24000, 16000), (32000, 24000)]


# ------------------------------------------------------------------------------------------------ CPU: oracle pinning
@pytest.mark.parametrize("orig,new", RATES)
def test_resample_oracle_and_filter_bank_vs_torchaudio(orig, new):
    """The numpy oracle and the product's host-side filter bank against torchaudio itself (the reference's call,
    meldataset.py:621-627): bank bit-identical, resampled signal within fp32 accumulation error."""
    import math
    import torchaudio
    from torchaudio.functional.functional import _get_sinc_resample_kernel
    from oracle import resample_np
    from pitchextractor_b200.resample import sinc_filter_bank
    k, w = _get_sinc_resample_kernel(orig, new, math.gcd(orig, new))
    h, up, down, width = sinc_filter_bank(orig, new)
    assert width == w and torch.equal(h, k[:, 0, :])
    rng = np.random.default_rng(orig + new)
    x = (0.3 * rng.standard_normal(5000)).astype(np.float32)
    ref = torchaudio.functional.resample(torch.from_numpy(x)[None], orig, new)[0].numpy()
    got = resample_np.resample(x, orig, new)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())  # fp32 conv1d of torchaudio vs fp64


def test_f0_and_mel_cache_formats(tmp_path):
    from pitchextractor_b200 import cache
    path = str(tmp_path / "clip.wav")
    f0 = np.linspace(0, 300, 77).astype(np.float32)
    assert cache.load_cached_f0(path, "-pyworld", 24000, 300) is None
    cache.save_f0_cache(path, f0, "pyworld", "-pyworld", 24000, 300)
    assert os.path.isfile(path + "_f0-pyworld.npy") and os.path.isfile(path + "_f0-pyworld.json")
    meta = json.load(open(path + "_f0-pyworld.json"))
    assert meta == {"backend": "pyworld", "cache_identifier": "-pyworld", "hop_length": 300, "sample_rate": 24000}
    np.testing.assert_array_equal(cache.load_cached_f0(path, "-pyworld", 24000, 300), f0)
    assert cache.load_cached_f0(path, "-pyworld", 22050, 300) is None      # metadata mismatch -> ignored
    assert cache.load_cached_f0(path, "-crepe", 24000, 300) is None        # other backend cascade -> other file
    np.save(path + "_f0.npy", f0[:10])                                      # legacy cache without metadata
    np.testing.assert_array_equal(cache.load_cached_f0(path, "-crepe", 24000, 300), f0[:10])
    np.testing.assert_array_equal(cache.slice_cached_f0(f0, 3000, 20, 300), f0[10:34])
    assert cache.slice_cached_f0(f0, 10 ** 9, 20, 300).shape == (0,)
    mp = {"sample_rate": 24000, "n_mels": 80, "n_fft": 1024, "win_length": 1024, "hop_length": 300}
    m = cache.mel_metadata(58624, 1, 24000, 24000, mp)
    mel = np.random.default_rng(0).random((80, 196)).astype(np.float32)
    assert cache.load_cached_mel(path, m) is None
    cache.save_mel_cache(path, mel, m)
    np.testing.assert_array_equal(cache.load_cached_mel(path, m), mel)
    assert cache.load_cached_mel(path, dict(m, audio_num_samples=1)) is None


def test_cache_files_interchange_with_the_live_reference(tmp_path):
    """Files written by the reference's own MelDataset are read by this package and vice versa."""
    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference tree not present")
    from pitchextractor_b200 import cache
    ns = refshim.load()
    mp = {"sample_rate": 24000, "n_mels": 80, "n_fft": 1024, "win_length": 1024, "hop_length": 300}
    ds = ns.meldataset.MelDataset([], mel_params=dict(mp), verbose=False)
    ident = ds.f0_extractor.cache_identifier
    a, b = str(tmp_path / "a.wav"), str(tmp_path / "b.wav")
    f0 = (100 + np.arange(50)).astype(np.float32)
    ds._save_f0_cache(a, f0, "pyworld")                                       # reference writes, we read
    np.testing.assert_array_equal(cache.load_cached_f0(a, ident, 24000, 300), f0)
    cache.save_f0_cache(b, f0 * 2, "pyworld", ident, 24000, 300)              # we write, the reference reads
    np.testing.assert_array_equal(ds._load_cached_f0(b), f0 * 2)
    wave_t = torch.zeros(58624)
    meta_ref = ds._build_mel_metadata(wave_t, 24000)
    assert meta_ref == cache.mel_metadata(58624, 1, 24000, 24000, ds.mel_params)
    mel = torch.rand(80, 196)
    ds._save_mel_cache(a, mel, meta_ref)
    np.testing.assert_array_equal(cache.load_cached_mel(a, meta_ref), mel.numpy())
    cache.save_mel_cache(b, mel.numpy() * 3, meta_ref)
    assert torch.equal(ds._load_cached_mel(b, meta_ref), mel * 3)


def _write_wav(path, data, sr):
    with wave.open(path, "wb") as w:
        w.setnchannels(data.shape[1] if data.ndim > 1 else 1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes((np.clip(data, -1, 1) * 32767.0).astype("<i2").tobytes())


def test_wav_metadata_and_segment_reads(tmp_path):
    from pitchextractor_b200 import MelDataset
    rng = np.random.default_rng(1)
    stereo = (0.2 * rng.standard_normal((100000, 2))).astype(np.float32)
    p = str(tmp_path / "s.wav")
    _write_wav(p, stereo, 44100)
    ds = MelDataset([p + "|0\n"], verbose=False, return_wave=True, device="cpu")
    assert ds._audio_metadata(p) == {"sample_rate": 44100, "frames": 100000, "channels": 2}
    seg, sr = ds._read_audio(p, 1000, 512)
    assert sr == 44100 and seg.shape == (512, 2)
    np.testing.assert_allclose(seg, np.round(stereo[1000:1512] * 32767.0) / 32768.0, atol=1.0 / 32768.0)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("orig,new", RATES)
def test_gpu_resample_matches_torchaudio(built_lib, orig, new):
    import torchaudio
    from oracle import resample_np
    from pitchextractor_b200.resample import resample
    rng = np.random.default_rng(orig)
    x = (0.3 * rng.standard_normal((3, 30011))).astype(np.float32)
    got = resample(torch.from_numpy(x).cuda(), orig, new).cpu().numpy()
    ref = torchaudio.functional.resample(torch.from_numpy(x), orig, new).numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), np.abs(got - ref).max()
    o = resample_np.resample(x[1], orig, new)
    assert np.abs(got[1] - o).max() <= 5e-6 * max(1.0, np.abs(o).max())
    # ragged batch: per-item lengths treat the padding as silence, exactly like resampling each item on its own
    lens = torch.tensor([30011, 12345, 7], dtype=torch.int32)
    xz = x.copy()
    for i, n in enumerate(lens.tolist()):
        xz[i, n:] = 0
    a = resample(torch.from_numpy(x).cuda(), orig, new, lengths=lens).cpu().numpy()
    b = resample(torch.from_numpy(xz).cuda(), orig, new).cpu().numpy()
    np.testing.assert_array_equal(a, b)


@pytest.mark.gpu
def test_dataset_ingests_a_44k_stereo_wav_with_cached_f0(built_lib, tmp_path):
    """meldataset.py:178-245 on a real file: random segment in source samples, mono mixdown, GPU resample to 24 kHz,
    F0 cache sliced at the segment, log-mel -- against the same steps done with torchaudio + the numpy oracle."""
    import torchaudio
    from oracle import logmel_np
    from pitchextractor_b200 import MelDataset, cache
    rng = np.random.default_rng(5)
    sr0, n = 44100, 44100 * 4
    t = np.arange(n) / sr0
    stereo = np.stack([0.3 * np.sin(2 * np.pi * 220 * t), 0.2 * np.sin(2 * np.pi * 330 * t)], 1).astype(np.float32)
    stereo += 1e-3 * rng.standard_normal(stereo.shape).astype(np.float32)
    p = str(tmp_path / "clip.wav")
    _write_wav(p, stereo, sr0)
    total_frames_24k = 1 + int(np.ceil(n * 24000 / sr0)) // 300
    f0_full = (150.0 + np.arange(total_frames_24k)).astype(np.float32)
    cache.save_f0_cache(p, f0_full, "pyworld", "-pyworld", 24000, 300)
    ds = MelDataset([p + "|0\n"], verbose=False, f0_params={"cache_identifier": "-pyworld"})
    mel, f0, sil = ds[0]
    assert mel.shape == (80, 192) and f0.shape == (192,) and float(sil.sum()) == 0.0
    # replay: the dataset drew (segment start, crop) from a generator seeded with 1 in this process
    r = np.random.RandomState(1)
    seg = int(np.ceil(((192 * 300) / 24000.0 + 1024 / 24000.0) * sr0))
    start = int(r.randint(0, n - seg + 1))
    pcm = (np.clip(stereo, -1, 1) * 32767.0).astype("<i2").astype(np.float32) / 32768.0   # as _write_wav quantises
    mono = pcm[start:start + seg].mean(axis=-1).astype(np.float32)
    res = torchaudio.functional.resample(torch.from_numpy(mono)[None], sr0, 24000)[0].numpy()
    ref = logmel_np.log_mel(res)
    crop = int(r.randint(0, ref.shape[1] - 192))
    assert np.abs(mel.cpu().numpy() - ref[:, crop:crop + 192]).max() <= 2e-4 * max(1.0, np.abs(ref).max())
    first = int(np.floor(round(start / sr0 * 24000) / 300.0))
    assert 140.0 <= f0.min().item() and abs(f0[0].item() - (150.0 + first + crop)) <= 3.0
