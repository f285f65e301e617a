"""Data side of the hot path with the reference's API (reference meldataset.py:42-137,629-677,790-875): ``MelDataset``
items, ``Collater`` and ``build_dataloader``.

B200-first change of data flow: the reference computes one torchaudio mel per item on CPU inside DataLoader workers
(meldataset.py:644).  Here the workers only slice waveforms and labels; the batched log-mel runs on the GPU
(``mel.LogMel``) -- either per batch inside ``Trainer.run`` when the collater returns waveforms
(``Collater(return_wave=True)``, the flag the reference declares but never uses, meldataset.py:796-798), or per item
through ``MelDataset._build_training_example`` for API compatibility.  Audio file decoding, resampling and the F0
extractor cascade are outside this path: items come from in-memory / synthetic sources
or from ``.npy`` waveform + cached-F0 pairs.
"""
import os

import numpy as np
import torch
from torch.utils.data import DataLoader

from .mel import DEFAULT_MEL_PARAMS, LogMel
from . import synthetic

MAX_MEL_LENGTH = 192


def align_length(values, target_frames):
    """Linear-interpolation resize of an F0 track that keeps unvoiced (zero) frames zero (f0_backends.py:788-806)."""
    values = np.asarray(values, dtype=np.float64)
    if target_frames <= 0:
        return np.zeros((0,), dtype=np.float32)
    if values.size == target_frames:
        return values.astype(np.float32)
    if values.size == 0:
        return np.zeros((target_frames,), dtype=np.float32)
    grid = np.linspace(0.0, values.size - 1, num=target_frames)
    out = np.interp(grid, np.arange(values.size, dtype=np.float64), values)
    unvoiced = values == 0.0
    if unvoiced.any():
        out[unvoiced[np.clip(np.round(grid).astype(int), 0, values.size - 1)]] = 0.0
    return out.astype(np.float32)


class MelDataset(torch.utils.data.Dataset):
    """Same constructor as the reference (meldataset.py:43-52).  ``data_list`` lines are ``path|...``; a path may be
    a ``.npy`` waveform (with a cached ``<path>_f0.npy`` label track next to it) or ``synthetic:<seed>``."""

    def __init__(self, data_list, sr=DEFAULT_MEL_PARAMS["sample_rate"], mel_params=None, f0_params=None,
                 data_augmentation=False, validation=False, verbose=True, synthetic_data=None, device="cuda",
                 return_wave=False):
        self.verbose = verbose
        self.data_list = [l.rstrip("\n").split("|")[0] for l in data_list]
        mel_params = dict(mel_params or {})
        if "win_len" in mel_params and "win_length" not in mel_params:
            mel_params["win_length"] = mel_params.pop("win_len")
        self.mel_params = dict(DEFAULT_MEL_PARAMS)
        self.mel_params.update(mel_params)
        self.sr = sr if sr is not None else self.mel_params["sample_rate"]
        self.mel_params["sample_rate"] = self.sr
        self.f0_params = f0_params or {}
        # the reference's F0Extractor.cache_identifier ("-" + "_".join(backend keys), f0_backends.py:756-757) names the
        # cache files; label generation itself is out of scope, so the identifier is configuration here
        self.f0_cache_identifier = str(self.f0_params.get("cache_identifier", ""))
        self.zero_value = float(self.f0_params.get("zero_fill_value", 0.0))
        self.mean, self.std = -4, 4
        self.max_mel_length = MAX_MEL_LENGTH
        self.validation = validation
        self.data_augmentation = data_augmentation and (not validation)
        self.device = device
        self.return_wave = return_wave
        self.requires_cuda_backend = False
        self.synthetic_config = synthetic_data or {}
        self._base_length = len(self.data_list)
        self._synthetic_count = 0
        if self.synthetic_config.get("enabled", False) and not (validation and not self.synthetic_config.get(
                "apply_to_validation", False)):
            self._synthetic_count = max(1, int(round(self._base_length * float(self.synthetic_config.get("ratio", 0.25)))))
        # The reference draws crops / segment starts / the augmentation gain from the process-global numpy / random
        # generators (seeded with 1 at import, meldataset.py:31-32), which torch re-seeds per DataLoader worker and per
        # epoch.  A private generator pickled into every worker would repeat the same draws in all of them, so it is
        # created lazily inside the process that serves the items (see _random()).
        self._rng = None
        self._rng_owner = None
        self._logmel = None
        hop = self.mel_params["hop_length"]
        self.segment_samples = int(np.ceil((self.max_mel_length * hop + self.mel_params["n_fft"])))

    def __len__(self):
        return self._base_length + self._synthetic_count

    def _random(self):
        """Per-process generator: in a DataLoader worker it is seeded from the worker's torch seed (base_seed +
        worker_id, fresh every epoch for non-persistent workers), in the main process with 1 as the reference does."""
        pid = os.getpid()
        if self._rng is None or self._rng_owner != pid:
            info = torch.utils.data.get_worker_info()
            seed = 1 if info is None else int(info.seed) % (2 ** 32)
            self._rng, self._rng_owner = np.random.RandomState(seed), pid
        return self._rng

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_logmel"] = None  # CUDA tables are rebuilt lazily in the worker / after unpickling
        state["_rng"] = state["_rng_owner"] = None
        return state

    # ---------------------------------------------------------------- item sources
    def _load_item(self, idx):
        if idx >= self._base_length:
            return synthetic.make_segment(np.random.default_rng([1234, idx]), self.segment_samples, self.sr,
                                          self.mel_params["hop_length"])
        path = self.data_list[idx]
        if path.startswith("synthetic:"):
            return synthetic.make_segment(np.random.default_rng([int(path.split(":")[1]), idx]), self.segment_samples,
                                          self.sr, self.mel_params["hop_length"])
        return self.path_to_wave_and_label(path)

    # ---------------------------------------------------------------- real audio (meldataset.py:178-245)
    def _audio_metadata(self, path):
        """-> {'sample_rate', 'frames', 'channels'} of a PCM ``.wav`` (stdlib ``wave``) or a ``.npy`` waveform (shape
        [n] or [n, channels]; its rate is read from ``<path>.sr`` if present, else the dataset rate)."""
        if path.endswith(".npy"):
            arr = np.load(path, mmap_mode="r")
            sr_file = path + ".sr"
            sr = int(open(sr_file).read().strip()) if os.path.isfile(sr_file) else self.sr
            return {"sample_rate": sr, "frames": int(arr.shape[0]), "channels": int(arr.shape[1]) if arr.ndim > 1 else 1}
        if path.lower().endswith(".wav"):
            import wave as _wave
            with _wave.open(path, "rb") as w:
                return {"sample_rate": w.getframerate(), "frames": w.getnframes(), "channels": w.getnchannels()}
        raise IndexError("unsupported audio container %r: this path reads PCM .wav (stdlib) and .npy waveforms; other "
                         "formats need the reference's soundfile/librosa loaders, which are out of scope" % path)

    def _read_audio(self, path, start_frame=0, num_frames=None):
        """-> (float32 [n] or [n, channels] in [-1, 1), sample_rate); ``num_frames`` None reads to the end."""
        if path.endswith(".npy"):
            arr = np.load(path, mmap_mode="r")
            end = arr.shape[0] if num_frames is None else min(arr.shape[0], start_frame + num_frames)
            return np.asarray(arr[start_frame:end], dtype=np.float32), self._audio_metadata(path)["sample_rate"]
        import wave as _wave
        with _wave.open(path, "rb") as w:
            sr, ch, width, total = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
            w.setpos(min(start_frame, total))
            raw = w.readframes(total - start_frame if num_frames is None else num_frames)
        if width == 2:
            data = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif width == 4:
            data = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
        elif width == 1:
            data = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif width == 3:
            b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            data = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
        else:
            raise IndexError("unsupported PCM sample width %d in %r" % (width, path))
        return (data.reshape(-1, ch) if ch > 1 else data), sr

    def path_to_wave_and_label(self, path):
        """Random training segment of a file + its cached F0 labels, as ``path_to_mel_and_label`` does up to the mel
        (meldataset.py:178-230): the segment is requested in SOURCE samples, mixed down to mono, resampled (on the GPU,
        ``resample.py``) when the file's rate differs from the dataset's, and the whole-file F0 cache is sliced at the
        segment's position.  -> (wave float32 [n] at self.sr, f0 or None)"""
        from . import cache
        meta = self._audio_metadata(path)
        source_sr, total = meta["sample_rate"], meta["frames"]
        hop = int(self.mel_params["hop_length"])
        window = int(self.mel_params.get("win_length") or self.mel_params.get("n_fft", hop))
        requested = (self.max_mel_length * hop) / float(self.sr) + max(window, hop) / float(self.sr)
        segment = int(np.ceil(requested * float(source_sr)))
        start, use_full = 0, True
        if 0 < segment < total:
            start = int(self._random().randint(0, total - segment + 1))   # random.randint(0, max_start), inclusive
            use_full = False
        wave, wave_sr = self._read_audio(path, start, None if use_full else segment)
        if wave.ndim > 1:
            wave = wave.mean(axis=-1)
        wave = wave.astype(np.float32)
        if wave_sr != self.sr:
            wave = self._resample(wave, wave_sr)
        start_resampled = 0 if use_full else int(round(start / float(source_sr) * self.sr))
        expected = None if use_full else int(np.ceil(len(wave) / max(hop, 1))) + 2
        f0 = cache.load_cached_f0(path, self.f0_cache_identifier, self.sr, hop)
        if f0 is None:
            if not self.f0_params.get("allow_missing_f0", False):
                raise RuntimeError("no cached F0 for %r (expected %s): F0 extraction (pyworld / CREPE / ...) is outside "
                                   "this path -- prepare the caches with the reference, or pass f0_params="
                                   "{'allow_missing_f0': True} to train the voicing of unlabeled audio as silence"
                                   % (path, cache.f0_cache_paths(path, self.f0_cache_identifier)[0]))
        else:
            f0 = cache.slice_cached_f0(f0, start_resampled, expected, hop)
        return wave, f0

    def _resample(self, wave, source_sr):
        from .resample import resample
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError("resampling runs on the GPU: construct the dataset with device='cuda'")
        return resample(torch.from_numpy(np.ascontiguousarray(wave)).to(dev), int(source_sr), int(self.sr)).cpu().numpy()

    def __getitem__(self, idx):
        wave, f0 = self._load_item(idx)
        if self.data_augmentation:  # random gain in [0.5, 1) (meldataset.py:232-234)
            wave = (0.5 + 0.5 * self._random().random_sample()) * np.asarray(wave, dtype=np.float32)
        if self.return_wave:
            return self._build_wave_example(wave, f0)
        return self._build_training_example(wave, self.sr, f0)

    # ---------------------------------------------------------------- label / crop logic (meldataset.py:652-677)
    def _labels(self, f0, mel_length):
        f0 = np.zeros((mel_length,), np.float32) if f0 is None else align_length(f0, mel_length)
        start = 0
        if mel_length > self.max_mel_length:
            start = int(self._random().randint(0, mel_length - self.max_mel_length))
            f0 = f0[start:start + self.max_mel_length]
        sil = (f0 == 0).astype(np.float32)
        f0 = np.where(np.isnan(f0), np.float32(self.zero_value), f0).astype(np.float32)
        return torch.from_numpy(f0), torch.from_numpy(sil), start

    def _build_wave_example(self, waveform, f0):
        """Waveform-level item for the fused path: (wave [L], f0 [<=192], is_silence [<=192], crop_start)."""
        waveform = np.asarray(waveform)
        if waveform.ndim > 1:
            waveform = waveform.mean(axis=-1)
        wave = torch.from_numpy(waveform.astype(np.float32))
        f0_t, sil, start = self._labels(f0, 1 + wave.shape[0] // self.mel_params["hop_length"])
        return wave, f0_t, sil, start

    def _build_training_example(self, waveform, sr, f0, cache_key=None, allow_cache=True):
        """Reference contract (meldataset.py:629-677): -> (mel [80, <=192], f0 [<=192], is_silence [<=192]); the mel is
        computed by the CUDA log-mel kernels."""
        waveform = np.asarray(waveform)
        if waveform.ndim > 1:
            waveform = waveform.mean(axis=-1)
        if sr != self.sr:  # meldataset.py:621-627, on the GPU
            waveform = self._resample(waveform.astype(np.float32), sr)
        wave, f0_t, sil, start = self._build_wave_example(waveform, f0)
        if self._logmel is None:
            self._logmel = LogMel(self.device, **self.mel_params)
        T = self._logmel.num_frames(wave.shape[0])
        T_out = min(T, self.max_mel_length)
        if cache_key is not None and allow_cache and not self.data_augmentation:
            # whole-file mel POWER cache in the reference's format (meldataset.py:638-648,679-786)
            from . import cache
            meta = cache.mel_metadata(wave.shape[0], 1, self.sr, self.sr, self.mel_params)
            power = cache.load_cached_mel(cache_key, meta)
            if power is None:
                full = self._logmel(wave[None].to(self.device))[0]
                cache.save_mel_cache(cache_key, torch.clamp(torch.exp(4.0 * full - 4.0) - 1e-5, min=0.0).cpu().numpy(), meta)
                mel = full[:, start:start + T_out]
            else:
                mel = ((torch.log(1e-5 + torch.from_numpy(power).to(self.device)) + 4.0) / 4.0)[:, start:start + T_out]
            return mel.contiguous(), f0_t, sil
        mel = self._logmel(wave[None].to(self.device), crop=torch.tensor([start], dtype=torch.int32), T_out=T_out)[0]
        return mel, f0_t, sil


class Collater(object):
    """Zero-pad to 192 frames and stack (meldataset.py:790-826).  With ``return_wave=True`` the batch carries raw
    waveforms instead of mels: ``(waves [B, L], f0s [B, 192], is_silences [B, 192], crops [B], lengths [B])`` -- crop
    offsets in frames and every item's own sample count, so that the batched GPU log-mel reflect-pads each item at its
    own end and zero-fills the frames past it, exactly like the per-item mel + zero padding of the reference."""

    def __init__(self, return_wave=False):
        self.return_wave = return_wave
        self.max_mel_length = MAX_MEL_LENGTH

    def __call__(self, batch):
        B = len(batch)
        f0s = torch.zeros((B, self.max_mel_length)).float()
        sils = torch.zeros((B, self.max_mel_length)).float()
        if self.return_wave:
            L = max(item[0].shape[0] for item in batch)
            waves = torch.zeros((B, L)).float()
            crops = torch.zeros((B,), dtype=torch.int32)
            lengths = torch.zeros((B,), dtype=torch.int32)
            for i, (wave, f0, sil, start) in enumerate(batch):
                waves[i, :wave.shape[0]] = wave
                f0s[i, :f0.shape[0]] = f0
                sils[i, :sil.shape[0]] = sil
                crops[i] = start
                lengths[i] = wave.shape[0]
            return waves, f0s, sils, crops, lengths
        nmels = batch[0][0].size(0)
        mels = torch.zeros((B, nmels, self.max_mel_length), device=batch[0][0].device).float()
        for i, (mel, f0, sil) in enumerate(batch):
            n = mel.size(1)
            mels[i, :, :n] = mel
            f0s[i, :n] = f0
            sils[i, :n] = sil
        return mels.unsqueeze(1), f0s, sils


def build_dataloader(path_list, validation=False, batch_size=4, num_workers=1, device="cpu", collate_config=None,
                     dataset_config=None):
    """Reference signature (meldataset.py:829-875)."""
    dataset_config = dict(dataset_config or {})
    opts = dataset_config.pop("dataloader", {}) or {}
    collate_config = dict(collate_config or {})
    dataset_config.setdefault("return_wave", bool(collate_config.get("return_wave", False)))
    dataset = MelDataset(path_list, validation=validation, **dataset_config)
    kwargs = dict(batch_size=batch_size, shuffle=(not validation), num_workers=num_workers,
                  drop_last=(not validation), collate_fn=Collater(**collate_config), pin_memory=(device != "cpu"))
    if not dataset.return_wave:
        kwargs["num_workers"] = 0  # per-item CUDA mels cannot run in forked workers; the waveform path can
        kwargs["pin_memory"] = False
    if opts.get("start_method") and kwargs["num_workers"] > 0:
        kwargs["multiprocessing_context"] = torch.multiprocessing.get_context(opts["start_method"])
    if opts.get("persistent_workers") is not None and kwargs["num_workers"] > 0:
        kwargs["persistent_workers"] = bool(opts["persistent_workers"])
    if opts.get("prefetch_factor") is not None and kwargs["num_workers"] > 0:
        kwargs["prefetch_factor"] = int(opts["prefetch_factor"])
    return DataLoader(dataset, **kwargs)
