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
        if path.endswith(".npy"):
            wave = np.load(path).astype(np.float32)
            f0_path = path[:-4] + "_f0.npy"
            f0 = np.load(f0_path) if os.path.isfile(f0_path) else None
            if wave.shape[0] > self.segment_samples:  # random segment, as meldataset.py:196-201
                start = int(self._random().randint(0, wave.shape[0] - self.segment_samples))
                hop = self.mel_params["hop_length"]
                if f0 is not None:
                    f0 = f0[start // hop: start // hop + 1 + self.segment_samples // hop + 4]
                wave = wave[start:start + self.segment_samples]
            return wave, f0
        raise IndexError("audio decoding is outside the accelerated path; provide .npy waveforms or synthetic: items "
                         "(got %r)" % path)

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
        if sr != self.sr:
            raise ValueError("resampling is outside the accelerated path (sr %d != %d)" % (sr, self.sr))
        wave, f0_t, sil, start = self._build_wave_example(waveform, f0)
        if self._logmel is None:
            self._logmel = LogMel(self.device, **self.mel_params)
        T = self._logmel.num_frames(wave.shape[0])
        T_out = min(T, self.max_mel_length)
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
