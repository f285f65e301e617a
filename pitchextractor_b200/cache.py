"""On-disk label / spectrogram caches in the reference's formats, so that datasets prepared by the
reference drop in unchanged and vice versa:

  <audio path>_f0<identifier>.npy    float32 F0 track in Hz, one value per hop (0 = unvoiced)      meldataset.py:519-522
  <audio path>_f0<identifier>.json   {"cache_identifier", "backend", "sample_rate", "hop_length"}  meldataset.py:606-619
  <audio path>_f0.npy                legacy cache without metadata (accepted as is)                meldataset.py:595-602
  <audio path>_mel.npy               float32 mel POWER spectrogram [n_mels, T] (before log / normalisation)
  <audio path>_mel_meta.json         {"audio_sample_rate", "audio_num_samples", "audio_num_channels",
                                      "dataset_sample_rate", "mel_params"}                          meldataset.py:679-704

``identifier`` is the reference's ``F0Extractor.cache_identifier``: "-" + "_".join(backend cache keys), "" when no
backend is configured (f0_backends.py:756-757).  Validation follows the reference (meldataset.py:566-604,706-742): a
cache whose metadata does not match is ignored.  One deliberate difference: the reference DELETES stale cache files (and,
on a mel-metadata mismatch, every cache of the dataset); this module never removes anything -- a mismatching cache simply
reads as "absent".
"""
import json
import os

import numpy as np

MEL_CACHE_SUFFIX, MEL_META_SUFFIX = "_mel.npy", "_mel_meta.json"


def f0_cache_paths(path, identifier=""):
    data = path + "_f0%s.npy" % identifier
    return data, data[:-4] + ".json", path + "_f0.npy"


def load_cached_f0(path, identifier, sample_rate, hop_length):
    """-> float32 F0 track, or None if there is no valid cache (meldataset.py:566-604)."""
    data_path, meta_path, legacy_path = f0_cache_paths(path, identifier)
    if os.path.isfile(data_path) and os.path.isfile(meta_path):
        try:
            with open(meta_path, "r", encoding="utf-8") as f:
                meta = json.load(f)
        except (OSError, json.JSONDecodeError):
            meta = None
        expected = {"cache_identifier": identifier, "sample_rate": int(sample_rate), "hop_length": int(hop_length)}
        if meta and all(meta.get(k) == v for k, v in expected.items()):
            try:
                return np.load(data_path).astype(np.float32)
            except (OSError, ValueError):
                pass
    if os.path.isfile(legacy_path):
        try:
            return np.load(legacy_path).astype(np.float32)
        except (OSError, ValueError):
            pass
    return None


def save_f0_cache(path, f0, backend_name, identifier, sample_rate, hop_length):
    """meldataset.py:606-619."""
    data_path, meta_path, _ = f0_cache_paths(path, identifier)
    np.save(data_path, np.asarray(f0, dtype=np.float32))
    meta = {"cache_identifier": identifier, "backend": backend_name, "sample_rate": int(sample_rate),
            "hop_length": int(hop_length)}
    with open(meta_path, "w", encoding="utf-8") as f:
        json.dump(meta, f, sort_keys=True)


def slice_cached_f0(f0, start_sample, expected_frames, hop_length):
    """The part of a whole-file F0 track that belongs to a segment starting at ``start_sample`` (meldataset.py:531-538)."""
    if expected_frames is None:
        return f0
    hop = max(int(hop_length), 1)
    start = max(0, int(np.floor(start_sample / float(hop))))
    if start >= f0.shape[0]:
        return np.zeros((0,), dtype=np.float32)
    return f0[start:min(f0.shape[0], start + int(expected_frames) + 4)]


def mel_metadata(num_samples, num_channels, audio_sample_rate, dataset_sample_rate, mel_params):
    """meldataset.py:679-701."""
    def ser(v):
        if isinstance(v, np.ndarray):
            return v.tolist()
        if isinstance(v, np.generic):
            return v.item()
        return v
    return {"audio_sample_rate": int(audio_sample_rate), "audio_num_samples": int(num_samples),
            "audio_num_channels": int(num_channels), "dataset_sample_rate": int(dataset_sample_rate),
            "mel_params": {k: ser(v) for k, v in mel_params.items()}}


def mel_cache_paths(path):
    return path + MEL_CACHE_SUFFIX, path + MEL_META_SUFFIX


def load_cached_mel(path, expected_metadata):
    """-> float32 mel power [n_mels, T] or None (meldataset.py:706-742)."""
    mel_path, meta_path = mel_cache_paths(path)
    if not (os.path.isfile(mel_path) and os.path.isfile(meta_path)):
        return None
    try:
        with open(meta_path, "r", encoding="utf-8") as f:
            if json.load(f) != expected_metadata:
                return None
        return np.load(mel_path)
    except (OSError, ValueError, json.JSONDecodeError):
        return None


def save_mel_cache(path, mel_power, metadata):
    """meldataset.py:778-786."""
    mel_path, meta_path = mel_cache_paths(path)
    np.save(mel_path, np.asarray(mel_power, dtype=np.float32))
    with open(meta_path, "w", encoding="utf-8") as f:
        json.dump(metadata, f, sort_keys=True)
