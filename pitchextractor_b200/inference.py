"""Chunked F0 inference and the evaluation metrics of the reference's notebooks.

``predict_f0`` follows ``predict_f0`` / ``waveform_to_mel`` of reference Utils/dynamic_pitch_behavior.ipynb (cell 5):
192-frame chunks every 144 frames (overlap 48), zero-padded tail, the classifier output of every chunk truncated to
the chunk's true length and concatenated (overlapping frames therefore appear twice, exactly as in the reference).
B200-first difference: the log-mel of the whole clip is one kernel launch and ALL chunks go through JDCNet as one
batch (eval mode: BatchNorm running statistics, no dropout) instead of one forward per chunk.

The metrics restate ``compute_metrics`` (same notebook cell) and Utils/dynamic_pitch_tools.py:79-104.
"""
import numpy as np
import torch

from .mel import LogMel

CHUNK_SIZE, CHUNK_OVERLAP, VOICING_THRESHOLD_HZ = 192, 48, 10.0


@torch.no_grad()
def predict_f0(model, audio, chunk_size=CHUNK_SIZE, overlap=CHUNK_OVERLAP, logmel=None, return_voicing=False):
    """audio: 1-D float array / tensor at the model's sample rate -> np.float32 F0 track in Hz (reference stitching)."""
    dev = next(model.parameters()).device
    logmel = logmel or LogMel(dev)
    wave = torch.as_tensor(np.asarray(audio, dtype=np.float32) if not torch.is_tensor(audio) else audio)
    mel = logmel(wave.reshape(1, -1).to(dev))[0]  # [80, T]
    total = mel.shape[-1]
    step = max(chunk_size - overlap, 1)
    starts = list(range(0, total, step))
    if not starts:
        return np.zeros((0,), dtype=np.float32)
    batch = torch.zeros(len(starts), 1, mel.shape[0], chunk_size, device=dev)
    for i, s in enumerate(starts):
        e = min(s + chunk_size, total)
        batch[i, 0, :, :e - s] = mel[:, s:e]
    was_training = model.training
    model.eval()
    cls, det = model(batch.transpose(-1, -2))
    model.train(was_training)
    cls = cls.squeeze(-1).float().cpu().numpy()
    det = det.float().cpu().numpy()
    f0 = np.concatenate([cls[i, :min(s + chunk_size, total) - s] for i, s in enumerate(starts)]).astype(np.float32)
    if return_voicing:
        sil = np.concatenate([det[i, :min(s + chunk_size, total) - s] for i, s in enumerate(starts)])
        return f0, sil.astype(np.float32)
    return f0


def hz_to_cents(f0):
    f0 = np.asarray(f0)
    cents = np.zeros_like(f0, dtype=np.float32)
    pos = f0 > 0
    cents[pos] = 1200.0 * np.log2(f0[pos] / 55.0)
    return cents


def circular_cents_distance(a, b):
    return np.mod(a - b + 600.0, 1200.0) - 600.0


def rms_cents_error(reference, prediction):
    n = min(reference.shape[0], prediction.shape[0])
    if n == 0:
        return float("nan")
    ref, pred = reference[:n], prediction[:n]
    voiced = ref > 0
    if not np.any(voiced):
        return float("nan")
    diff = hz_to_cents(np.clip(pred[voiced], a_min=1e-5, a_max=None)) - hz_to_cents(ref[voiced])
    return float(np.sqrt(np.mean(diff ** 2)))


def compute_metrics(reference, prediction, voicing_threshold_hz=VOICING_THRESHOLD_HZ):
    """Raw pitch accuracy, raw chroma accuracy (both within 50 cents), voicing decision accuracy (prediction above
    ``voicing_threshold_hz`` == voiced) and octave-error rate, over the frames both tracks have."""
    n = min(reference.shape[0], prediction.shape[0])
    ref, pred = reference[:n], prediction[:n]
    ref_v = ref > 0
    vuv = float(np.count_nonzero(ref_v == (pred > voicing_threshold_hz)) / max(n, 1))
    nv = int(np.count_nonzero(ref_v))
    if nv == 0:
        return {"RPA": float("nan"), "RCA": float("nan"), "VUV": vuv, "OctaveError": float("nan")}
    rc = hz_to_cents(ref[ref_v])
    pc = hz_to_cents(np.clip(pred[ref_v], a_min=1e-5, a_max=None))
    d = pc - rc
    octv = np.round(d / 1200.0)
    octave_err = (np.abs(d) > 50.0) & (octv != 0) & (np.abs(d - octv * 1200.0) <= 50.0)
    return {"RPA": float(np.count_nonzero(np.abs(d) <= 50.0) / nv),
            "RCA": float(np.count_nonzero(np.abs(circular_cents_distance(pc, rc)) <= 50.0) / nv),
            "VUV": vuv, "OctaveError": float(np.count_nonzero(octave_err) / nv)}


def voicing_accuracy(sil_logit, reference_f0):
    """Detector-head accuracy: sigmoid(logit) > 0.5 means silence / unvoiced (label ``f0 == 0``, meldataset.py:659-665)."""
    n = min(sil_logit.shape[0], reference_f0.shape[0])
    return float(np.mean((sil_logit[:n] > 0.0) == (reference_f0[:n] == 0)))


def estimate_tracking_delay_ms(reference, prediction, frame_period_ms):
    """Lag (ms) at which the mean-removed prediction correlates best with the mean-removed reference track
    (Utils/dynamic_pitch_tools.py:107-124); positive = the prediction trails the reference."""
    n = min(reference.shape[0], prediction.shape[0])
    if n == 0:
        return float("nan")
    ref = reference[:n] - np.mean(reference[:n])
    pred = prediction[:n] - np.mean(prediction[:n])
    if np.allclose(ref, 0) or np.allclose(pred, 0):
        return float("nan")
    xc = np.correlate(pred, ref, mode="full")
    return float((int(np.argmax(xc)) - (n - 1)) * frame_period_ms)


def compute_overshoot_cents(reference, prediction):
    """Peak of the predicted track relative to the reference's final value, in cents
    (Utils/dynamic_pitch_tools.py:127-136)."""
    n = min(reference.shape[0], prediction.shape[0])
    if n == 0:
        return float("nan")
    target, peak = reference[:n][-1], np.max(prediction[:n])
    if target <= 0 or peak <= 0:
        return float("nan")
    return float(1200.0 * np.log2(peak / target))
