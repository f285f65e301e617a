"""Seeded inputs shared by the golden-fixture generator and the tests (no reference import needed)."""
import numpy as np
import torch

SEG = 58624
CROP_SEED = 7


def logmel_signals():
    rng = np.random.default_rng(0)
    t = np.arange(SEG) / 24000.0
    return {
        "noise": (0.1 * rng.standard_normal(SEG)).astype(np.float32),
        "tone220": (0.5 * np.sin(2 * np.pi * 220.0 * t)).astype(np.float32),
        "harm3": (0.3 * np.sin(2 * np.pi * 180 * t) + 0.2 * np.sin(2 * np.pi * 360 * t)
                  + 0.1 * np.sin(2 * np.pi * 540 * t) + 1e-3 * rng.standard_normal(SEG)).astype(np.float32),
        "silence": np.zeros(SEG, np.float32),
        "short": (0.2 * rng.standard_normal(24000)).astype(np.float32),  # 81 frames: no crop, collater pads
    }


def f0_track(name):
    """Label track handed to _build_training_example (one value per 12.5 ms hop, a few frames of slack)."""
    n = {"short": 83}.get(name, 200)
    rng = np.random.default_rng(abs(hash(name)) % 1000 if False else len(name))
    f0 = 100.0 + 150.0 * rng.random(n)
    f0[rng.random(n) < 0.3] = 0.0
    return f0.astype(np.float64)


def crop_start(T):
    """What np.random.randint(0, T - 192) yields right after np.random.seed(CROP_SEED)."""
    np.random.seed(CROP_SEED)
    return int(np.random.randint(0, T - 192)) if T > 192 else 0


def align_cases():
    rng = np.random.default_rng(5)
    a = 200.0 * rng.random(200)
    a[rng.random(200) < 0.4] = 0.0
    return [(a, 196), (a[:150], 196), (a, 81), (np.zeros(0), 10), (a[:196], 196), (a, 0)]


def model_config(model_type):
    return dict(model_type=model_type, num_layers=4, dropout=0.1, nhead=8, dim_feedforward=1536, max_len=2048)


def model_state_dict(model_type):
    """Deterministic weights in the reference state_dict layout, built by this repo's parameter container on CPU."""
    from pitchextractor_b200.model import JDCNet
    torch.manual_seed(1234)
    m = JDCNet(num_class=1, sequence_model_config=model_config(model_type))
    g = torch.Generator().manual_seed(99)
    for n, p in m.named_parameters():
        if p.dim() == 1 and "model.bias_" not in n:
            p.data.add_(0.1 * torch.randn(p.shape, generator=g))
    for n, b in m.named_buffers():
        if n.endswith("running_mean"):
            b.copy_(0.05 * torch.randn(b.shape, generator=g))
        if n.endswith("running_var"):
            b.copy_(0.8 + 0.4 * torch.rand(b.shape, generator=g))
    return {k: v.detach().clone().contiguous() for k, v in m.state_dict().items()}


def model_inputs(B=2):
    g = torch.Generator().manual_seed(42)
    mel = torch.randn(B, 1, 80, 192, generator=g) * 0.5
    sil = (torch.rand(B, 192, generator=g) < 0.25).float()
    f0 = (torch.rand(B, 192, generator=g) * 200.0 + 100.0) * (1 - sil)
    return mel, f0, sil
