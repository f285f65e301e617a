"""CPU port of one reference optimisation step (oracle; test infrastructure and the bench's CPU baseline only).

Follows reference ``Trainer.run`` (trainer.py:219-252) on CPU, where the reference itself disables AMP and gradient
checkpointing (trainer.py:64,103): per-sample log-mel as ``MelDataset._build_training_example`` does it
(meldataset.py:644,650; torchaudio == torch.stft + |.|^2 + filterbank matmul, torchaudio/functional/functional.py:
122-144, torchaudio/transforms/_transforms.py:417), ``Collater`` (meldataset.py:804-826), fp32 forward + losses +
backward through ``oracle.jdcnet_torch`` and AdamW / OneCycleLR as built by reference ``optimizers.py:54-76``.
"""
import math

import numpy as np
import torch

from . import jdcnet_torch as J
from . import logmel_np


def fbanks_fp32():
    """fp32 filterbank computed with torchaudio's own op order (functional.py:544-572)."""
    all_freqs = torch.linspace(0, 24000 // 2, 513)
    m_pts = torch.linspace(2595.0 * math.log10(1.0 + 0.0 / 700.0), 2595.0 * math.log10(1.0 + 12000.0 / 700.0), 82)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    return torch.max(torch.zeros(1), torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))


_FB = {}


def log_mel_torch(wave):
    """wave [L] fp32 tensor -> [80, T] fp32, the reference's own arithmetic (torch.stft path), on wave's device."""
    dev = wave.device
    if dev not in _FB:
        _FB[dev] = (fbanks_fp32().to(dev), torch.hann_window(1024).to(dev))
    fb, win = _FB[dev]
    spec = torch.stft(wave, 1024, 300, 1024, window=win, center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    mel = (spec.abs().pow(2.0).transpose(-1, -2) @ fb).transpose(-1, -2)
    return (torch.log(1e-5 + mel) + 4.0) / 4.0


class ReferenceStep:
    """Holds fp32 leaf parameters (reference state_dict layout) + torch AdamW / OneCycleLR and runs CPU steps."""

    def __init__(self, state_dict, cfg, lambda_f0=0.1, max_lr=3e-4, epochs=100, steps_per_epoch=1000, device="cpu",
                 autocast_dtype=None):
        """device: where the fp32 restatement is evaluated (the host for the CPU baseline; a CUDA device lets the tests
        compare long trajectories in seconds -- TF32 must then be disabled by the caller).  autocast_dtype: run the
        forward under torch.autocast (yardstick: what torch's own mixed precision does to the same trajectory)."""
        self.cfg, self.lambda_f0 = cfg, lambda_f0
        self.device, self.autocast_dtype = torch.device(device), autocast_dtype
        self.sd = {}
        for k, v in state_dict.items():
            t = v.detach().clone().to(self.device)
            if t.dtype.is_floating_point:
                t = t.float().contiguous()
                if "running" not in k and not k.endswith(".pe"):
                    t.requires_grad_(True)
            self.sd[k] = t
        leaves = [t for t in self.sd.values() if t.requires_grad]
        self.opt = torch.optim.AdamW(leaves, lr=1e-4, weight_decay=5e-4, betas=(0.9, 0.98), eps=1e-9)
        self.sched = torch.optim.lr_scheduler.OneCycleLR(self.opt, max_lr=max_lr, epochs=epochs,
                                                         steps_per_epoch=steps_per_epoch, pct_start=0.0,
                                                         final_div_factor=5)

    def step(self, waves, f0s, crops, dropout=True):
        """waves [B, L] fp32, f0s [B, T_full] Hz, crops [B] -> dict of python floats (trainer.py:250-252)."""
        items = []
        for w, f0, c in zip(waves, f0s, crops):
            mel = log_mel_torch(torch.as_tensor(w).to(self.device))
            f0a = logmel_np.align_length(np.asarray(f0), mel.shape[1])
            c = int(c)
            f0c = f0a[c:c + 192]
            items.append((mel[:, c:c + 192].cpu().numpy(), f0c, (f0c == 0).astype(np.float32)))
        mels, f0b, silb = (torch.from_numpy(a).to(self.device) for a in logmel_np.collate(items))
        self.opt.zero_grad(set_to_none=True)
        import contextlib
        ctx = (torch.autocast(self.device.type, dtype=self.autocast_dtype) if self.autocast_dtype is not None
               else contextlib.nullcontext())
        with ctx:
            cls, det = J.jdcnet_forward(self.sd, mels.transpose(-1, -2), self.cfg, training=True,
                                        p_scale=1.0 if dropout else 0.0, update_running=True)
            loss, lf, ls = J.losses(cls.float(), det.float(), f0b, silb, self.lambda_f0)
        loss.backward()
        self.opt.step()
        self.sched.step()
        return {"loss": loss.item(), "f0": lf.item(), "sil": ls.item()}
