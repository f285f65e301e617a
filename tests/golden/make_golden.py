"""Generate the golden fixtures from the LIVE reference (run in the build container, where /root/reference exists):

    python tests/golden/make_golden.py

Writes tests/golden/logmel_golden.npz and tests/golden/jdcnet_golden.npz.  Inputs are seeded and re-creatable without
the reference (tests/golden_inputs.py); outputs come from the unmodified reference modules imported through
oracle/refshim.py: MelDataset._build_training_example / Collater (meldataset.py:629-677,804-826), JDCNet.forward
(model.py:75-122) and the Trainer.run loss arithmetic (trainer.py:237-239).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import refshim  # noqa: E402
import golden_inputs as GI  # noqa: E402


def main():
    ns = refshim.load()
    # ---------------- log-mel through the reference dataset code
    ds = ns.meldataset.MelDataset([], verbose=False)
    out = {}
    for name, wave in GI.logmel_signals().items():
        f0 = GI.f0_track(name)
        np.random.seed(GI.CROP_SEED)  # the reference crops with the global numpy RNG (meldataset.py:669)
        mel, f0_t, sil = ds._build_training_example(wave, 24000, f0, cache_key=None, allow_cache=False)
        out["mel_" + name] = mel.numpy()
        out["f0_" + name] = f0_t.numpy()
        out["sil_" + name] = sil.numpy()
        full = (torch.log(1e-5 + ds.to_melspec(torch.from_numpy(wave))) + 4) / 4
        out["full_" + name] = full.numpy()
    batch = [(torch.from_numpy(out["mel_" + n]), torch.from_numpy(out["f0_" + n]), torch.from_numpy(out["sil_" + n]))
             for n in ("noise", "short")]
    mels, f0s, sils = ns.meldataset.Collater()(batch)
    out["collate_mels"], out["collate_f0s"], out["collate_sils"] = mels.numpy(), f0s.numpy(), sils.numpy()
    al = ds.f0_extractor.align_length
    for i, (vals, n) in enumerate(GI.align_cases()):
        out["align_%d" % i] = al(vals, n)
    np.savez_compressed(os.path.join(HERE, "logmel_golden.npz"), **out)
    # ---------------- JDCNet forward / loss / gradient norms through the reference model
    out = {}
    for mt in ("transformer", "bilstm"):
        sd = GI.model_state_dict(mt)
        cfg = GI.model_config(mt)
        ref = ns.model.JDCNet(num_class=1, sequence_model_config=dict(cfg))
        ref.load_state_dict(sd)
        for mod in ref.modules():  # dropout off: RNG streams cannot be shared between implementations
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, (torch.nn.LSTM, torch.nn.MultiheadAttention)):
                mod.dropout = 0.0
        mel, f0, sil = GI.model_inputs()
        out[mt + "_sd_checksum"] = np.array([sum(float(v.double().abs().sum()) for v in sd.values())])
        ref.eval()
        with torch.no_grad():
            c, d = ref(mel.transpose(-1, -2))
        out[mt + "_eval_cls"], out[mt + "_eval_det"] = c.numpy(), d.numpy()
        ref.train()
        c, d = ref(mel.transpose(-1, -2))
        l1 = torch.nn.SmoothL1Loss()
        ce = torch.nn.BCEWithLogitsLoss()
        loss_f0 = 0.1 * l1(c.squeeze(), f0)
        loss_sil = ce(d, sil)
        (loss_f0 + loss_sil).backward()
        out[mt + "_train_cls"], out[mt + "_train_det"] = c.detach().numpy(), d.detach().numpy()
        out[mt + "_losses"] = np.array([loss_f0.item() + loss_sil.item(), loss_f0.item(), loss_sil.item()])
        names = [n for n, _ in ref.named_parameters()]
        out[mt + "_grad_norms"] = np.array([p.grad.norm().item() for _, p in ref.named_parameters()])
        out[mt + "_grad_names"] = np.array(names)
        # a few full gradient tensors (small ones) for direction checks
        for n, p in ref.named_parameters():
            if n in ("conv_block.0.weight", "classifier.weight", "detector.weight", "pool_block.0.weight"):
                out[mt + "_grad_" + n] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "jdcnet_golden.npz"), **out)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
