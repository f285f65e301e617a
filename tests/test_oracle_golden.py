"""Pin the oracle (oracle/) against the golden fixtures generated from the live reference (tests/golden/make_golden.py)
and, when the reference tree is present (build container), against the live reference itself."""
import os

import numpy as np
import pytest
import torch

import golden_inputs as GI
from oracle import jdcnet_torch as J
from oracle import logmel_np, refshim, train_step

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def lg():
    return np.load(os.path.join(GOLD, "logmel_golden.npz"))


@pytest.fixture(scope="module")
def jg():
    return np.load(os.path.join(GOLD, "jdcnet_golden.npz"), allow_pickle=False)


def test_logmel_oracle_vs_golden(lg):
    for name, wave in GI.logmel_signals().items():
        ref = lg["full_" + name]
        got = logmel_np.log_mel(wave)
        assert got.shape == ref.shape
        # the golden is torchaudio fp32; the oracle is fp64 -> torchaudio's own rounding is the yardstick
        tol = 1e-4 if name in ("noise", "silence", "short", "harm3") else 3e-4
        assert np.abs(got - ref).max() <= tol, name
        got32 = train_step.log_mel_torch(torch.from_numpy(wave)).numpy()
        assert np.abs(got32 - ref).max() <= 2e-5, name  # same fp32 arithmetic as the reference


def test_training_example_and_collate_vs_golden(lg):
    sigs = GI.logmel_signals()
    items = {}
    for name, wave in sigs.items():
        T = 1 + len(wave) // 300
        mel, f0, sil = logmel_np.build_training_example(wave, GI.f0_track(name), crop_start=GI.crop_start(T))
        assert mel.shape == lg["mel_" + name].shape, name
        np.testing.assert_array_equal(f0, lg["f0_" + name])
        np.testing.assert_array_equal(sil, lg["sil_" + name])
        tol = 1e-4 if name != "tone220" else 3e-4
        assert np.abs(mel - lg["mel_" + name]).max() <= tol
        items[name] = (mel, f0, sil)
    mels, f0s, sils = logmel_np.collate([items["noise"], items["short"]])
    assert mels.shape == (2, 1, 80, 192)
    np.testing.assert_array_equal(f0s, lg["collate_f0s"])
    np.testing.assert_array_equal(sils, lg["collate_sils"])
    assert np.abs(mels - lg["collate_mels"]).max() <= 1e-4
    assert (mels[1, 0, :, 81:] == 0).all()


def test_align_length_vs_golden(lg):
    from pitchextractor_b200.meldataset import align_length
    for i, (vals, n) in enumerate(GI.align_cases()):
        np.testing.assert_array_equal(logmel_np.align_length(vals, n), lg["align_%d" % i])
        np.testing.assert_array_equal(align_length(vals, n), lg["align_%d" % i])  # host-side product code too


@pytest.mark.parametrize("mt", ["transformer", "bilstm"])
def test_jdcnet_oracle_vs_golden(jg, mt):
    sd = GI.model_state_dict(mt)
    chk = sum(float(v.double().abs().sum()) for v in sd.values())
    assert abs(chk - jg[mt + "_sd_checksum"][0]) <= 1e-6 * chk, "seeded init drifted: regenerate the goldens"
    cfg = J.default_config(mt)
    mel, f0, sil = GI.model_inputs()
    cls, det = J.jdcnet_forward(sd, mel.transpose(-1, -2), cfg, training=False)
    assert np.abs(cls.numpy() - jg[mt + "_eval_cls"]).max() <= 2e-4
    assert np.abs(det.numpy() - jg[mt + "_eval_det"]).max() <= 2e-4
    res = J.loss_and_grads(sd, mel, f0, sil, cfg)
    assert np.abs(res["cls"].numpy() - jg[mt + "_train_cls"]).max() <= 2e-4
    assert np.abs(res["det"].numpy() - jg[mt + "_train_det"]).max() <= 2e-4
    np.testing.assert_allclose([res["loss"].item(), res["f0"].item(), res["sil"].item()], jg[mt + "_losses"], rtol=1e-5)
    names = [str(n) for n in jg[mt + "_grad_names"]]
    norms = jg[mt + "_grad_norms"]
    for n, ref_norm in zip(names, norms):
        got = res["grads"][n].norm().item()
        assert abs(got - ref_norm) <= 2e-3 * ref_norm + 1e-7, (n, got, ref_norm)
    for n in ("conv_block.0.weight", "classifier.weight", "detector.weight", "pool_block.0.weight"):
        ref = jg[mt + "_grad_" + n]
        got = res["grads"][n].numpy()
        assert np.abs(got - ref).max() <= 2e-3 * np.abs(ref).max(), n


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
def test_oracle_vs_live_reference():
    ns = refshim.load()
    ds = ns.meldataset.MelDataset([], verbose=False)
    rng = np.random.default_rng(3)
    w = (0.1 * rng.standard_normal(58624)).astype(np.float32)
    ref = ((torch.log(1e-5 + ds.to_melspec(torch.from_numpy(w))) + 4) / 4).numpy()
    assert np.abs(logmel_np.log_mel(w) - ref).max() <= 1e-5
    assert np.abs(train_step.log_mel_torch(torch.from_numpy(w)).numpy() - ref).max() <= 1e-5
    from pitchextractor_b200.mel import mel_filterbank
    assert torch.equal(mel_filterbank(24000, 1024, 80), ds.to_melspec.mel_scale.fb)
    for mt in ("transformer", "bilstm"):
        cfg = GI.model_config(mt)
        torch.manual_seed(0)
        m = ns.model.JDCNet(num_class=1, sequence_model_config=dict(cfg))
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, (torch.nn.LSTM, torch.nn.MultiheadAttention)):
                mod.dropout = 0.0
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        x = torch.randn(2, 1, 192, 80)
        m.eval()
        with torch.no_grad():
            a, b = m(x)
        c, d = J.jdcnet_forward(sd, x, J.default_config(mt), training=False)
        assert (a - c).abs().max() <= 1e-4 and (b - d).abs().max() <= 1e-4
        m.train()
        a, b = m(x)
        c, d = J.jdcnet_forward(sd, x, J.default_config(mt), training=True, p_scale=0.0)
        assert (a - c).abs().max() <= 1e-4 and (b - d).abs().max() <= 1e-4
        # parameter container: same keys / shapes / order as the reference
        from pitchextractor_b200.model import JDCNet
        ours = JDCNet(num_class=1, sequence_model_config=dict(cfg)).state_dict()
        assert list(ours.keys()) == list(sd.keys())
        assert all(ours[k].shape == sd[k].shape for k in sd)
