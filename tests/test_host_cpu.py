"""Host-side logic that needs no GPU: C-ABI surface, data-parallel reducer under gloo (world_size 2), data pipeline."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "pitchextractor_b200.h")).read()
    declared = set(re.findall(r"^(?:int|long long) (pe_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert len(declared) >= 20
    lib = ctypes.CDLL(built_lib)
    for sym in declared:
        assert hasattr(lib, sym), sym
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (pe_[a-z0-9_]+)", out))
    assert exported == declared, (exported ^ declared)
    assert lib.pe_version() >= 100
    lib.pe_workspace_bytes.restype = ctypes.c_longlong
    ws = lambda op, B, T, L: lib.pe_workspace_bytes(op.encode(), B, T, L)
    assert ws("pe_logmel_tc", 3, 0, 58625) == 3 * 58628 * 4
    assert ws("pe_logmel_f32", 2, 196, 1024) == 2 * 196 * 513 * 4
    assert ws("pe_adamw", 1, 0, 0) == 0 and ws("pe_logmel_tc", 0, 0, 0) == -1


def test_no_cpu_fallback():
    from pitchextractor_b200 import JDCNet, LogMel
    m = JDCNet(num_class=1, sequence_model_config=dict(model_type="transformer", num_layers=1, dropout=0.1, nhead=8,
                                                       dim_feedforward=64, max_len=256))
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 192, 80))
    with pytest.raises(RuntimeError):
        LogMel("cpu")
    if not torch.cuda.is_available():
        lib = ctypes.CDLL(os.path.join(ROOT, "pitchextractor_b200", "libpe_b200.so"))
        assert lib.pe_check_device() != 0  # compute entry points refuse to run without an sm_100 device


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pitchextractor_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_bucket_ranges_cover_arena():
    from pitchextractor_b200.parallel import bucket_ranges
    names = ["conv_block.0.weight", "res_block1.conv.0.weight", "sequence_classifier.layer_norm.weight",
             "sequence_detector.layer_norm.weight", "classifier.weight", "detector.weight"]
    offs = [0, 64, 128, 192, 256, 320]
    b = bucket_ranges(names, offs, 384)
    assert [t for _, _, t in b] == ["sequence_detector+heads", "sequence_classifier", "trunk"]
    cover = sorted((lo, hi) for lo, hi, _ in b)
    assert cover[0][0] == 0 and cover[-1][1] == 384
    assert all(cover[i][1] == cover[i + 1][0] for i in range(len(cover) - 1))


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from pitchextractor_b200.parallel import init_from_env, GradReducer, bucket_ranges, broadcast_parameters
rank, world, _ = init_from_env("gloo")
names = ["conv_block.0.weight", "sequence_classifier.a", "sequence_detector.a", "classifier.weight"]
offs = [0, 100, 300, 700]
total = 1000
flat = torch.arange(total, dtype=torch.float32) * (rank + 1)
red = GradReducer(flat, bucket_ranges(names, offs, total))
red.begin_step()
red.ready("sequence_detector+heads")
red.ready("sequence_classifier")
red.wait()            # 'trunk' was never announced: wait() must still reduce it
expect = torch.arange(total, dtype=torch.float32) * sum(r + 1 for r in range(world))
assert torch.equal(flat, expect), (rank, (flat - expect).abs().max())
p = torch.full((10,), float(rank))
broadcast_parameters(p)
assert (p == 0).all()
dist.barrier()
print("rank", rank, "ok")
"""


def test_grad_reducer_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    port = 29000 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out


def test_dataset_wave_items_and_collate():
    from pitchextractor_b200 import Collater, MelDataset
    ds = MelDataset(["synthetic:%d|x\n" % i for i in range(3)], verbose=False, return_wave=True)
    assert len(ds) == 3 and ds.segment_samples == 58624
    items = [ds[i] for i in range(3)]
    wave, f0, sil, start = items[0]
    assert wave.shape == (58624,) and f0.shape == (192,) and sil.shape == (192,) and 0 <= start < 4
    assert torch.equal(sil, (f0 == 0).float())
    waves, f0s, sils, crops, lengths = Collater(return_wave=True)(items)
    assert waves.shape == (3, 58624) and f0s.shape == (3, 192) and sils.shape == (3, 192) and crops.dtype == torch.int32
    assert lengths.tolist() == [58624] * 3


def test_synthetic_segments_are_seeded_and_labelled():
    from pitchextractor_b200 import synthetic
    w1, f1 = synthetic.make_batch(2, seed=5)
    w2, f2 = synthetic.make_batch(2, seed=5)
    assert np.array_equal(w1, w2) and np.array_equal(f1, f2)
    assert w1.shape == (2, 58624) and f1.shape == (2, 196)
    voiced = f1[f1 > 0]
    assert 80.0 < voiced.min() and voiced.max() < 340.0
    assert 0.05 < (f1 == 0).mean() < 0.4
    assert np.abs(w1).max() < 1.0


def test_metrics_match_reference_definitions():
    """RMSE-cents / RPA / RCA / VUV / OctaveError vs hand-computed cases and, when present, the reference's own
    Utils/dynamic_pitch_tools.py."""
    from pitchextractor_b200 import inference as I
    rng = np.random.default_rng(0)
    ref = (100.0 + 200.0 * rng.random(500)).astype(np.float32)
    ref[rng.random(500) < 0.2] = 0.0
    pred = ref * (2.0 ** (rng.normal(0.0, 30.0, 500) / 1200.0)).astype(np.float32)
    pred[::17] *= 2.0          # octave errors
    pred[ref == 0] = 3.0       # predicted unvoiced
    m = I.compute_metrics(ref, pred)
    voiced = ref > 0
    cents = 1200.0 * np.log2(pred[voiced] / ref[voiced])
    assert abs(m["RPA"] - np.mean(np.abs(cents) <= 50.0)) < 1e-6
    assert m["VUV"] == 1.0 and m["RCA"] >= m["RPA"] and 0.03 < m["OctaveError"] < 0.08
    assert abs(I.rms_cents_error(ref, pred) - np.sqrt(np.mean(cents ** 2))) < 1e-2
    from oracle import refshim
    if refshim.available():
        import importlib, sys
        sys.path.insert(0, refshim.REF_ROOT)
        T = importlib.import_module("Utils.dynamic_pitch_tools")
        assert abs(T.rms_cents_error(ref, pred) - I.rms_cents_error(ref, pred)) < 1e-6
        np.testing.assert_array_equal(T.hz_to_cents(ref), I.hz_to_cents(ref))
        np.testing.assert_allclose(T.circular_cents_distance(cents, 0 * cents), I.circular_cents_distance(cents, 0 * cents))
        step = np.concatenate([np.full(40, 220.0), np.full(60, 330.0)])
        lagged = np.concatenate([np.full(47, 220.0), 330.0 + 25.0 * np.exp(-np.arange(53) / 6.0)])
        assert T.estimate_tracking_delay_ms(step, lagged, 12.5) == I.estimate_tracking_delay_ms(step, lagged, 12.5)
        assert T.compute_overshoot_cents(step, lagged) == I.compute_overshoot_cents(step, lagged)


def test_lag_and_overshoot_known_answers():
    """estimate_tracking_delay_ms / compute_overshoot_cents (Utils/dynamic_pitch_tools.py:107-136) on hand cases."""
    from pitchextractor_b200 import inference as I
    rng = np.random.default_rng(1)
    ref = 20.0 * rng.standard_normal(400) + 200.0  # white around 200 Hz: a sharp correlation peak
    pred = np.concatenate([np.full(5, ref[0]), ref[:-5]])           # the prediction trails by 5 frames
    assert I.estimate_tracking_delay_ms(ref, pred, 12.5) == 62.5
    assert I.estimate_tracking_delay_ms(ref, ref, 10.0) == 0.0
    assert np.isnan(I.estimate_tracking_delay_ms(np.full(10, 3.0), ref[:10], 1.0))
    assert np.isnan(I.estimate_tracking_delay_ms(ref[:0], ref[:0], 1.0))
    step = np.concatenate([np.full(10, 220.0), np.full(10, 440.0)])
    assert abs(I.compute_overshoot_cents(step, step * 2.0 ** (100.0 / 1200.0)) - 100.0) < 1e-9
    assert np.isnan(I.compute_overshoot_cents(np.zeros(4), np.ones(4)))
