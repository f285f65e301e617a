"""Equal-steps training parity (north_star: "per-step loss ... within a stated bf16/fp32 tolerance, with pitch RMSE /
voicing accuracy on a held-out synthetic set matching the reference after equal steps").

The CUDA Trainer and the fp32 restatement of the reference step (strict fp32 on the GPU) start from the same state_dict
and see the same synthetic waveform batches for N optimisation steps at the reference's own learning rate
(Configs/config.yml:27, max_lr 3e-4, OneCycleLR pct_start 0), dropout off.  Harness: tests/equal_steps.py.

Stated tolerances
  * per step, F0 loss (lambda * SmoothL1, the dominant term): |cuda - fp32| <= 2e-3 * fp32
  * per step, total loss: |cuda - fp32| <= 2e-2 * fp32, or -- once rounding noise has been amplified along the
    trajectory (training is chaotic) -- no more than 1.5x the gap torch's own bf16 autocast shows at the same step window
  * held-out set (32 unseen segments, eval mode): RMSE-cents within 1 %, voicing accuracy within 3 points or within the
    bf16-autocast yardstick's own gap, RPA / VUV within 2 points
  * the model must have moved: F0 loss down by >= 10 % over the run.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _windows(g, n=25):
    return [float(np.mean(g[i:i + n])) for i in range(0, len(g), n)]


@pytest.mark.parametrize("model_type,steps,with_amp", [("transformer", 300, True), ("bilstm", 120, False)])
def test_equal_steps_trajectory_and_heldout_metrics(built_lib, model_type, steps, with_amp):
    import equal_steps as ES
    res = ES.run(model_type, steps=steps, B=16, max_lr=3e-4, with_amp=with_amp, log=print)
    curve, held = res["curve"], res["heldout"]
    # --- per-step losses
    for r in curve:
        assert abs(r["cuda"]["f0"] - r["fp32"]["f0"]) <= 2e-3 * abs(r["fp32"]["f0"]), r
    gap = np.array(res["gap_cuda_vs_fp32"])
    amp = np.array(res["gap_amp_vs_fp32"]) if with_amp else np.zeros_like(gap)
    print("total-loss gap per 25-step window  cuda:", ["%.4f" % v for v in _windows(gap)])
    if with_amp:
        print("total-loss gap per 25-step window  torch bf16 autocast:", ["%.4f" % v for v in _windows(amp)])
    assert gap[:20].max() <= 2e-2, gap[:20]
    for i, (gc, ga) in enumerate(zip(_windows(gap), _windows(amp))):
        assert gc <= max(2e-2, 1.5 * ga), (i, gc, ga)
    # --- the model moved
    f0_first = np.mean([r["fp32"]["f0"] for r in curve[:5]])
    f0_last = np.mean([r["fp32"]["f0"] for r in curve[-5:]])
    assert f0_last <= 0.9 * f0_first, (f0_first, f0_last)
    # --- held-out metrics
    a, b = held["cuda"], held["fp32"]
    print("held-out cuda", a)
    print("held-out fp32", b)
    assert abs(a["rmse_cents"] - b["rmse_cents"]) <= 1e-2 * b["rmse_cents"]
    assert abs(a["mae_hz_voiced"] - b["mae_hz_voiced"]) <= 1e-2 * b["mae_hz_voiced"]
    assert abs(a["RPA"] - b["RPA"]) <= 0.02 and abs(a["VUV"] - b["VUV"]) <= 0.02
    tol_v = 0.03
    if with_amp:
        print("held-out torch bf16 autocast", held["amp_bf16"])
        tol_v = max(tol_v, abs(held["amp_bf16"]["voicing_acc_detector"] - b["voicing_acc_detector"]))
    assert abs(a["voicing_acc_detector"] - b["voicing_acc_detector"]) <= tol_v
