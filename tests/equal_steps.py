"""Equal-steps training parity harness (helper, not a test file): the CUDA Trainer vs the fp32 restatement of the reference
step (oracle/train_step.py: trainer.py:219-252, optimizers.py:54-76) from the same initial state_dict on the same
synthetic waveform batches, dropout off on both sides (RNG streams cannot be shared), followed by the held-out metrics
north_star names: pitch RMSE in cents (Utils/dynamic_pitch_tools.py:92-104) and voicing accuracy.

Three trajectories are produced:
  cuda      pitchextractor_b200 (bf16 tensor-core operands, fp32 accumulation / master weights)
  fp32      the oracle in strict fp32 (TF32 off) -- the reference arithmetic
  amp_bf16  the same oracle under torch.autocast(bfloat16) -- the yardstick: how far torch's OWN mixed precision drifts
            from fp32 over the same steps (training is chaotic: rounding differences grow along the trajectory, so the
            per-step tolerance is stated relative to this drift).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def heldout_metrics(cls, det, f0_ref):
    from pitchextractor_b200 import inference
    p = cls.squeeze(-1).float().cpu().numpy().reshape(-1)
    d = det.float().cpu().numpy().reshape(-1)
    r = f0_ref.reshape(-1)
    m = inference.compute_metrics(r, p)
    m["rmse_cents"] = inference.rms_cents_error(r, p)
    m["mae_hz_voiced"] = float(np.mean(np.abs(p[r > 0] - r[r > 0])))
    m["voicing_acc_detector"] = inference.voicing_accuracy(d, r)
    return m


def run(model_type="transformer", steps=300, B=16, max_lr=3e-4, pool=24, oracle_device="cuda", with_amp=True,
        heldout=32, log=None):
    import golden_inputs as GI
    from oracle import jdcnet_torch as J, train_step as TS, logmel_np
    from pitchextractor_b200 import JDCNet, Trainer, build_optimizer, synthetic
    from pitchextractor_b200.meldataset import align_length
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = GI.model_state_dict(model_type)
    cfg = J.default_config(model_type)
    sched = dict(max_lr=max_lr, epochs=1, steps_per_epoch=max(steps, 2))
    refs = {"fp32": TS.ReferenceStep(sd, cfg, device=oracle_device, **sched)}
    if with_amp:
        refs["amp_bf16"] = TS.ReferenceStep(sd, cfg, device=oracle_device, autocast_dtype=torch.bfloat16, **sched)
    model = JDCNet(num_class=1, sequence_model_config=GI.model_config(model_type))
    model.load_state_dict(sd)
    model = model.cuda()
    opt, sch = build_optimizer({"params": model.parameters(), "optimizer_params": {},
                                "scheduler_params": dict(sched, pct_start=0.0)})
    tr = Trainer(model=model, optimizer=opt, scheduler=sch, loss_config={"lambda_f0": 0.1}, device="cuda")
    model.engine.dropout_enabled = False
    model.train()
    batches = []
    for i in range(pool):
        waves, f0 = synthetic.make_batch(B, seed=1000 + i)
        crops = ((np.arange(B) + i) % 4).astype(np.int32)
        f0c = np.stack([align_length(f0[b], f0.shape[1])[crops[b]:crops[b] + 192] for b in range(B)]).astype(np.float32)
        batches.append((waves, f0, crops, f0c, (f0c == 0).astype(np.float32)))
    curve = []
    for s in range(steps):
        w, f, c, f0c, sil = batches[s % pool]
        row = {"step": s}
        for name, ref in refs.items():
            row[name] = ref.step(w, f, c, dropout=False)
        row["cuda"] = tr.run(tuple(torch.from_numpy(x) for x in (w, f0c, sil, c)))
        curve.append(row)
        if log and (s % 25 == 0 or s == steps - 1):
            log("step %3d " % s + " | ".join("%s %.4f (f0 %.4f sil %.4f)" % (k, v["loss"], v["f0"], v["sil"])
                                             for k, v in row.items() if k != "step"))
    # held-out evaluation (eval mode: BatchNorm running statistics, no dropout)
    waves, f0_full = synthetic.make_batch(heldout, seed=777777)
    f0_ref = np.stack([align_length(f0_full[b], f0_full.shape[1])[:192] for b in range(heldout)]).astype(np.float32)
    mels = np.stack([logmel_np.log_mel(waves[b])[:, :192] for b in range(heldout)]).astype(np.float32)[:, None]
    out = {"model": model_type, "steps": steps, "batch": B, "max_lr": max_lr, "curve": curve, "heldout": {}}
    with torch.no_grad():
        for name, ref in refs.items():
            sd_r = {k: v.detach() for k, v in ref.sd.items()}
            x = torch.from_numpy(mels).to(ref.device).transpose(-1, -2)
            cls, det = J.jdcnet_forward(sd_r, x, cfg, training=False)
            out["heldout"][name] = heldout_metrics(cls, det, f0_ref)
        model.eval()
        cls, det = model(torch.from_numpy(mels).cuda().transpose(-1, -2))
        out["heldout"]["cuda"] = heldout_metrics(cls, det, f0_ref)
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old

    def gaps(a, b, key="loss"):
        return [abs(r[a][key] - r[b][key]) / max(abs(r[b][key]), 1e-12) for r in curve]

    out["gap_cuda_vs_fp32"] = gaps("cuda", "fp32")
    if with_amp:
        out["gap_amp_vs_fp32"] = gaps("amp_bf16", "fp32")
    return out


if __name__ == "__main__":
    import json
    mt = sys.argv[1] if len(sys.argv) > 1 else "transformer"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    lr = float(sys.argv[4]) if len(sys.argv) > 4 else 3e-4
    res = run(mt, steps, B, lr, log=lambda m: print(m, flush=True))
    g = np.array(res["gap_cuda_vs_fp32"])
    a = np.array(res.get("gap_amp_vs_fp32", g * 0))
    for lo in range(0, steps, 50):
        print("steps %3d-%3d: cuda-vs-fp32 gap mean %.4f max %.4f | torch-amp-vs-fp32 mean %.4f max %.4f" % (
            lo, min(lo + 50, steps) - 1, g[lo:lo + 50].mean(), g[lo:lo + 50].max(), a[lo:lo + 50].mean(),
            a[lo:lo + 50].max()))
    print(json.dumps(res["heldout"], indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "equal_steps_%s.json" % mt), "w"))
