"""Equal-steps training parity: the CUDA Trainer vs the CPU port of the reference step (oracle/train_step.py), same
initial state_dict, same synthetic waveform batches, same AdamW / OneCycleLR, then held-out pitch RMSE (cents), RPA,
VUV and detector voicing accuracy of both models (north_star: "pitch RMSE / voicing accuracy on a held-out synthetic set
matching the reference after equal steps").  Dropout is off on both sides (their RNG streams cannot be shared), so the
two trajectories differ only by bf16 tensor-core arithmetic vs fp32.  Writes gpurun_out/equal_steps_parity.json.

    python tools/equal_steps_parity.py [steps] [batch] [model_type]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import golden_inputs as GI
from oracle import jdcnet_torch as J, train_step as TS, logmel_np
from pitchextractor_b200 import JDCNet, Trainer, build_optimizer, synthetic, inference
from pitchextractor_b200.meldataset import align_length

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mt = sys.argv[3] if len(sys.argv) > 3 else "transformer"
MAX_LR = 5e-3  # large enough for the F0 regression to leave its initial plateau within the step budget
torch.set_num_threads(os.cpu_count() or 1)
sd = GI.model_state_dict(mt)
cfg = J.default_config(mt)

def batch(seed):
    waves, f0 = synthetic.make_batch(B, seed=seed)
    crops = (np.arange(B) + seed) % 4
    return waves, f0, crops.astype(np.int32)

ref = TS.ReferenceStep(sd, cfg, max_lr=MAX_LR, epochs=1, steps_per_epoch=max(steps, 2))
model = JDCNet(num_class=1, sequence_model_config=GI.model_config(mt))
model.load_state_dict(sd)
model = model.cuda()
opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {},
                              "scheduler_params": {"max_lr": MAX_LR, "pct_start": 0.0, "epochs": 1,
                                                   "steps_per_epoch": max(steps, 2)}})
tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda")
model.engine.dropout_enabled = False
model.train()
curve = []
t_cpu = t_gpu = 0.0
for s in range(steps):
    w, f, c = batch(1000 + s)
    t0 = time.perf_counter(); a = ref.step(w, f, c, dropout=False); t_cpu += time.perf_counter() - t0
    f0 = np.stack([align_length(f[b], f.shape[1])[c[b]:c[b] + 192] for b in range(B)]).astype(np.float32)
    sil = (f0 == 0).astype(np.float32)
    t0 = time.perf_counter(); g = tr.run(tuple(torch.from_numpy(x) for x in (w, f0, sil, c))); t_gpu += time.perf_counter() - t0
    curve.append({"step": s, "cpu": a, "cuda": g})
    if s % 10 == 0 or s == steps - 1:
        print("step %3d  cpu loss %.4f (f0 %.4f sil %.4f) | cuda loss %.4f (f0 %.4f sil %.4f)" % (
            s, a["loss"], a["f0"], a["sil"], g["loss"], g["f0"], g["sil"]), flush=True)

# held-out evaluation (eval mode, running BatchNorm statistics)
waves, f0_full = synthetic.make_batch(16, seed=777777)
f0_ref = np.stack([align_length(f0_full[b], f0_full.shape[1])[:192] for b in range(16)]).astype(np.float32)
mels = np.stack([logmel_np.log_mel(waves[b])[:, :192] for b in range(16)]).astype(np.float32)[:, None]
with torch.no_grad():
    sd_cpu = {k: v.detach() for k, v in ref.sd.items()}
    c_cls, c_det = J.jdcnet_forward(sd_cpu, torch.from_numpy(mels).transpose(-1, -2), cfg, training=False)
    model.eval()
    g_cls, g_det = model(torch.from_numpy(mels).cuda().transpose(-1, -2))
def metrics(cls, det):
    p = cls.squeeze(-1).float().cpu().numpy().reshape(-1)
    d = det.float().cpu().numpy().reshape(-1)
    r = f0_ref.reshape(-1)
    m = inference.compute_metrics(r, p)
    m["rmse_cents"] = inference.rms_cents_error(r, p)
    m["mae_hz_voiced"] = float(np.mean(np.abs(p[r > 0] - r[r > 0])))
    m["voicing_acc_detector"] = inference.voicing_accuracy(d, r)
    return m
res = {"model": mt, "steps": steps, "batch": B, "max_lr": MAX_LR, "cpu_seconds": t_cpu, "cuda_seconds": t_gpu,
       "heldout_cpu_port": metrics(c_cls, c_det), "heldout_cuda": metrics(g_cls, g_det),
       "final_loss_cpu": curve[-1]["cpu"], "final_loss_cuda": curve[-1]["cuda"],
       "max_rel_loss_gap": max(abs(x["cpu"]["loss"] - x["cuda"]["loss"]) / abs(x["cpu"]["loss"]) for x in curve),
       "curve": curve}
print(json.dumps({k: v for k, v in res.items() if k != "curve"}, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "equal_steps_parity_%s.json" % mt), "w"))
