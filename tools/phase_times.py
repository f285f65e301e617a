"""GPU time of the phases of one graph-replayed training step (PE_PHASES=1 cuts the captured step at the phase marks;
each segment still runs its two-stream graph branches).  Usage: python tools/phase_times.py [transformer|bilstm] [B]"""
import collections, os, sys, torch
os.environ["PE_PHASES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import JDCNet, Trainer, build_optimizer

model_type = sys.argv[1] if len(sys.argv) > 1 else "transformer"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = dict(model_type=model_type, num_layers=4, dropout=0.1, nhead=8, dim_feedforward=1536, max_len=2048)
torch.manual_seed(0)
model = JDCNet(num_class=1, sequence_model_config=cfg).cuda()
opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {}, "scheduler_params": {}})
tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda")
model.train()
waves = torch.randn(B, 58624, device="cuda") * 0.1
f0 = torch.rand(B, 192, device="cuda") * 200 + 100
sil = (torch.rand(B, 192, device="cuda") < 0.2).float()
batch = (waves, f0 * (1 - sil), sil, torch.zeros(B, dtype=torch.int32, device="cuda"))
for _ in range(8):
    tr.run_async(batch)
torch.cuda.synchronize()
eng = model.engine
eng.phase_events.clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    tr.run_async(batch)
e1.record()
torch.cuda.synchronize()
tot = collections.OrderedDict()
for evs in eng.phase_events:
    for tag, a, b in evs:
        tot[tag] = tot.get(tag, 0.0) + a.elapsed_time(b)
print("%s B=%d: %.3f ms per step (cut into %d graph segments)" % (model_type, B, e0.elapsed_time(e1) / n, len(tot)))
for tag, v in tot.items():
    print("  %-14s %7.3f ms" % (tag, v / n))
