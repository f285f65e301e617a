"""GPU time per C-ABI entry point inside one training step (CUDA events around every call, warm caches)."""
import collections, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pitchextractor_b200 import JDCNet, Trainer, build_optimizer, _lib

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
crops = torch.zeros(B, dtype=torch.int32, device="cuda")
batch = (waves, f0 * (1 - sil), sil, crops)
for _ in range(3):
    tr.run_async(batch)
torch.cuda.synchronize()
_lib.TIMING = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    tr.run_async(batch)
e1.record()
torch.cuda.synchronize()
tot = collections.defaultdict(float); cnt = collections.Counter()
for name, a, b in _lib.TIMING:
    tot[name] += a.elapsed_time(b); cnt[name] += 1
_lib.TIMING = None
total = e0.elapsed_time(e1) / n
print("%s B=%d: step %.3f ms (with event overhead); sum of calls %.3f ms" % (model_type, B, total, sum(tot.values()) / n))
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print("  %-24s %8.3f ms  %5.1f%%  calls/step %d" % (k, v / n, 100 * v / n / total, cnt[k] // n))
