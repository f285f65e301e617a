"""Per-launch time of the tile-engine calls in one training step (CUDA events, warm), in launch order."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import JDCNet, Trainer, build_optimizer, ops
cfg = dict(model_type="transformer", num_layers=4, dropout=0.1, nhead=8, dim_feedforward=1536, max_len=2048)
torch.manual_seed(0)
B = 64
model = JDCNet(num_class=1, sequence_model_config=cfg).cuda()
opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {}, "scheduler_params": {}})
tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda")
model.train()
batch = (torch.randn(B, 58624, device="cuda") * 0.1, torch.rand(B, 192, device="cuda") * 200, torch.zeros(B, 192, device="cuda"),
         torch.zeros(B, dtype=torch.int32, device="cuda"))
for _ in range(3):
    tr.run_async(batch)
torch.cuda.synchronize()
ops.PROFILE = []
tr.run_async(batch)
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
tot = 0
for i, (tag, a, b, fl) in enumerate(prof):
    ms = a.elapsed_time(b)
    tot += ms
    print("%3d %-6s %7.1f us  %6.1f GFLOP  %6.0f TFLOP/s" % (i, tag, ms * 1e3, fl / 1e9, fl / ms / 1e9))
print("total %.3f ms" % tot)
