"""Host-side jitter of the end-to-end epoch loop (Trainer.run_pipelined): wall time of repeated 20-step loops and the
per-yield gaps of the first one."""
import os, sys, time, gc, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import JDCNet, Trainer, build_optimizer
cfg = dict(model_type="transformer", num_layers=4, dropout=0.1, nhead=8, dim_feedforward=1536, max_len=2048)
torch.manual_seed(0)
model = JDCNet(num_class=1, sequence_model_config=cfg).cuda()
opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {}, "scheduler_params": {}})
tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda")
model.train()
B = 64
pool = []
for s in range(3):
    g = torch.Generator().manual_seed(s)
    w = (torch.randn(B, 58624, generator=g) * 0.1).pin_memory()
    sil = (torch.rand(B, 192, generator=g) < 0.2).float()
    f0 = ((torch.rand(B, 192, generator=g) * 200 + 100) * (1 - sil)).pin_memory()
    pool.append((w, f0, sil.pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()))
def host_loop(steps, stamps=None):
    for out in tr.run_pipelined(pool[i % 3] for i in range(steps)):
        if stamps is not None:
            stamps.append(time.perf_counter())
warm = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = [tuple(t.cuda() for t in b) for b in pool]
for i in range(5):
    tr.run_async(dev[i % 3])
torch.cuda.synchronize()
host_loop(warm)
res = []
for rep in range(6):
    torch.cuda.synchronize()
    st = []
    t0 = time.perf_counter(); host_loop(20, st); torch.cuda.synchronize(); t1 = time.perf_counter()
    res.append((t1 - t0) * 1e3)
    if rep == 0:
        print("gaps of the first loop (ms):", " ".join("%.1f" % ((b - a) * 1e3) for a, b in zip([t0] + st[:-1], st)))
print("20-step loops (ms):", " ".join("%.0f" % r for r in res))
