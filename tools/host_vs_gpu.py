"""Host enqueue time vs GPU time of one training step, and back-to-back per-launch time of the step's GEMM shapes."""
import os, sys, time, torch
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
for _ in range(5):
    tr.run_async(batch)
torch.cuda.synchronize()
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n):
    tr.run_async(batch)
t1 = time.perf_counter(); e1.record()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.3f ms/step | gpu %.3f ms/step | wall %.3f ms/step" % ((t1 - t0) / n * 1e3, e0.elapsed_time(e1) / n, (t2 - t0) / n * 1e3))

def rep(M, N, K, reps=50, **kw):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, b, out, M, N, K)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps): ops.gemm(a, b, out, M, N, K)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    ref = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): torch.matmul(a, b.t(), out=ref)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): torch.matmul(a, b.t(), out=ref)
    e1.record(); torch.cuda.synchronize()
    us2 = e0.elapsed_time(e1) / reps * 1e3
    print("gemm %5d x %4d x %4d  back-to-back %6.1f us (%4.0f TFLOP/s) | cuBLAS %6.1f us (%4.0f TFLOP/s)" % (
        M, N, K, us, 2.0 * M * N * K / us / 1e6, us2, 2.0 * M * N * K / us2 / 1e6))
for shp in [(128, 256, 64), (12288, 1536, 512), (12288, 512, 512), (12288, 512, 1536), (12288, 256, 2048), (12288, 2048, 256)]:
    rep(*shp)
