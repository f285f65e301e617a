"""Narrow-output convolutions (Cout <= 128): time per call with the plain epilogue and with the fused column statistics
(BatchNorm forward sums; BatchNorm-backward first pass, direct and through MaxPool(1,2)).  """
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import ops
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def run(B, H, W, C1, C2, Cout):
    x = torch.randn(B, H, W, C1, device="cuda").to(torch.bfloat16)
    x2 = torch.randn(B, H, W, C2, device="cuda").to(torch.bfloat16) if C2 else None
    w = (torch.randn(Cout, 9 * C1 + C2, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    xb = torch.randn(B, H, W, Cout, device="cuda").to(torch.bfloat16)
    xw = torch.randn(B, H, 2 * W, Cout, device="cuda").to(torch.bfloat16)
    sums = torch.zeros(2, Cout, device="cuda", dtype=torch.float64)
    scale = torch.randn(Cout, device="cuda"); shift = torch.randn(Cout, device="cuda")
    cases = {"plain": {}, "stats1": dict(stats=sums),
             "stats2": dict(stats=sums, stats_mode=2, stats_x=xb, stats_scale=scale, stats_shift=shift, stats_slope=0.01),
             "stats3": dict(stats=sums, stats_mode=3, stats_x=xw, stats_scale=scale, stats_shift=shift, stats_slope=0.01)}
    res = []
    for name, kw in cases.items():
        for _ in range(3): ops.conv3x3(x, w, out, x2=x2, **kw)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): ops.conv3x3(x, w, out, x2=x2, **kw)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        res.append("%s %6.1f us (%4.0f TF/s)" % (name, us, 2.0 * B * H * W * Cout * (9 * C1 + C2) / us / 1e6))
    print("conv W=%d %d(+%d)->%d: " % (W, C1, C2, Cout) + " | ".join(res))

run(64, 192, 80, 64, 0, 64); run(64, 192, 40, 64, 0, 128); run(64, 192, 40, 128, 64, 128); run(64, 192, 40, 128, 0, 64)
