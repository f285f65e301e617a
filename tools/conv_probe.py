import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import ops, _lib as L
dbg = torch.zeros(148, 16, dtype=torch.int64, device="cuda")
lib = L.lib()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def conv(B, H, W, C1, C2, Cout, reps=10):
    x = torch.randn(B, H, W, C1, device="cuda").to(torch.bfloat16)
    x2 = torch.randn(B, H, W, C2, device="cuda").to(torch.bfloat16) if C2 else None
    w = torch.randn(Cout, 9 * C1 + C2, device="cuda").to(torch.bfloat16)
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    for _ in range(2): ops.conv3x3(x, w, out, x2=x2)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): ops.conv3x3(x, w, out, x2=x2)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    ops.DEBUG_BUFFER = dbg; dbg.zero_()
    ops.conv3x3(x, w, out, x2=x2); torch.cuda.synchronize(); ops.DEBUG_BUFFER = None
    d = dbg.float().mean(0).tolist()
    tiles = B * H * W / 128 / min(148, B * H * W / 128)
    kb = tiles * (9 * C1 // 64 + C2 // 64)
    fl = 2.0 * B * H * W * Cout * (9 * C1 + C2)
    print("conv B=%3d W=%2d C=%3d+%3d->%3d: %7.1f us %5.0f TFLOP/s | %6.1f k-blocks/CTA, %5.0f cyc/k-block | mma waits opnd %4.1f%% acc %4.1f%% | tma waits slot %4.1f%%" % (
        B, W, C1, C2, Cout, us, fl / us / 1e6, kb, d[0] / kb, 100 * d[1] / d[0], 100 * d[2] / d[0], 100 * d[3] / d[0]))
for B in (8, 64):
    conv(B, 192, 80, 64, 0, 64); conv(B, 192, 40, 64, 0, 128); conv(B, 192, 40, 128, 64, 128); conv(B, 192, 10, 256, 192, 256)
