"""Where the tile engine's MMA / TMA threads wait (tuning aid, GPU only)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import ops, _lib
dbg = torch.zeros(148, 16, dtype=torch.int64, device="cuda")
lib = _lib.lib()

def show(tag):
    torch.cuda.synchronize()
    d = dbg.float().mean(0).tolist()
    print("%-34s main-loop %8.0f cyc | mma waits operands %5.1f%% | mma waits accumulator %5.1f%% | tma waits slot %5.1f%%" % (
        tag, d[0], 100 * d[1] / max(d[0], 1), 100 * d[2] / max(d[0], 1), 100 * d[3] / max(d[0], 1)), flush=True)
    t0 = dbg[:, 4]
    print("      CTA entry skew %.2f us | setup %.0f cyc | entry->mma end %.0f cyc | entry->exit %.0f cyc (max %.0f)" % (
        (t0.max() - t0.min()).item() / 1e3, d[5], d[6], d[7], dbg[:, 7].max().item()), flush=True)

def gemm(M, N, K):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, out, M, N, K); torch.cuda.synchronize()
    ops.DEBUG_BUFFER = dbg; dbg.zero_()
    ops.gemm(a, b, out, M, N, K); show("gemm M=%d N=%d K=%d" % (M, N, K)); ops.DEBUG_BUFFER = None

def conv(B, H, W, C1, C2, Cout):
    x = torch.randn(B, H, W, C1, device="cuda").to(torch.bfloat16)
    x2 = torch.randn(B, H, W, C2, device="cuda").to(torch.bfloat16) if C2 else None
    w = torch.randn(Cout, 9 * C1 + C2, device="cuda").to(torch.bfloat16)
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    ops.conv3x3(x, w, out, x2=x2); torch.cuda.synchronize()
    ops.DEBUG_BUFFER = dbg; dbg.zero_()
    ops.conv3x3(x, w, out, x2=x2); show("conv W=%d C=%d->%d" % (W, C1, Cout)); ops.DEBUG_BUFFER = None

gemm(12288, 1536, 512); gemm(12288, 512, 1536); gemm(12288, 256, 2048)
conv(64, 192, 80, 64, 0, 64); conv(64, 192, 40, 128, 64, 128); conv(64, 192, 10, 256, 192, 256)

def wgrad(B, H, W, C, Cout, taps=9):
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, H, W, Cout, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(Cout, taps * C, device="cuda")
    ops.conv_wgrad(dy, x, dw, taps=taps); torch.cuda.synchronize()
    ops.DEBUG_BUFFER = dbg; dbg.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv_wgrad(dy, x, dw, taps=taps); e1.record()
    show("wgrad W=%d C=%d Cout=%d" % (W, C, Cout)); ops.DEBUG_BUFFER = None
    print("     %.1f us  %.0f TFLOP/s" % (e0.elapsed_time(e1) * 1e3, 2.0 * B * H * W * C * Cout * taps / e0.elapsed_time(e1) / 1e9))

wgrad(64, 192, 80, 64, 64); wgrad(64, 192, 40, 64, 128); wgrad(64, 192, 40, 128, 128); wgrad(64, 192, 20, 192, 192); wgrad(64, 192, 10, 256, 256)
