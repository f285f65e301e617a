"""Cycles per k-block of the tile engine as a function of tile width / K / stages (tuning aid, GPU only)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import ops

def run(M, N, K, reps=5, **kw):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm(a, b, out, M, N, K, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gemm(a, b, out, M, N, K, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

for N in (64, 128, 256):
    for K in (512, 4096):
        M = 148 * 128 * 4
        t = run(M, N, K)
        tiles_per_cta = (M // 128) * max(1, N // 256) / 148
        kb = K // 64
        print("N=%3d K=%4d  %.1f us  %.0f TFLOP/s  cycles/k-block ~ %.0f (at 1.9 GHz)" % (
            N, K, t * 1e6, 2.0 * M * N * K / t / 1e12, t * 1.9e9 / (tiles_per_cta * kb)), flush=True)
