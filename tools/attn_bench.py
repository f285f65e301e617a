"""Attention forward / backward time per call at the step's shape (B items x 8 heads, T = 192), CUDA events, warm."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, HD, T = 8, 64, 192
D = H * HD
qkv = torch.randn(B * T, 3 * D, device="cuda").to(torch.bfloat16)
dctx = (torch.randn(B * T, D, device="cuda") * 0.1).to(torch.bfloat16)
ctx = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device="cuda")
dqkv = torch.zeros(B * T, 3 * D, device="cuda", dtype=torch.bfloat16)
delta = torch.empty(B, H, T, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for p in (0.0, 0.1):
    for name, fn in (("fwd", lambda: ops.attn_fwd(qkv, B, T, H, ctx, lse, p, 7)),
                     ("bwd", lambda: ops.attn_bwd(qkv, ctx, dctx, lse, B, T, H, dqkv, delta, p, 7))):
        for _ in range(5): fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(50): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        fl = (4 if name == "fwd" else 10) * T * T * HD * B * H
        print("attn %s B=%d p_drop=%.1f: %6.1f us  %5.0f TFLOP/s" % (name, B, p, us, fl / us / 1e6))
