"""A handful of tile-engine launches at the step's shapes (for ncu --set full captures): three forward convolutions,
two encoder GEMMs, one weight gradient; each launched twice (the second one is the warm one)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import ops

def conv(B, H, W, C1, C2, Cout):
    x = torch.randn(B, H, W, C1, device="cuda").to(torch.bfloat16)
    x2 = torch.randn(B, H, W, C2, device="cuda").to(torch.bfloat16) if C2 else None
    w = torch.randn(Cout, 9 * C1 + C2, device="cuda").to(torch.bfloat16)
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.conv3x3(x, w, out, x2=x2)

def gemm(M, N, K):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm(a, b, out, M, N, K)

def wgrad(B, H, W, C, Cout):
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, H, W, Cout, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(Cout, 9 * C, device="cuda")
    for _ in range(2):
        ops.conv_wgrad(dy, x, dw, taps=9)

conv(64, 192, 80, 64, 0, 64); conv(64, 192, 40, 128, 64, 128); conv(64, 192, 10, 256, 192, 256)
gemm(12288, 1536, 512); gemm(12288, 512, 1536); wgrad(64, 192, 20, 192, 192)
torch.cuda.synchronize()
print("done")
