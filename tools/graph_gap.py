"""Launch gap of the tile engine: the same GEMM back to back on a stream vs replayed from a CUDA graph."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import ops
M, N, K = 12288, 1536, 512
a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def run(n):
    for _ in range(n): ops.gemm(a, b, out, M, N, K)
run(5); torch.cuda.synchronize()
e0.record(); run(40); e1.record(); torch.cuda.synchronize()
print("stream launches: %.2f us per GEMM" % (e0.elapsed_time(e1) / 40 * 1e3))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run(3)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        run(40)
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print("graph replay:    %.2f us per GEMM" % (e0.elapsed_time(e1) / 40 * 1e3))
