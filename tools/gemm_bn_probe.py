"""A/B of the tile width for the encoder GEMM shapes (M = 12288): PE_TC_BN=256 vs 128 (GPU only)."""
import os, sys, subprocess
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from pitchextractor_b200 import ops, _lib as L
    M = 12288
    for N, K, kw in ((512, 512, dict(f32=True)), (512, 1536, dict(f32=True)), (1536, 512, {}), (512, 1536, {}), (512, 512, {})):
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.float32 if kw.get("f32") else torch.bfloat16)
        bias = torch.zeros(N, device="cuda")
        for _ in range(3): ops.gemm(a, b, out, M, N, K, bias=bias)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.gemm(a, b, out, M, N, K, bias=bias)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 20
        print("BN=%s N=%4d K=%4d out=%s: %.1f us %.0f TFLOP/s" % (os.environ.get("PE_TC_BN"), N, K, out.dtype, t * 1e3, 2.0 * M * N * K / t / 1e9), flush=True)
else:
    for bn in ("256", "128"):
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, PE_TC_BN=bn))
