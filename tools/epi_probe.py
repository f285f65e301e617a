import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import ops, _lib as L
dbg = torch.zeros(148, 16, dtype=torch.int64, device="cuda")
lib = L.lib()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def rep(tag, M, N, K, out_dtype=torch.bfloat16, **kw):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=out_dtype)
    for _ in range(3): ops.gemm(a, b, out, M, N, K, **kw)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20): ops.gemm(a, b, out, M, N, K, **kw)
    e1.record(); torch.cuda.synchronize()
    ops.DEBUG_BUFFER = dbg; dbg.zero_()
    ops.gemm(a, b, out, M, N, K, **kw); torch.cuda.synchronize(); ops.DEBUG_BUFFER = None
    d = dbg.float().mean(0).tolist()
    tiles = (M // 128) * ((N + 255) // 256) / 148
    print("%-34s %5.1f us | %.1f tiles/CTA | mma loop %6.0f (wait opnd %5.0f, acc %5.0f) | exit %6.0f | epi: wait %6.0f ld %6.0f work %6.0f last %6.0f" % (
        tag, e0.elapsed_time(e1) / 20 * 1e3, tiles, d[0], d[1], d[2], d[7], d[8], d[9], d[10], d[11]))
M = 12288
bias512 = torch.randn(512, device="cuda"); bias1536 = torch.randn(1536, device="cuda")
aux = torch.randn(M, 512, device="cuda").to(torch.bfloat16)
u = torch.empty(M, 1536, device="cuda", dtype=torch.bfloat16)
rep("outproj plain bf16", M, 512, 512)
rep("outproj plain f32", M, 512, 512, out_dtype=torch.float32)
rep("outproj bias f32", M, 512, 512, out_dtype=torch.float32, bias=bias512)
rep("outproj bias+drop f32", M, 512, 512, out_dtype=torch.float32, bias=bias512, p_drop=0.1, seed=5)
rep("outproj bias+aux f32", M, 512, 512, out_dtype=torch.float32, bias=bias512, aux=aux, aux_mode=L.PE_AUX_ADD)
rep("outproj full f32", M, 512, 512, out_dtype=torch.float32, bias=bias512, p_drop=0.1, seed=5, aux=aux, aux_mode=L.PE_AUX_ADD)
rep("ffn1 plain", M, 1536, 512)
rep("ffn1 bias", M, 1536, 512, bias=bias1536)
rep("ffn1 bias+gelu", M, 1536, 512, bias=bias1536, act=L.PE_ACT_GELU)
rep("ffn1 bias+gelu_save", M, 1536, 512, bias=bias1536, act=L.PE_ACT_GELU_SAVE_GRAD, out2=u)
rep("ffn1 full", M, 1536, 512, bias=bias1536, act=L.PE_ACT_GELU_SAVE_GRAD, out2=u, p_drop=0.1, seed=5)
