"""Achieved HBM bandwidth of the memory-bound trunk passes at the training shapes (B=64, T=192); inputs cycle over
enough buffers to defeat the 126 MB L2."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200._lib import call, ptr, stream
c_int, c_ll, c_f, c_u, c_ull = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_uint, ctypes.c_ulonglong
B, T = 64, 192
rows = B * T
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
NBUF = 4
def timeit(fn, reps=12):
    for i in range(2): fn(i % NBUF)
    torch.cuda.synchronize(); e0.record()
    for i in range(reps): fn(i % NBUF)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
tot_f = tot_b = 0.0
for (W, C, k, drop) in ((80, 64, 1, 0), (80, 64, 2, 0), (40, 128, 1, 0), (40, 128, 2, 0), (20, 192, 1, 0), (20, 192, 2, 0), (10, 256, 1, 0), (10, 256, 4, 1)):
    Wo = W // k
    xs = [torch.randn(rows, W, C, device="cuda").to(torch.bfloat16) for _ in range(NBUF)]
    outs = [torch.empty(rows, Wo, C, device="cuda", dtype=torch.bfloat16) for _ in range(NBUF)]
    dxs = [torch.empty(rows, W, C, device="cuda", dtype=torch.bfloat16) for _ in range(NBUF)]
    sc = torch.rand(C, device="cuda") + 0.5; sh = torch.randn(C, device="cuda") * 0.1
    mu = torch.randn(C, device="cuda") * 0.1; rs = torch.rand(C, device="cuda") + 0.5
    sums = torch.zeros(2, C, device="cuda", dtype=torch.float64); coef = torch.zeros(2, C, device="cuda")
    dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
    thr, scl = (32768, 2.0) if drop else (0, 1.0)
    f = lambda i: call("pe_bn_act_pool_fwd", ptr(xs[i]), c_ll(rows), c_int(W), c_int(C), c_int(k), ptr(sc), ptr(sh), c_f(0.01), c_u(thr), c_f(scl),
                       c_ull(7), ptr(outs[i]), c_ll(C), c_int(0), None, None, stream())
    us_f = timeit(f)
    bytes_f = rows * (W + Wo) * C * 2
    def bwd(i, ready):
        call("pe_bn_act_pool_bwd", ptr(xs[i]), c_ll(rows), c_int(W), c_int(C), c_int(k), ptr(sc), ptr(sh), ptr(mu), ptr(rs), c_f(0.01), c_u(thr),
             c_f(scl), c_ull(7), ptr(outs[i]), c_ll(C), c_int(0), None, ptr(sums), c_int(ready), ptr(coef), ptr(dg), ptr(db), None, None, c_ll(0), c_int(0), c_int(0), ptr(dxs[i]), stream())
    us_b1 = timeit(lambda i: bwd(i, 1))
    us_b0 = timeit(lambda i: bwd(i, 0))
    bytes_b = rows * (2 * W + Wo) * C * 2
    tot_f += us_f; tot_b += us_b1
    print("W=%2d C=%3d k=%d: fwd %6.1f us %5.2f TB/s | bwd apply (sums fused) %6.1f us %5.2f TB/s | bwd reduce+apply %6.1f us" % (
        W, C, k, us_f, bytes_f / us_f / 1e6, us_b1, bytes_b / us_b1 / 1e6, us_b0))
print("sum fwd %.0f us, bwd apply %.0f us" % (tot_f, tot_b))
