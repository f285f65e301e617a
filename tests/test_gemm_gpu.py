"""tcgen05 tile engine vs a plain torch fp32 reference of the same op (GPU)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


def _close(got, ref, rtol=2e-2, atol=None):
    got, ref = got.float(), ref.float()
    atol = atol if atol is not None else 1e-2 * ref.abs().max().item()
    err = (got - ref).abs()
    bad = err > atol + rtol * ref.abs()
    assert not bad.any(), "max err %.4g (ref max %.4g), %d bad" % (err.max().item(), ref.abs().max().item(),
                                                                   int(bad.sum()))


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 512), (384, 1536, 512), (200, 80, 72), (12288, 512, 1536)])
def test_gemm_kmajor(built_lib, M, N, K):
    from pitchextractor_b200 import ops
    Kp = (K + 7) // 8 * 8
    a = _rand((M, Kp), 1)[:, :K]
    b = _rand((N, Kp), 2)[:, :K]
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(a, b, out, M, N, K)
    torch.cuda.synchronize()
    _close(out, a.float() @ b.float().t())


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 192), (512, 1536, 1000)])
def test_gemm_mn_major(built_lib, a_mn, b_mn, M, N, K):
    from pitchextractor_b200 import ops
    a = _rand((M, K), 3)
    b = _rand((N, K), 4)
    a_arg = a.t().contiguous() if a_mn else a
    b_arg = b.t().contiguous() if b_mn else b
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a_arg, b_arg, out, M, N, K, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    _close(out, a.float() @ b.float().t())


def test_gemm_splitk_atomic(built_lib):
    from pitchextractor_b200 import ops, _lib as L
    M, N, K = 1536, 512, 4096  # weight-gradient shape: contraction over tokens
    dy = _rand((K, M), 5)
    x = _rand((K, N), 6)
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(dy, x, out, M, N, K, a_mn=True, b_mn=True, splits=8, out_mode=L.PE_OUT_F32_ATOMIC)
    torch.cuda.synchronize()
    _close(out, dy.float().t() @ x.float())


def test_gemm_epilogues(built_lib):
    from pitchextractor_b200 import ops, _lib as L
    M, N, K = 384, 1536, 512
    a, b = _rand((M, K), 7, 0.5), _rand((N, K), 8, 0.1)
    bias = torch.randn(N, device="cuda")
    ref = a.float() @ b.float().t() + bias
    # bias + GELU with pre-activation copy
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, out, M, N, K, bias=bias, act=L.PE_ACT_GELU, out2=pre)
    torch.cuda.synchronize()
    _close(pre, ref)
    _close(out, torch.nn.functional.gelu(ref))
    # residual add, fp32 out
    res = _rand((M, N), 9)
    out32 = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(a, b, out32, M, N, K, bias=bias, aux=res, aux_mode=L.PE_AUX_ADD)
    torch.cuda.synchronize()
    _close(out32, ref + res.float())
    # gelu-grad multiply
    ops.gemm(a, b, out32, M, N, K, aux=res, aux_mode=L.PE_AUX_GELU_GRAD)
    torch.cuda.synchronize()
    x = res.float().requires_grad_()
    torch.nn.functional.gelu(x).sum().backward()
    _close(out32, (a.float() @ b.float().t()) * x.grad)
    # gelu with saved derivative * dropout factor, and the multiply epilogue that consumes it
    dact = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, out, M, N, K, bias=bias, act=L.PE_ACT_GELU_SAVE_GRAD, out2=dact, p_drop=0.25, seed=99)
    torch.cuda.synchronize()
    xr = ref.clone().requires_grad_()
    torch.nn.functional.gelu(xr).sum().backward()
    keep = dact != 0
    assert abs(keep.float().mean().item() - 0.75) < 0.02
    _close(out.float()[keep], (torch.nn.functional.gelu(ref) / 0.75)[keep])
    _close(dact.float()[keep], (xr.grad / 0.75)[keep])
    assert (out[~keep] == 0).all()
    ops.gemm(a, b, out32, M, N, K, aux=dact, aux_mode=L.PE_AUX_MUL)
    torch.cuda.synchronize()
    _close(out32, (a.float() @ b.float().t()) * dact.float())
    # dropout: deterministic in (seed, index); keep-rate and scaling
    ops.gemm(a, b, out32, M, N, K, bias=bias, p_drop=0.25, seed=1234)
    out_b = torch.empty_like(out32)
    ops.gemm(a, b, out_b, M, N, K, bias=bias, p_drop=0.25, seed=1234)
    torch.cuda.synchronize()
    assert torch.equal(out32, out_b)
    kept = out32 != 0
    assert abs(kept.float().mean().item() - 0.75) < 0.01
    _close(out32[kept], (ref / 0.75)[kept])


@pytest.mark.parametrize("B,H,W,C1,C2,Cout", [(1, 16, 16, 64, 0, 64), (2, 192, 80, 64, 0, 64), (2, 192, 40, 128, 64, 128),
                                              (2, 192, 20, 192, 128, 192), (3, 192, 10, 256, 192, 256),
                                              (1, 10, 12, 64, 0, 128)])
def test_conv3x3(built_lib, B, H, W, C1, C2, Cout):
    from pitchextractor_b200 import ops
    x = _rand((B, H, W, C1), 10)
    w = _rand((Cout, 3, 3, C1), 11, 0.05)
    x2 = _rand((B, H, W, C2), 12) if C2 else None
    w2 = _rand((Cout, C2), 13, 0.05) if C2 else None
    wcat = torch.cat([w.reshape(Cout, -1)] + ([w2] if C2 else []), dim=1).contiguous()
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    ops.conv3x3(x, wcat, out, x2=x2)
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), padding=1)
    if C2:
        ref = ref + torch.nn.functional.conv2d(x2.float().permute(0, 3, 1, 2), w2.float()[:, :, None, None])
    _close(out, ref.permute(0, 2, 3, 1))


@pytest.mark.parametrize("B,H,W,C,Cout,taps", [(2, 192, 80, 64, 64, 9), (2, 192, 40, 64, 128, 9), (2, 192, 20, 192, 192, 9),
                                               (2, 192, 10, 256, 256, 9), (2, 192, 40, 64, 128, 1), (1, 10, 12, 64, 72, 9)])
def test_conv_wgrad(built_lib, B, H, W, C, Cout, taps):
    from pitchextractor_b200 import ops
    x = _rand((B, H, W, C), 14)
    dy = _rand((B, H, W, Cout), 15, 0.1)
    dw = torch.zeros(Cout, taps * C, device="cuda", dtype=torch.float32)
    ops.conv_wgrad(dy, x, dw, taps=taps)
    torch.cuda.synchronize()
    xs = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wt = torch.zeros(Cout, C, 3 if taps == 9 else 1, 3 if taps == 9 else 1, device="cuda", requires_grad=True)
    y = torch.nn.functional.conv2d(xs, wt, padding=1 if taps == 9 else 0)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = wt.grad.permute(0, 2, 3, 1).reshape(Cout, taps * C)
    _close(dw, ref)


def test_conv_epilogue_column_statistics(built_lib):
    """BatchNorm statistics / first BatchNorm-backward pass fused into the convolution epilogue."""
    from pitchextractor_b200 import ops
    B, H, W, C1, Cout = 2, 192, 40, 64, 128
    x = _rand((B, H, W, C1), 20)
    w = _rand((Cout, 9 * C1), 21, 0.05)
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    sums = torch.zeros(2, Cout, device="cuda", dtype=torch.float64)
    ops.conv3x3(x, w, out, stats=sums)
    torch.cuda.synchronize()
    o = out.double().reshape(-1, Cout)
    assert torch.allclose(sums[0], o.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(sums[1], (o * o).sum(0), rtol=1e-4, atol=1e-2)
    # mode 2: g = v * lrelu'(xb * scale + shift); sums = (sum g, sum g * xb)
    xb = _rand((B, H, W, Cout), 22)
    scale = torch.randn(Cout, device="cuda")
    shift = torch.randn(Cout, device="cuda") * 0.1
    sums2 = torch.zeros(2, Cout, device="cuda", dtype=torch.float64)
    ops.conv3x3(x, w, out, stats=sums2, stats_mode=2, stats_x=xb, stats_scale=scale, stats_shift=shift, stats_slope=0.01)
    torch.cuda.synchronize()
    z = xb.float().reshape(-1, Cout) * scale + shift
    g = out.float().reshape(-1, Cout) * torch.where(z > 0, 1.0, 0.01)
    assert torch.allclose(sums2[0], g.double().sum(0), rtol=1e-3, atol=1e-2)
    assert torch.allclose(sums2[1], (g.double() * xb.double().reshape(-1, Cout)).sum(0), rtol=1e-3, atol=1e-2)
    # mode 3: the same through MaxPool(1,2): the BN input is twice as wide, g goes to the first maximum of the pair
    xw = _rand((B, H, 2 * W, Cout), 23)
    sums3 = torch.zeros(2, Cout, device="cuda", dtype=torch.float64)
    ops.conv3x3(x, w, out, stats=sums3, stats_mode=3, stats_x=xw, stats_scale=scale, stats_shift=shift, stats_slope=0.01)
    torch.cuda.synchronize()
    pair = xw.float().reshape(-1, 2, Cout)
    pre = pair * scale + shift
    act = torch.where(pre > 0, pre, 0.01 * pre)
    second = act[:, 1] > act[:, 0]
    pre_s = torch.where(second, pre[:, 1], pre[:, 0])
    x_s = torch.where(second, pair[:, 1], pair[:, 0])
    g3 = out.float().reshape(-1, Cout) * torch.where(pre_s > 0, 1.0, 0.01)
    assert torch.allclose(sums3[0], g3.double().sum(0), rtol=1e-3, atol=1e-2)
    assert torch.allclose(sums3[1], (g3.double() * x_s.double()).sum(0), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("env", [{"PE_TC_PAIR": "1"}, {"PE_PDL": "1"}])
def test_opt_in_launch_modes(built_lib, env):
    """The opt-in launch modes of the tile engine -- CTA pairs (tcgen05 cta_group::2, clusters of two) and programmatic
    dependent launch -- are switched by environment variables read once per process, so they are exercised by re-running
    this file's GEMM / convolution / column-statistics checks in a child process."""
    import os
    import subprocess
    import sys
    child = dict(os.environ, **env)
    cmd = [sys.executable, "-m", "pytest", __file__, "-q", "-x", "-m", "gpu", "-k",
           "kmajor or mn_major or conv3x3 or column_statistics or epilogues", "-p", "no:cacheprovider"]
    r = subprocess.run(cmd, env=child, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
