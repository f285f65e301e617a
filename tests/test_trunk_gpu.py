"""Memory-bound conv-trunk passes (trunk.cu) vs plain torch fp32 references of the same op (GPU), through the C-ABI."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False  # the torch references below are meant to be fp32
torch.backends.cuda.matmul.allow_tf32 = False

c_int, c_ll, c_f, c_d, c_u, c_ull = (ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_double, ctypes.c_uint,
                                     ctypes.c_ulonglong)
SLOPE = 0.01


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(dtype)


def test_stem_conv_forward_statistics_and_weight_gradient(built_lib):
    """Conv2d(1->64, 3x3, pad 1) on a strided (transposed-view) input, the fused BatchNorm statistics and dw
    (reference model.py:24-25)."""
    from pitchextractor_b200._lib import call, ptr, stream
    B, T, Fm = 3, 37, 80  # ragged sizes: 3 * 37 * 80 pixels is not a multiple of any block shape
    mel = _rand((B, 1, Fm, T), 1, dtype=torch.float32)  # reference batch layout [B,1,80,T]
    x = mel.transpose(-1, -2)                           # model input view [B,1,T,80] (trainer.py:235)
    w = _rand((64, 9), 2, 0.3, torch.float32)
    y = torch.empty(B, T, Fm, 64, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(2, 64, device="cuda", dtype=torch.float64)
    call("pe_stem_conv_fwd", ptr(x), c_ll(x.stride(0)), c_ll(x.stride(2)), c_ll(x.stride(3)), c_int(B), c_int(T), c_int(Fm),
         ptr(w), ptr(y), ptr(stats), stream())
    ref = F.conv2d(x, w.view(64, 1, 3, 3), padding=1).permute(0, 2, 3, 1)  # NHWC
    assert torch.allclose(y.float(), ref, rtol=1e-2, atol=1e-2)
    yb = y.double().reshape(-1, 64)
    assert torch.allclose(stats[0], yb.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(stats[1], (yb * yb).sum(0), rtol=1e-5, atol=1e-3)
    dy = _rand((B, T, Fm, 64), 3)
    dw = torch.zeros(64, 9, device="cuda")
    call("pe_stem_conv_wgrad", ptr(x), c_ll(x.stride(0)), c_ll(x.stride(2)), c_ll(x.stride(3)), c_int(B), c_int(T), c_int(Fm),
         ptr(dy), ptr(dw), stream())
    xr = x.clone().requires_grad_(False)
    wr = w.view(64, 1, 3, 3).clone().requires_grad_(True)
    F.conv2d(xr, wr, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    ref_dw = wr.grad.view(64, 9)
    assert torch.allclose(dw, ref_dw, rtol=2e-3, atol=1e-3 * ref_dw.abs().max().item())


def _bn_pool_reference(x, gamma, beta, k, dout, aux_k=None, aux_dout=None):
    """torch autograd through BatchNorm2d(train) -> LeakyReLU -> MaxPool2d((1,k)) (+ an auxiliary MaxPool2d((1,aux_k)) of
    the BN INPUT, model.py:45-49) on NCHW fp32; returns y, dx, dgamma, dbeta, batch mean / rstd."""
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)  # [B, C, H, W]
    g = gamma.clone().requires_grad_(True)
    b = beta.clone().requires_grad_(True)
    mean = xr.mean((0, 2, 3))
    var = xr.var((0, 2, 3), unbiased=False)
    z = F.leaky_relu(F.batch_norm(xr, None, None, g, b, True, 0.1, 1e-5), SLOPE)
    y = F.max_pool2d(z, (1, k))
    loss = (y * dout.float().permute(0, 3, 1, 2)).sum()
    if aux_k:
        loss = loss + (F.max_pool2d(xr, (1, aux_k)) * aux_dout.float().permute(0, 3, 1, 2)).sum()
    loss.backward()
    return (y.detach().permute(0, 2, 3, 1), xr.grad.permute(0, 2, 3, 1), g.grad, b.grad, mean.detach(),
            torch.rsqrt(var.detach() + 1e-5))


@pytest.mark.parametrize("W,C,k,aux_k", [(80, 64, 1, 0), (40, 128, 2, 0), (10, 256, 4, 0), (20, 192, 2, 10), (6, 24, 2, 0)])
def test_bn_act_pool_forward_backward(built_lib, W, C, k, aux_k):
    """BN(train)+LeakyReLU+MaxPool(1,k) forward and its two-pass backward; W = 10, k = 4 leaves two columns unpooled
    (model.py:39); aux_k: the auxiliary max-pool gradient joins dx through the saved arg-max positions."""
    from pitchextractor_b200._lib import call, ptr, stream
    B, H = 2, 9
    rows, Wo = B * H, W // k
    x = _rand((B, H, W, C), 10)
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda") * 0.2
    dout = _rand((B, H, Wo, C), 11)
    aux_dout = _rand((B, H, W // aux_k, C), 12) if aux_k else None
    y_ref, dx_ref, dg_ref, db_ref, mean, rstd = _bn_pool_reference(x, gamma, beta, k, dout, aux_k, aux_dout)
    scale = gamma * rstd
    shift = beta - mean * scale
    y = torch.empty(B, H, Wo, C, device="cuda", dtype=torch.bfloat16)
    call("pe_bn_act_pool_fwd", ptr(x), c_ll(rows), c_int(W), c_int(C), c_int(k), ptr(scale), ptr(shift), c_f(SLOPE), c_u(0),
         c_f(1.0), c_ull(0), ptr(y), c_ll(C), c_int(0), None, None, stream())
    assert torch.allclose(y.float(), y_ref, rtol=2e-2, atol=2e-2)
    idx = None
    if aux_k:
        idx = torch.empty(B, H, W // aux_k, C, device="cuda", dtype=torch.uint8)
        aux_y = torch.empty(B, H, W // aux_k, C, device="cuda", dtype=torch.bfloat16)
        call("pe_bn_act_pool_fwd", ptr(x), c_ll(rows), c_int(W), c_int(C), c_int(aux_k), None, None, c_f(SLOPE), c_u(0),
             c_f(1.0), c_ull(0), ptr(aux_y), c_ll(C), c_int(0), None, ptr(idx), stream())
        assert torch.equal(aux_y, F.max_pool2d(x.float().permute(0, 3, 1, 2), (1, aux_k)).permute(0, 2, 3, 1).to(torch.bfloat16))
    sums = torch.zeros(2, C, device="cuda", dtype=torch.float64)
    coef = torch.zeros(2, C, device="cuda")
    dg = torch.zeros(C, device="cuda")
    db = torch.zeros(C, device="cuda")
    dx = torch.empty_like(x)
    call("pe_bn_act_pool_bwd", ptr(x), c_ll(rows), c_int(W), c_int(C), c_int(k), ptr(scale), ptr(shift), ptr(mean), ptr(rstd),
         c_f(SLOPE), c_u(0), c_f(1.0), c_ull(0), ptr(dout), c_ll(C), c_int(0), None, ptr(sums), c_int(0), ptr(coef), ptr(dg),
         ptr(db), ptr(idx), ptr(aux_dout), c_ll(C), c_int(0), c_int(aux_k), ptr(dx), stream())
    torch.cuda.synchronize()
    tol = 3e-2 * dx_ref.abs().max().item()
    assert (dx.float() - dx_ref).abs().max().item() <= tol, ((dx.float() - dx_ref).abs().max().item(), tol)
    assert torch.allclose(db, db_ref, rtol=2e-2, atol=2e-2 * db_ref.abs().max().item())
    assert torch.allclose(dg, dg_ref, rtol=2e-2, atol=2e-2 * dg_ref.abs().max().item())


def test_bn_act_pool_dropout_is_replayed_by_the_backward(built_lib):
    """Dropout after the pool (model.py:40): the backward regenerates the forward mask from (seed, element index)."""
    from pitchextractor_b200._lib import call, ptr, stream, drop_thresh
    B, H, W, C, k = 2, 5, 10, 256, 4
    rows, Wo = B * H, W // k
    x = _rand((B, H, W, C), 20)
    scale = torch.rand(C, device="cuda") + 0.5
    shift = torch.randn(C, device="cuda") * 0.2
    thr, sc = drop_thresh(0.5)
    y0 = torch.empty(B, H, Wo, C, device="cuda", dtype=torch.bfloat16)
    y1 = torch.empty_like(y0)
    for out, t in ((y0, 0), (y1, thr)):
        call("pe_bn_act_pool_fwd", ptr(x), c_ll(rows), c_int(W), c_int(C), c_int(k), ptr(scale), ptr(shift), c_f(SLOPE), c_u(t),
             c_f(sc if t else 1.0), c_ull(77), ptr(out), c_ll(C), c_int(0), None, None, stream())
    keep = y1 != 0
    frac = keep.float().mean().item()
    assert 0.45 < frac < 0.55, frac
    assert torch.allclose(y1.float()[keep], (y0.float() * sc)[keep], rtol=1e-2, atol=1e-2)
    # backward with dout = 1: dx is non-zero only under kept outputs (BN statistics terms removed via zero coefficients
    # is not possible through the ABI, so compare the routed gradient count instead)
    mean = torch.zeros(C, device="cuda")
    rstd = torch.ones(C, device="cuda")
    dout = torch.ones(B, H, Wo, C, device="cuda", dtype=torch.bfloat16)
    sums = torch.zeros(2, C, device="cuda", dtype=torch.float64)
    coef = torch.zeros(2, C, device="cuda")
    dg = torch.zeros(C, device="cuda")
    db = torch.zeros(C, device="cuda")
    dx = torch.empty_like(x)
    call("pe_bn_act_pool_bwd", ptr(x), c_ll(rows), c_int(W), c_int(C), c_int(k), ptr(scale), ptr(shift), ptr(mean), ptr(rstd),
         c_f(SLOPE), c_u(thr), c_f(sc), c_ull(77), ptr(dout), c_ll(C), c_int(0), None, ptr(sums), c_int(0), ptr(coef), ptr(dg),
         ptr(db), None, None, c_ll(0), c_int(0), c_int(0), ptr(dx), stream())
    torch.cuda.synchronize()
    # sum of routed gradients g = dout * mask * scale_drop * lrelu'  ->  dbeta; with lrelu' in {1, slope}
    routed = db.sum().item()
    upper = keep.float().sum().item() * sc
    assert SLOPE * upper * 0.99 <= routed <= upper * 1.01, (routed, upper)
