"""Attention kernels: tcgen05 (T = 192) and SIMT formulations vs a torch fp32 reference, and vs each other with the
(shared, hash-based) probability dropout enabled."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu
H, HD = 8, 64


def _ref(qkv, B, T, dctx=None):
    D = H * HD
    x = qkv.float().clone().requires_grad_(True)
    q, k, v = [t.reshape(B, T, H, HD).transpose(1, 2) for t in x.view(B, T, 3 * D).chunk(3, dim=-1)]
    att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    ctx = (att @ v).transpose(1, 2).reshape(B * T, D)
    if dctx is None:
        return ctx.detach(), None
    ctx.backward(dctx.float())
    return ctx.detach(), x.grad


def _run(qkv, B, T, dctx, p_drop=0.0, seed=0, simt=False):
    from pitchextractor_b200 import ops
    D = H * HD
    ctx = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device="cuda", dtype=torch.float32)
    dqkv = torch.zeros(B * T, 3 * D, device="cuda", dtype=torch.bfloat16)
    delta = torch.empty(B, H, T, device="cuda", dtype=torch.float32)
    ops.attn_fwd(qkv, B, T, H, ctx, lse, p_drop, seed, force_simt=simt)
    ops.attn_bwd(qkv, ctx, dctx, lse, B, T, H, dqkv, delta, p_drop, seed, force_simt=simt)
    torch.cuda.synchronize()
    return ctx, lse, dqkv


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.mark.parametrize("simt", [False, True])
@pytest.mark.parametrize("B,T", [(3, 192), (2, 64)])
def test_attention_vs_torch(built_lib, B, T, simt):
    g = torch.Generator(device="cuda").manual_seed(B * 100 + T)
    qkv = (torch.randn(B * T, 3 * H * HD, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    dctx = (torch.randn(B * T, H * HD, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    ctx, lse, dqkv = _run(qkv, B, T, dctx, simt=simt)
    rctx, rgrad = _ref(qkv, B, T, dctx)
    assert _rel(ctx, rctx) < 1e-2, _rel(ctx, rctx)
    D = H * HD
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        assert _rel(dqkv[:, sl], rgrad[:, sl]) < 2e-2, (name, _rel(dqkv[:, sl], rgrad[:, sl]))


def test_attention_dropout_tc_matches_simt(built_lib):
    B, T = 2, 192
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.randn(B * T, 3 * H * HD, device="cuda", generator=g).to(torch.bfloat16)
    dctx = (torch.randn(B * T, H * HD, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    a = _run(qkv, B, T, dctx, p_drop=0.1, seed=77, simt=False)
    b = _run(qkv, B, T, dctx, p_drop=0.1, seed=77, simt=True)
    c = _run(qkv, B, T, dctx, p_drop=0.0, seed=77, simt=False)
    assert _rel(a[0], b[0]) < 1e-2 and _rel(a[2], b[2]) < 2e-2, (_rel(a[0], b[0]), _rel(a[2], b[2]))
    assert (a[1] - b[1]).abs().max() < 1e-2
    assert _rel(a[0], c[0]) > 5e-2  # dropout really changes the output


def test_colsum(built_lib):
    from pitchextractor_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    for M, N in ((12288, 1536), (1000, 512), (77, 64)):
        x = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
        out = torch.zeros(N, device="cuda")
        ops.colsum(x, out)
        torch.cuda.synchronize()
        ref = x.float().sum(0)
        assert (out - ref).abs().max() <= 1e-3 * ref.abs().max() + 1e-2
