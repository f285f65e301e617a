"""Yardstick, not a test: the reference's forward/backward as plain PyTorch library calls (cuDNN / cuBLAS / SDPA-free
nn.functional ops of oracle/jdcnet_torch.py, i.e. the same op sequence reference model.py runs) on the SAME B200, with
bf16 autocast ("time the reference on the B200 via torch -- that is the real bar").  Run on the GPU box:
    python tests/torch_gpu_yardstick.py [--batch 64] [--model transformer]
Prints one JSON line.  Nothing in the product imports this file."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_inputs as GI  # noqa: E402
from oracle import jdcnet_torch as J  # noqa: E402


def _transformer_sdpa(sd, p, x, cfg, training, p_scale):
    """oracle/jdcnet_torch._transformer with the attention core replaced by torch's fused scaled_dot_product_attention
    (what nn.MultiheadAttention dispatches to on CUDA), so that the yardstick is not handicapped by the math path."""
    import torch.nn.functional as F
    B, T, D = x.shape
    H = cfg["nhead"]
    pd = cfg["dropout"] * p_scale if training else 0.0
    x = F.layer_norm(x + sd[p + ".pos_encoding.pe"][:, :T], (D,), sd[p + ".layer_norm.weight"], sd[p + ".layer_norm.bias"])
    for l in range(cfg["num_layers"]):
        q = p + ".model.layers.%d." % l
        qkv = F.linear(x, sd[q + "self_attn.in_proj_weight"], sd[q + "self_attn.in_proj_bias"])
        qh, kh, vh = [t.reshape(B, T, H, D // H).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
        ctx = F.scaled_dot_product_attention(qh, kh, vh, dropout_p=pd).transpose(1, 2).reshape(B, T, D)
        sa = F.linear(ctx, sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"])
        x = F.layer_norm(x + F.dropout(sa, pd, training), (D,), sd[q + "norm1.weight"], sd[q + "norm1.bias"])
        h = F.dropout(F.gelu(F.linear(x, sd[q + "linear1.weight"], sd[q + "linear1.bias"])), pd, training)
        ff = F.linear(h, sd[q + "linear2.weight"], sd[q + "linear2.bias"])
        x = F.layer_norm(x + F.dropout(ff, pd, training), (D,), sd[q + "norm2.weight"], sd[q + "norm2.bias"])
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--model", default="transformer")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--channels-last", action="store_true")
    args = ap.parse_args()
    J._transformer = _transformer_sdpa
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True  # reference train.py:28
    cfg = J.default_config(args.model)
    sd = {k: v.to(dev) for k, v in GI.model_state_dict(args.model).items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype.is_floating_point and "running" not in k and not k.endswith(".pe")}
    state = dict(sd)
    state.update(params)
    opt = torch.optim.AdamW(list(params.values()), lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=5e-4, fused=True)
    B = args.batch
    wave = torch.randn(B, 58624, device=dev) * 0.1
    f0 = torch.rand(B, 192, device=dev) * 300
    sil = (f0 < 60).float()
    window = torch.hann_window(1024, device=dev)
    fb = torch.rand(513, 80, device=dev)

    def step():
        spec = torch.stft(wave, 1024, 300, 1024, window=window, center=True, pad_mode="reflect", return_complex=True)
        mel = (spec.abs().pow(2.0).transpose(-1, -2) @ fb).transpose(-1, -2)[:, None, :, :192]
        mel = (torch.log(1e-5 + mel) + 4.0) / 4.0
        opt.zero_grad(set_to_none=True)
        x_in = mel.transpose(-1, -2)
        if args.channels_last:
            x_in = x_in.contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            cls, det = J.jdcnet_forward(state, x_in, cfg, training=True, p_scale=1.0, update_running=True)
            total, _, _ = J.losses(cls.float(), det.float(), f0, sil)
        total.backward()
        opt.step()
        return total

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "torch library ops on cuda (bf16 autocast, eager, fused SDPA%s)" % (", channels_last" if args.channels_last else ""), "model": args.model, "batch": B,
                      "ms_per_step": ms, "segments_per_s": B / ms * 1e3, "loss": float(loss)}))


if __name__ == "__main__":
    main()
