"""Plain-torch fp32 functional restatement of the reference JDCNet + training losses (oracle; test infrastructure only).

Written against the reference sources (cited per function); it consumes a reference-layout ``state_dict`` so the same
weights drive the reference ``model.JDCNet``, this restatement and the CUDA engine.  Pinned against the live reference
in ``tests/test_oracle_vs_reference.py`` (bit-exact forward on CPU) and against ``tests/golden/jdcnet_*.npz``.

Dropout: the reference uses torch's RNG (model.py:40,56,224,235), which no other implementation can reproduce, so this
restatement takes ``p_scale`` (0 disables every dropout) and otherwise uses torch's own RNG.
"""
import math

import torch
import torch.nn.functional as F

LRELU = 0.01


def default_config(model_type="transformer"):
    # Configs/config.yml:16-24 (shared block) + SequenceModel ctor defaults (model.py:199-210)
    return dict(model_type=model_type, hidden_size=384, num_layers=4, dropout=0.1, bidirectional=True, nhead=8,
                dim_feedforward=1536, max_len=2048, input_size=512)


def _bn(sd, prefix, x, training, momentum=0.1, eps=1e-5, update_running=True):
    """nn.BatchNorm2d (model.py:25,37,54,150,159)."""
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training and not update_running:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training, momentum, eps)
    if training and update_running and (prefix + ".num_batches_tracked") in sd:
        sd[prefix + ".num_batches_tracked"] += 1
    return y


def _res_block(sd, p, x, training, upd):
    """ResBlock.forward (model.py:169-175): BN -> LReLU -> MaxPool(1,2) -> [conv, BN, LReLU, conv] + 1x1 shortcut."""
    x = F.max_pool2d(F.leaky_relu(_bn(sd, p + ".pre_conv.0", x, training, update_running=upd), LRELU), (1, 2))
    y = F.conv2d(x, sd[p + ".conv.0.weight"], padding=1)
    y = F.leaky_relu(_bn(sd, p + ".conv.1", y, training, update_running=upd), LRELU)
    y = F.conv2d(y, sd[p + ".conv.3.weight"], padding=1)
    if (p + ".conv1by1.weight") in sd:
        return y + F.conv2d(x, sd[p + ".conv1by1.weight"])
    return y + x


def _transformer(sd, p, x, cfg, training, p_scale):
    """SequenceModel.forward, transformer branch (model.py:253-255): LayerNorm(x + pe) then post-LN encoder layers
    (torch/nn/modules/transformer.py:961-982, norm_first=False, activation gelu, batch_first)."""
    B, T, D = x.shape
    H = cfg["nhead"]
    pd = cfg["dropout"] * p_scale if training else 0.0
    x = F.layer_norm(x + sd[p + ".pos_encoding.pe"][:, :T], (D,), sd[p + ".layer_norm.weight"], sd[p + ".layer_norm.bias"])
    for l in range(cfg["num_layers"]):
        q = p + ".model.layers.%d." % l
        qkv = F.linear(x, sd[q + "self_attn.in_proj_weight"], sd[q + "self_attn.in_proj_bias"])
        qh, kh, vh = [t.reshape(B, T, H, D // H).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
        att = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(D // H), dim=-1)
        att = F.dropout(att, pd, training)
        ctx = (att @ vh).transpose(1, 2).reshape(B, T, D)
        sa = F.linear(ctx, sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"])
        x = F.layer_norm(x + F.dropout(sa, pd, training), (D,), sd[q + "norm1.weight"], sd[q + "norm1.bias"])
        h = F.dropout(F.gelu(F.linear(x, sd[q + "linear1.weight"], sd[q + "linear1.bias"])), pd, training)
        ff = F.linear(h, sd[q + "linear2.weight"], sd[q + "linear2.bias"])
        x = F.layer_norm(x + F.dropout(ff, pd, training), (D,), sd[q + "norm2.weight"], sd[q + "norm2.bias"])
    return x


def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one nn.LSTM layer; gate order i, f, g, o (torch/nn/modules/rnn.py:842-847)."""
    B, T, _ = x.shape
    Hn = w_hh.shape[1]
    gi = F.linear(x, w_ih, b_ih + b_hh)
    h = x.new_zeros(B, Hn)
    c = x.new_zeros(B, Hn)
    outs = [None] * T
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        g = gi[:, t] + h @ w_hh.t()
        i, f, gg, o = g.chunk(4, dim=-1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, dim=1)


def _bilstm(sd, p, x, cfg, training, p_scale):
    """SequenceModel.forward, bilstm branch (model.py:250-252): nn.LSTM(batch_first, bidirectional, inter-layer dropout)."""
    pd = (cfg["dropout"] if cfg["num_layers"] > 1 else 0.0) * p_scale if training else 0.0
    for l in range(cfg["num_layers"]):
        outs = []
        for sfx, rev in (("", False), ("_reverse", True)):
            outs.append(_lstm_dir(x, sd["%s.model.weight_ih_l%d%s" % (p, l, sfx)], sd["%s.model.weight_hh_l%d%s" % (p, l, sfx)],
                                  sd["%s.model.bias_ih_l%d%s" % (p, l, sfx)], sd["%s.model.bias_hh_l%d%s" % (p, l, sfx)], rev))
        x = torch.cat(outs, dim=-1)
        if l < cfg["num_layers"] - 1:
            x = F.dropout(x, pd, training)
    return x


def jdcnet_forward(sd, x, cfg, training=True, p_scale=0.0, update_running=False):
    """JDCNet.forward (model.py:75-122).  x: [B, 1, T, 80] -> (classifier [B, T, num_class], detector [B, T])."""
    upd = update_running
    T = x.shape[-2]
    seq = _transformer if cfg["model_type"] == "transformer" else _bilstm
    # conv_block (model.py:23-28)
    y = F.conv2d(x, sd["conv_block.0.weight"], padding=1)
    y = F.leaky_relu(_bn(sd, "conv_block.1", y, training, update_running=upd), LRELU)
    c0 = F.conv2d(y, sd["conv_block.3.weight"], padding=1)
    r1 = _res_block(sd, "res_block1", c0, training, upd)
    r2 = _res_block(sd, "res_block2", r1, training, upd)
    r3 = _res_block(sd, "res_block3", r2, training, upd)
    # pool_block (model.py:36-41)
    pb = F.max_pool2d(F.leaky_relu(_bn(sd, "pool_block.0", r3, training, update_running=upd), LRELU), (1, 4))
    pb = F.dropout(pb, 0.5 * p_scale, training)
    cls_in = pb.permute(0, 2, 1, 3).reshape(-1, T, 512)
    cls = seq(sd, "sequence_classifier", cls_in, cfg, training, p_scale)
    cls = F.linear(cls, sd["classifier.weight"], sd["classifier.bias"])
    # detector branch (model.py:103-117)
    cat = torch.cat((F.max_pool2d(c0, (1, 40)), F.max_pool2d(r1, (1, 20)), F.max_pool2d(r2, (1, 10)), pb), dim=1)
    d = F.conv2d(cat, sd["detector_conv.0.weight"])
    d = F.leaky_relu(_bn(sd, "detector_conv.1", d, training, update_running=upd), LRELU)
    d = F.dropout(d, 0.5 * p_scale, training)
    det_in = d.permute(0, 2, 1, 3).reshape(-1, T, 512)
    det = seq(sd, "sequence_detector", det_in, cfg, training, p_scale)
    det = F.linear(det, sd["detector.weight"], sd["detector.bias"]).sum(-1)
    return cls, det


def losses(cls, det, f0, sil, lambda_f0=0.1):
    """trainer.py:237-239 with the criteria of train.py:104-106."""
    loss_f0 = lambda_f0 * F.smooth_l1_loss(cls.squeeze(), f0)
    loss_sil = F.binary_cross_entropy_with_logits(det, sil)
    return loss_f0 + loss_sil, loss_f0, loss_sil


def loss_and_grads(sd, mel, f0, sil, cfg, lambda_f0=0.1, training=True):
    """One forward/backward of Trainer.run (trainer.py:226-246) without AMP / optimizer; mel: [B, 1, 80, T]."""
    params = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k and not k.endswith(".pe"))
              for k, v in sd.items()}
    cls, det = jdcnet_forward(params, mel.transpose(-1, -2), cfg, training=training, p_scale=0.0)
    total, lf, ls = losses(cls, det, f0, sil, lambda_f0)
    leaves = {k: v for k, v in params.items() if v.requires_grad}
    grads = torch.autograd.grad(total, list(leaves.values()), allow_unused=True)
    return dict(loss=total.detach(), f0=lf.detach(), sil=ls.detach(), cls=cls.detach(), det=det.detach(),
                grads={k: g for k, g in zip(leaves.keys(), grads) if g is not None})
