"""JDCNet with the reference's constructor, ``forward`` contract and ``state_dict`` layout (reference model.py:13-256),
executed by the sm_100a kernel engine (``engine.py``) instead of torch.nn layers.

The module tree below only *names and owns* parameters/buffers so that checkpoints interchange with the reference
(``conv_block.0.weight`` ... ``sequence_classifier.model.layers.3.norm2.bias`` ...).  All parameters live in ONE flat
fp32 arena (gradients in a second one): the optimizer, the bf16 weight cast and the data-parallel all-reduce are single
passes over contiguous memory.  Convolution weights are stored channels-last ([Cout][kh][kw][Cin] in memory, exposed
with the reference's [Cout, Cin, kh, kw] shape), which is the tap-major operand layout the implicit-GEMM kernels read.
"""
import math

import torch
from torch import nn

from .engine import Engine

TRUNK_CHANNELS = ((64, 128), (128, 192), (192, 256))  # ResBlock in/out (model.py:31-33)


class _Node(nn.Module):
    """Anonymous container; children are attached by name so dotted state_dict keys match the reference."""

    def forward(self, *a, **k):  # pragma: no cover - containers are never called
        raise RuntimeError("parameter container")


def _descend(root, path):
    node = root
    for name in path:
        if name not in node._modules:
            node.add_module(name, _Node())
        node = node._modules[name]
    return node


class SinusoidalPositionalEncoding(nn.Module):
    """Buffer-only module (reference model.py:178-193); the add is fused into the first LayerNorm kernel."""

    def __init__(self, d_model, max_len=2000):
        super().__init__()
        pos = torch.arange(max_len, dtype=torch.float32)[:, None]
        freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
        table = torch.zeros(max_len, d_model)
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", table[None])


class SequenceModel(nn.Module):
    """Parameter owner for the BiLSTM / Transformer temporal block (reference model.py:196-256)."""

    def __init__(self, input_size, model_type="bilstm", hidden_size=384, num_layers=2, dropout=0.3, bidirectional=True,
                 nhead=8, dim_feedforward=1024, max_len=2000):
        super().__init__()
        self.model_type = model_type.lower()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.bidirectional, self.nhead, self.dim_feedforward = bidirectional, nhead, dim_feedforward
        self.dropout = dropout
        # fail at construction, not at the first step on the GPU, for shapes the sm_100a kernels are not built for
        if self.model_type == "bilstm" and (hidden_size != 384 or not bidirectional or input_size != 512):
            raise NotImplementedError("the LSTM kernels are built for input_size 512, hidden_size 384, bidirectional "
                                      "(reference defaults, model.py:199-210); got input_size=%d hidden_size=%d "
                                      "bidirectional=%s" % (input_size, hidden_size, bidirectional))
        if self.model_type == "bilstm" and not 1 <= num_layers <= 16:
            raise NotImplementedError("1..16 LSTM layers are supported (dropout site ids)")
        if self.model_type == "transformer":  # registered before `model`, as in the reference (state_dict key order)
            if input_size != 512 or input_size % nhead or input_size // nhead != 64:
                raise NotImplementedError("the attention kernels are built for d_model 512 with head_dim 64 "
                                          "(d_model / nhead); got d_model=%d nhead=%d" % (input_size, nhead))
            if not 1 <= num_layers <= 7:
                raise NotImplementedError("1..7 encoder layers are supported (dropout site ids)")
            if dim_feedforward % 64:
                raise NotImplementedError("dim_feedforward must be a multiple of 64")
            self.pos_encoding = SinusoidalPositionalEncoding(input_size, max_len=max_len)
        self.model = _Node()
        if self.model_type == "bilstm":
            self.lstm_dropout = dropout if num_layers > 1 else 0.0
            ndir = 2 if bidirectional else 1
            for layer in range(num_layers):
                in_dim = input_size if layer == 0 else hidden_size * ndir
                for sfx in ("", "_reverse")[:ndir]:
                    w_ih = torch.empty(4 * hidden_size, in_dim)
                    w_hh = torch.empty(4 * hidden_size, hidden_size)
                    nn.init.orthogonal_(w_ih)
                    nn.init.orthogonal_(w_hh)
                    # registration order follows nn.LSTM: w_ih, w_hh, b_ih, b_hh per direction-layer
                    self.model.register_parameter("weight_ih_l%d%s" % (layer, sfx), nn.Parameter(w_ih))
                    self.model.register_parameter("weight_hh_l%d%s" % (layer, sfx), nn.Parameter(w_hh))
                    self.model.register_parameter("bias_ih_l%d%s" % (layer, sfx), nn.Parameter(torch.randn(4 * hidden_size)))
                    self.model.register_parameter("bias_hh_l%d%s" % (layer, sfx), nn.Parameter(torch.randn(4 * hidden_size)))
            self._output_dim = hidden_size * ndir
        elif self.model_type == "transformer":
            d, ff = input_size, dim_feedforward
            proto_in_proj = nn.init.xavier_uniform_(torch.empty(3 * d, d))  # shared by the cloned layers (model.py:239)
            for layer in range(num_layers):
                node = _descend(self.model, ("layers", str(layer)))
                attn = _descend(node, ("self_attn",))
                attn.register_parameter("in_proj_weight", nn.Parameter(proto_in_proj.clone()))
                attn.register_parameter("in_proj_bias", nn.Parameter(torch.zeros(3 * d)))
                self._linear(_descend(attn, ("out_proj",)), d, d)
                self._linear(_descend(node, ("linear1",)), ff, d)
                self._linear(_descend(node, ("linear2",)), d, ff)
                for nm in ("norm1", "norm2"):
                    ln = _descend(node, (nm,))
                    ln.register_parameter("weight", nn.Parameter(torch.ones(d)))
                    ln.register_parameter("bias", nn.Parameter(torch.zeros(d)))
            self.layer_norm = _Node()
            self.layer_norm.register_parameter("weight", nn.Parameter(torch.ones(d)))
            self.layer_norm.register_parameter("bias", nn.Parameter(torch.zeros(d)))
            self._output_dim = input_size
        else:
            raise ValueError(f"Unsupported sequence model type: {model_type}")

    @staticmethod
    def _linear(node, out_f, in_f):
        node.register_parameter("weight", nn.Parameter(nn.init.kaiming_uniform_(torch.empty(out_f, in_f))))
        node.register_parameter("bias", nn.Parameter(torch.zeros(out_f)))

    @property
    def output_dim(self):
        return self._output_dim

    def forward(self, x):
        raise RuntimeError("SequenceModel runs inside JDCNet's kernel engine; call JDCNet.forward")


def _bn(node, c):
    node.register_parameter("weight", nn.Parameter(torch.ones(c)))
    node.register_parameter("bias", nn.Parameter(torch.zeros(c)))
    node.register_buffer("running_mean", torch.zeros(c))
    node.register_buffer("running_var", torch.ones(c))
    node.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


def _conv(node, cout, cin, k):
    w = nn.init.xavier_normal_(torch.empty(cout, cin, k, k))
    node.register_parameter("weight", nn.Parameter(w.contiguous(memory_format=torch.channels_last)))


class ResBlock(nn.Module):
    """Parameter owner for one pre-activation residual block (reference model.py:143-175)."""

    def __init__(self, in_channels, out_channels, leaky_relu_slope=0.01):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.downsample = in_channels != out_channels
        if not self.downsample:
            raise ValueError("identity-shortcut ResBlocks are not used by JDCNet and are not built")
        self.pre_conv = _Node()
        _bn(_descend(self.pre_conv, ("0",)), in_channels)
        self.conv = _Node()
        _conv(_descend(self.conv, ("0",)), out_channels, in_channels, 3)
        _bn(_descend(self.conv, ("1",)), out_channels)
        _conv(_descend(self.conv, ("3",)), out_channels, out_channels, 3)
        self.conv1by1 = _Node()
        _conv(self.conv1by1, out_channels, in_channels, 1)

    def forward(self, x):
        raise RuntimeError("ResBlock runs inside JDCNet's kernel engine; call JDCNet.forward")


class JDCNet(nn.Module):
    """Joint detection / classification network (reference model.py:13-122), B200 kernel engine underneath.

    ``forward(x)`` with x ``[B, 1, T, 80]`` returns ``(classifier [B, T, num_class], detector [B, T])`` and is
    differentiable w.r.t. the parameters (one autograd node for the whole network).
    """

    def __init__(self, num_class=722, leaky_relu_slope=0.01, sequence_model_config=None):
        super().__init__()
        self.num_class = num_class
        self.leaky_relu_slope = leaky_relu_slope
        sequence_model_config = dict(sequence_model_config or {})
        self.conv_block = _Node()
        _conv(_descend(self.conv_block, ("0",)), 64, 1, 3)
        _bn(_descend(self.conv_block, ("1",)), 64)
        _conv(_descend(self.conv_block, ("3",)), 64, 64, 3)
        self.res_block1 = ResBlock(*TRUNK_CHANNELS[0])
        self.res_block2 = ResBlock(*TRUNK_CHANNELS[1])
        self.res_block3 = ResBlock(*TRUNK_CHANNELS[2])
        self.pool_block = _Node()
        _bn(_descend(self.pool_block, ("0",)), 256)
        self.detector_conv = _Node()
        _conv(_descend(self.detector_conv, ("0",)), 256, 640, 1)
        _bn(_descend(self.detector_conv, ("1",)), 256)
        sequence_model_config.setdefault("input_size", 512)
        self.sequence_classifier = SequenceModel(**sequence_model_config)
        self.sequence_detector = SequenceModel(**sequence_model_config)
        self.classifier = _Node()
        SequenceModel._linear(self.classifier, num_class, self.sequence_classifier.output_dim)
        self.detector = _Node()
        SequenceModel._linear(self.detector, 2, self.sequence_detector.output_dim)
        self._engine = None

    # ------------------------------------------------------------------ engine plumbing
    @property
    def engine(self):
        dev = self.conv_block._modules["0"].weight.device
        if dev.type != "cuda":
            raise RuntimeError("pitchextractor_b200.JDCNet computes on CUDA (sm_100a) only -- move the model to a B200 "
                               "with .to('cuda'); there is no CPU fallback")
        if self._engine is None or self._engine.device != dev:
            self._engine = Engine(self, dev)
        return self._engine

    def forward(self, x):
        return self.engine.autograd_forward(x)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        if self._engine is not None:
            self._engine.invalidate_bf16()  # the fp32 master weights changed under the bf16 working copy
        return out

    def train_step_loss(self, mel, f0, sil, lambda_f0=0.1, grad_scale=1.0):
        """Fused forward + losses + backward for Trainer.run: fills ``.grad`` of every parameter and returns a
        device tensor [total, lambda*f0, sil]."""
        return self.engine.train_step(mel, f0, sil, lambda_f0, grad_scale)
