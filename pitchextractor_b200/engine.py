"""Kernel engine of JDCNet: owns the flat parameter / gradient arenas and the activation buffers and enqueues the
sm_100a kernels of one forward (+ backward) pass on the current CUDA stream.

Data layout in HBM
  * parameters: one flat fp32 arena (master) + one flat bf16 arena (tensor-core operands, re-cast every step) + one
    flat fp32 gradient arena; every ``nn.Parameter`` / ``.grad`` is a view.
  * activations: NHWC bf16 ([B, T, F, C]; T = 192 frames on H, mel bins on W), sequence tensors [B*T, D] bf16, pre-LN
    residual sums fp32.  Everything the backward pass needs is kept (no recompute): 180 GB of HBM make activation
    checkpointing (reference trainer.py:227-233) unnecessary.
  * dropout masks are never stored: Philox keyed by (site seed, element index) is replayed in the backward kernels.
"""
import contextlib
import ctypes
import os

import torch

from . import _lib as L
from . import ops
from ._lib import call, ptr, stream

c_int, c_ll, c_f, c_d, c_u, c_ull = (ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_double, ctypes.c_uint,
                                     ctypes.c_ulonglong)
BN_EPS, BN_MOMENTUM, LN_EPS = 1e-5, 0.1, 1e-5
NULL = None


def _align(n, a=64):
    return (n + a - 1) // a * a


class _OutputGrad(torch.autograd.Function):
    """One autograd node for the whole network: forward ran in the engine, backward replays the engine's backward
    kernels and accumulates into the parameters' ``.grad`` views."""

    @staticmethod
    def forward(ctx, engine, token, *params):
        ctx.engine = engine
        ctx.token = token
        cls, det = engine._out_cls, engine._out_det
        return cls.clone(), det.clone()

    @staticmethod
    def backward(ctx, dcls, ddet):
        eng = ctx.engine
        if ctx.token != eng._fwd_token:
            raise RuntimeError("pitchextractor_b200: backward through a stale forward (activations were overwritten)")
        eng.backward_from_output_grads(dcls, ddet)
        return (None, None) + tuple(None for _ in range(len(eng.params)))


class Engine:
    def __init__(self, model, device):
        L.check(L.lib().pe_check_device(), "pe_check_device")
        self.model = model
        self.device = device
        self.slope = float(model.leaky_relu_slope)
        sc = model.sequence_classifier
        self.seq_type = sc.model_type
        self.seq_dim = sc.output_dim
        self.num_layers = sc.num_layers
        self.nhead = sc.nhead
        self.ff = sc.dim_feedforward
        self.p_seq = float(sc.dropout)
        self.p_trunk = 0.5  # nn.Dropout(p=0.5) in pool_block / detector_conv (model.py:40,56)
        # num_class == 1 (Configs/config.yml:17, the only value the reference's Trainer can train): fused heads + losses
        # kernel.  num_class > 1 (the reference constructor's default 722): the classifier head is a tile-engine GEMM,
        # available through forward() / autograd; the fused training step needs the scalar regression head.
        self.num_class = int(model.num_class)
        if self.seq_type == "bilstm" and (sc.hidden_size != 384 or not sc.bidirectional):
            raise NotImplementedError("the LSTM recurrence kernels are built for hidden_size 384, bidirectional")
        self.p_lstm = float(getattr(sc, "lstm_dropout", 0.0))
        # dropout site ids (the low 8 bits of a site's seed): 1, 2 trunk; 8 per encoder layer, the classifier stack from
        # 16, the detector stack right behind it; LSTM inter-layer dropouts from 128.  SequenceModel bounds num_layers
        # so that no two sites share an id.
        self._site_c = 16
        self._site_d = 16 + 8 * self.num_layers
        assert self._site_d + 8 * self.num_layers <= 128 or self.seq_type != "transformer"
        self.step_seed = 0x5EED0000
        self.salt_slot, self._salt = L.new_salt_slot(), 0
        self.dropout_enabled = True
        self._bufs = {}
        self._fwd_token = 0
        self.on_grads_ready = None  # callback(tag) for the data-parallel gradient reducer
        self._side_stream = None
        self._phase_cb = None     # see _mark()
        self._nvtx = os.environ.get("PE_NVTX") == "1"
        self._nvtx_open = False
        self.phase_events = None  # PE_PHASES=1: per replayed step, [(tag, start event, end event), ...]
        self._side_pending = False
        self.reducer = None         # GradReducer: begin_step() / ready(tag) / wait() bracket every training step
        # Training steps are captured into a CUDA graph after two eager warm-up steps and replayed from then on
        # (~270 launches per step; eagerly the host needs ~10 ms to enqueue them).  PE_CUDA_GRAPH=0 disables it.
        self.use_graph = os.environ.get("PE_CUDA_GRAPH", "1") != "0"
        # BiLSTM recurrences: persistent kernel (default) or one launch per time step (PE_LSTM_STEPWISE=1)
        self.lstm_persistent = os.environ.get("PE_LSTM_STEPWISE", "0") != "1"
        # persistent backward while its batch tiles (64 columns at most, 6 per launch) fit one launch; above that the
        # 128-column tiles lose to one launch per time step (20.2 vs 17.0 ms at B = 512)
        self.lstm_persistent_bwd_max_batch = int(os.environ.get("PE_LSTM_BWD_MAXB", "384"))
        self._graphs = {}
        self.bf16_fresh = False
        self._capturing = False
        self._pack()

    # ------------------------------------------------------------------ parameter arenas
    def _pack(self):
        named = list(self.model.named_parameters())
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += _align(p.numel())
        self.total = total
        dev = self.device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_bf16 = torch.zeros(total, device=dev, dtype=torch.bfloat16)
        self.view, self.gview, self.bview, self.offset = {}, {}, {}, {}
        for name, p, off in zip(self.names, self.params, offs):
            n = p.numel()
            if p.dim() == 4:  # conv weight: channels-last physical order [Cout][kh][kw][Cin]
                co, ci, kh, kw = p.shape
                mk = lambda buf: buf[off:off + n].view(co, kh, kw, ci).permute(0, 3, 1, 2)
            else:
                mk = lambda buf, shape=p.shape: buf[off:off + n].view(shape)
            v = mk(self.flat)
            v.copy_(p.data)
            p.data = v
            p.grad = mk(self.flat_grad)
            self.view[name], self.gview[name], self.bview[name], self.offset[name] = v, p.grad, mk(self.flat_bf16), off
        # fused operand layouts of the convolution weights (bf16)
        self.wops = {}
        for blk, cin, cout in (("conv_block", 64, 64),):
            self.wops[blk + ".3.dgrad"] = torch.empty(cin, 9 * cout, device=dev, dtype=torch.bfloat16)
        for i, (cin, cout) in enumerate(((64, 128), (128, 192), (192, 256)), 1):
            r = "res_block%d" % i
            self.wops[r + ".B.fwd"] = torch.empty(cout, 9 * cout + cin, device=dev, dtype=torch.bfloat16)
            self.wops[r + ".B.dgrad"] = torch.empty(cout, 9 * cout, device=dev, dtype=torch.bfloat16)
            self.wops[r + ".A.dgrad"] = torch.empty(cin, 9 * cout + cout, device=dev, dtype=torch.bfloat16)
        nbn = 9
        self.bn_sums = torch.zeros(2 * nbn, 2, 256, device=dev, dtype=torch.float64)  # fwd + bwd scratch per BN
        self.bn_aff = torch.zeros(nbn, 4, 256, device=dev, dtype=torch.float32)       # scale, shift, mean, rstd
        self.bn_coef = torch.zeros(nbn, 2, 256, device=dev, dtype=torch.float32)      # pass-2 coefficients (backward)
        self.bn_index = {}
        self.loss_acc = torch.zeros(2, device=dev, dtype=torch.float64)
        self.loss_out = torch.zeros(3, device=dev, dtype=torch.float32)
        from .optimizers import register_engine
        register_engine(self)

    def mat(self, name, rows, cols, arena="bf16"):
        """[rows, cols] row-major view of a parameter's storage in the bf16 / fp32 / grad arena."""
        buf = {"bf16": self.flat_bf16, "f32": self.flat, "grad": self.flat_grad}[arena]
        off = self.offset[name]
        return buf[off:off + rows * cols].view(rows, cols)

    def attach_grads(self):
        """(Re)attach the ``.grad`` views (an optimizer's zero_grad(set_to_none=True) drops them)."""
        for name, p in zip(self.names, self.params):
            if p.grad is None or p.grad.data_ptr() != self.gview[name].data_ptr():
                p.grad = self.gview[name]

    def zero_grad(self):
        self.flat_grad.zero_()
        self.attach_grads()

    def invalidate_bf16(self):
        """The fp32 master weights were written by something other than FusedAdamW.step (state_dict load, broadcast)."""
        self.bf16_fresh = False

    def cast_weights(self):
        """bf16 working copy of the fp32 master weights.  FusedAdamW.step writes it in the optimizer pass itself
        (``bf16_fresh``); one forward consumes that freshness, so any other writer of the weights (another optimizer,
        load_state_dict, in-place edits) is picked up by the cast of the following forward."""
        if not self.bf16_fresh:
            call("pe_cast_bf16", ptr(self.flat), ptr(self.flat_bf16), c_ll(self.total), stream())
        self.bf16_fresh = False

    def refresh_weights(self):
        """Tensor-core operand layouts of the convolution weights (from the fp32 master weights)."""
        V = self.view
        prep = lambda w, co, ci, w2, c2, fwd, dgrad: call(
            "pe_conv_weight_prep", ptr(w), c_int(co), c_int(ci), ptr(w2), c_int(c2), ptr(fwd), ptr(dgrad), stream())
        prep(V["conv_block.3.weight"], 64, 64, None, 0, None, self.wops["conv_block.3.dgrad"])
        for i, (cin, cout) in enumerate(((64, 128), (128, 192), (192, 256)), 1):
            r = "res_block%d" % i
            wA, wB, wS = V[r + ".conv.0.weight"], V[r + ".conv.3.weight"], V[r + ".conv1by1.weight"]
            prep(wB, cout, cout, wS, cin, self.wops[r + ".B.fwd"], None)
            prep(wB, cout, cout, None, 0, None, self.wops[r + ".B.dgrad"])
            prep(wA, cout, cin, wS, cout, None, self.wops[r + ".A.dgrad"])

    # ------------------------------------------------------------------ buffers
    def buf(self, name, shape, dtype=torch.bfloat16):
        t = self._bufs.get(name)
        if t is None or t.shape != torch.Size(shape) or t.dtype != dtype:
            if t is not None and self._graphs:
                # a captured step holds the address of the buffer that is about to be released (the batch size changed,
                # e.g. the short last batch of an epoch): drop the graphs, they are re-captured after two eager steps
                self._graphs = {}
            t = torch.empty(shape, device=self.device, dtype=dtype)
            self._bufs[name] = t
        return t

    def _seed(self, site):
        return (self.salt_slot << 56) | (((self.step_seed << 8) + site) & ((1 << 56) - 1))

    def _drop(self, p, training):
        if not training or not self.dropout_enabled or p <= 0.0:
            return 0, 1.0
        return L.drop_thresh(p)

    # ------------------------------------------------------------------ BatchNorm helpers
    def _bn_slot(self, prefix):
        if prefix not in self.bn_index:
            self.bn_index[prefix] = len(self.bn_index)
        return self.bn_index[prefix]

    def _bn_fwd_sums(self, prefix, training):
        """fp64 [2][C] accumulator a producing convolution fills with this BN's batch statistics (training only)."""
        return self.bn_sums[2 * self._bn_slot(prefix)] if training else None

    def _bn_prepare(self, prefix, x, rows, C, training, stats_fused=False):
        """Batch statistics (training) or running statistics (eval) -> scale / shift / mean / rstd for this BN."""
        i = self._bn_slot(prefix)
        aff = self.bn_aff[i]
        sc, sh, mu, rs = aff[0], aff[1], aff[2], aff[3]
        m = self.model.get_submodule(prefix)
        if training:
            sums = self.bn_sums[2 * i]
            if not stats_fused:
                call("pe_bn_stats", ptr(x), c_ll(rows), c_int(C), ptr(sums), stream())
            call("pe_bn_finalize", ptr(sums), c_d(float(rows)), ptr(m.weight), ptr(m.bias), c_f(BN_EPS),
                 c_f(BN_MOMENTUM), ptr(sc), ptr(sh), ptr(mu), ptr(rs), ptr(m.running_mean), ptr(m.running_var),
                 ptr(m.num_batches_tracked), c_int(C), stream())
        else:
            call("pe_bn_eval_params", ptr(m.weight), ptr(m.bias), ptr(m.running_mean), ptr(m.running_var), c_f(BN_EPS),
                 ptr(sc), ptr(sh), ptr(mu), ptr(rs), c_int(C), stream())
        return sc, sh, mu, rs

    def _act_pool(self, x, rows, W, C, k, aff, out=None, ld_out=0, c_off=0, out_seq=None, drop=(0, 1.0), seed=0,
                  argmax=None):
        sc, sh = (aff[0], aff[1]) if aff is not None else (None, None)
        call("pe_bn_act_pool_fwd", ptr(x), c_ll(rows), c_int(W), c_int(C), c_int(k), ptr(sc), ptr(sh), c_f(self.slope),
             c_u(drop[0]), c_f(drop[1]), c_ull(seed), ptr(out), c_ll(ld_out), c_int(c_off), ptr(out_seq), ptr(argmax),
             stream())

    def _bn_bwd_fused(self, prefix, x, pooled=False):
        """Epilogue arguments that make a convolution accumulate the first pass of this BN's backward (sum g, sum g*x)
        while it writes the gradient w.r.t. the BN output (k = 1) or w.r.t. its MaxPool(1,2) output (pooled); no dropout."""
        i = self._bn_slot(prefix)
        aff = self.bn_aff[i]
        return dict(stats=self.bn_sums[2 * i + 1], stats_mode=3 if pooled else 2, stats_x=x, stats_scale=aff[0],
                    stats_shift=aff[1], stats_slope=self.slope)

    def _act_pool_bwd(self, prefix, x, rows, W, C, k, dx, dout=None, ld_dout=0, c_off=0, dout_seq=None, drop=(0, 1.0),
                      seed=0, sums_ready=False, aux=None):
        """aux = (argmax, dout, ld, c_off, k_aux): gradient of the auxiliary max-pool of the same tensor, added here."""
        aux_idx, aux_dout, aux_ld, aux_off, aux_k = aux if aux is not None else (None, None, 0, 0, 0)
        i = self._bn_slot(prefix)
        aff = self.bn_aff[i]
        sums = self.bn_sums[2 * i + 1]
        g = self.gview
        call("pe_bn_act_pool_bwd", ptr(x), c_ll(rows), c_int(W), c_int(C), c_int(k), ptr(aff[0]), ptr(aff[1]),
             ptr(aff[2]), ptr(aff[3]), c_f(self.slope), c_u(drop[0]), c_f(drop[1]), c_ull(seed), ptr(dout),
             c_ll(ld_dout), c_int(c_off), ptr(dout_seq), ptr(sums), c_int(int(sums_ready)), ptr(self.bn_coef[i]),
             ptr(g[prefix + ".weight"]), ptr(g[prefix + ".bias"]), ptr(aux_idx), ptr(aux_dout), c_ll(aux_ld),
             c_int(aux_off), c_int(aux_k), ptr(dx), stream())

    # ------------------------------------------------------------------ forward
    def _prep_input(self, x):
        """Accept the reference layouts: model input [B,1,T,80] (possibly the transposed view of a [B,1,80,T] mel)."""
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError("JDCNet expects input of shape [B, 1, T, n_mels]")
        if x.device != self.device or x.dtype != torch.float32:
            x = x.to(self.device, torch.float32)
        B, _, T, F = x.shape
        if F != 80:
            raise ValueError("the conv trunk is built for 80 mel bins (width 80 -> 40 -> 20 -> 10 -> 2)")
        if T > 256:
            raise ValueError("sequence kernels are built for segments of at most 256 frames (reference uses 192)")
        return x, B, T, F

    def forward_core(self, x, training):
        x, B, T, F = self._prep_input(x)
        self._x, self._B, self._T = x, B, T
        self._training = training
        self._fwd_token += 1
        self.step_seed += 1
        if training:
            self.bn_sums.zero_()
        if not self._capturing:
            self.cast_weights()
        self.refresh_weights()
        V, W16 = self.view, self.bview
        BT = B * T
        st = stream
        # ---- conv_block (model.py:23-28)
        Y1 = self.buf("Y1", (B, T, 80, 64))
        call("pe_stem_conv_fwd", ptr(x), c_ll(x.stride(0)), c_ll(x.stride(2)), c_ll(x.stride(3)), c_int(B), c_int(T),
             c_int(F), ptr(V["conv_block.0.weight"]), ptr(Y1), ptr(self._bn_fwd_sums("conv_block.1", training)), st())
        aff = self._bn_prepare("conv_block.1", Y1, BT * 80, 64, training, stats_fused=True)
        Z1 = self.buf("Z1", (B, T, 80, 64))
        self._act_pool(Y1, BT, 80, 64, 1, aff, out=Z1, ld_out=64)
        R = self.buf("R0", (B, T, 80, 64))
        ops.conv3x3(Z1, self.mat("conv_block.3.weight", 64, 576), R,
                    stats=self._bn_fwd_sums("res_block1.pre_conv.0", training))
        # ---- residual blocks (model.py:143-175)
        width = 80
        for i, (cin, cout) in enumerate(((64, 128), (128, 192), (192, 256)), 1):
            r = "res_block%d" % i
            aff = self._bn_prepare(r + ".pre_conv.0", R, BT * width, cin, training, stats_fused=True)
            P = self.buf("P%d" % i, (B, T, width // 2, cin))
            self._act_pool(R, BT, width, cin, 2, aff, out=P, ld_out=cin)
            width //= 2
            U = self.buf("U%d" % i, (B, T, width, cout))
            ops.conv3x3(P, self.mat(r + ".conv.0.weight", cout, 9 * cin), U, stats=self._bn_fwd_sums(r + ".conv.1", training))
            aff = self._bn_prepare(r + ".conv.1", U, BT * width, cout, training, stats_fused=True)
            Vv = self.buf("V%d" % i, (B, T, width, cout))
            self._act_pool(U, BT, width, cout, 1, aff, out=Vv, ld_out=cout)
            R = self.buf("R%d" % i, (B, T, width, cout))
            nxt = "res_block%d.pre_conv.0" % (i + 1) if i < 3 else "pool_block.0"
            ops.conv3x3(Vv, self.wops[r + ".B.fwd"], R, x2=P, stats=self._bn_fwd_sums(nxt, training))
        # ---- pool_block + auxiliary max-pools + concat (model.py:36-49,103-108)
        CAT = self.buf("CAT", (B, T, 2, 640))
        SEQC = self.buf("SEQC", (BT, 512))
        aff = self._bn_prepare("pool_block.0", R, BT * 10, 256, training, stats_fused=True)
        self._act_pool(R, BT, 10, 256, 4, aff, out=CAT, ld_out=640, c_off=384, out_seq=SEQC,
                       drop=self._drop(self.p_trunk, training), seed=self._seed(1))
        # (the arg-max positions are kept, one byte each, so that the backward pass need not read R_k again)
        for kk, (Wk, Ck, kw, off) in enumerate(((80, 64, 40, 0), (40, 128, 20, 64), (20, 192, 10, 192))):
            self._act_pool(self._bufs["R%d" % kk], BT, Wk, Ck, kw, None, out=CAT, ld_out=640, c_off=off,
                           argmax=self.buf("AUXIDX%d" % kk, (BT, 2, Ck), torch.uint8))
        # ---- detector_conv (model.py:52-57): 1x1 conv == GEMM over the 640 concatenated channels
        DD = self.buf("DD", (BT * 2, 256))
        ops.gemm(CAT.view(BT * 2, 640), self.mat("detector_conv.0.weight", 256, 640), DD, BT * 2, 256, 640,
                 stats=self._bn_fwd_sums("detector_conv.1", training))
        aff = self._bn_prepare("detector_conv.1", DD, BT * 2, 256, training, stats_fused=True)
        SEQD = self.buf("SEQD", (BT, 512))
        self._act_pool(DD, BT, 2, 256, 1, aff, out=None, out_seq=SEQD, drop=self._drop(self.p_trunk, training),
                       seed=self._seed(2))
        self._mark("trunk_fwd")
        # ---- sequence models (model.py:94,113)
        if self.seq_type == "transformer":
            # the two encoder stacks are independent: the detector's runs on a second stream, so that the tail of one
            # stack's kernels (partial last wave, epilogue drain) is filled by the other's
            with self._forked() as side:
                with torch.cuda.stream(side):
                    Hd = self._transformer_fwd("sequence_detector", "d", SEQD, B, T, training, self._site_d)
                Hc = self._transformer_fwd("sequence_classifier", "c", SEQC, B, T, training, self._site_c)
        else:
            Hc, Hd = self._bilstm_fwd(SEQC, SEQD, B, T, training)
        self._Hc, self._Hd = Hc, Hd
        self._mark("encoder_fwd")
        return Hc, Hd

    _PHASES = ("trunk_fwd", "encoder_fwd", "heads_loss", "encoder_bwd", "trunk_bwd")

    def _mark(self, tag):
        """Phase boundary (all streams joined here): tools/phase_times.py cuts the captured step at these points; with
        PE_NVTX=1 the phases of an eagerly launched step (PE_CUDA_GRAPH=0) show up as NVTX ranges in a timeline."""
        if self._nvtx_open:  # (only inside a training step: forward_core alone also passes its marks)
            torch.cuda.nvtx.range_pop()
            i = self._PHASES.index(tag)
            self._nvtx_open = i + 1 < len(self._PHASES)
            if self._nvtx_open:
                torch.cuda.nvtx.range_push(self._PHASES[i + 1])
        if self._phase_cb is not None:
            self._phase_cb(tag)

    @contextlib.contextmanager
    def _forked(self, join=True):
        """A second stream that starts after everything enqueued so far on the current stream; joined back on exit, or
        (join=False) at the next _join_side().  Under CUDA-graph capture this becomes a parallel branch of the graph."""
        if os.environ.get("PE_TWO_STREAMS", "1") == "0":
            yield torch.cuda.current_stream()
            return
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        main, side = torch.cuda.current_stream(), self._side_stream
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        self._side_pending = True
        try:
            yield side
        finally:
            if join:
                self._join_side()

    def _join_side(self):
        if self._side_pending:
            ev = torch.cuda.Event()
            ev.record(self._side_stream)
            torch.cuda.current_stream().wait_event(ev)
            self._side_pending = False

    def _transformer_fwd(self, prefix, tag, X, B, T, training, site0):
        V, W16 = self.view, self.bview
        M, D, FF, H = B * T, 512, self.ff, self.nhead
        sm = self.model.get_submodule(prefix)
        pe = sm.pos_encoding.pe
        drop = self._drop(self.p_seq, training)
        adrop = L.attn_drop_thresh(self.p_seq if drop[0] else 0.0)
        pdrop = self.p_seq if drop[0] else 0.0
        Hcur = self.buf(tag + "H0", (M, D))
        stats = self.buf(tag + "lnstats", (2 * self.num_layers + 1, 2, M), torch.float32)
        call("pe_layernorm_fwd", None, ptr(X), ptr(pe), c_int(T), c_int(D), ptr(V[prefix + ".layer_norm.weight"]),
             ptr(V[prefix + ".layer_norm.bias"]), c_f(LN_EPS), c_ll(M), ptr(Hcur), ptr(stats[0, 0]), ptr(stats[0, 1]),
             stream())
        for l in range(self.num_layers):
            q = "%s.model.layers.%d." % (prefix, l)
            t = "%s%d" % (tag, l)
            site = site0 + 8 * l
            QKV = self.buf(t + "QKV", (M, 3 * D))
            ops.gemm(Hcur, W16[q + "self_attn.in_proj_weight"], QKV, M, 3 * D, D, bias=V[q + "self_attn.in_proj_bias"])
            CTX = self.buf(t + "CTX", (M, D))
            LSE = self.buf(t + "LSE", (B, H, T), torch.float32)
            call("pe_attn_fwd", ptr(QKV), c_int(B), c_int(T), c_int(H), c_int(64), c_u(adrop[0]), c_f(adrop[1]),
                 c_ull(self._seed(site + 0)), ptr(CTX), ptr(LSE), c_int(0), stream())
            S1 = self.buf(t + "S1", (M, D), torch.float32)
            ops.gemm(CTX, W16[q + "self_attn.out_proj.weight"], S1, M, D, D, bias=V[q + "self_attn.out_proj.bias"],
                     p_drop=pdrop, seed=self._seed(site + 1), aux=Hcur, aux_mode=L.PE_AUX_ADD)
            H1 = self.buf(t + "H1", (M, D))
            call("pe_layernorm_fwd", ptr(S1), None, None, c_int(T), c_int(D), ptr(V[q + "norm1.weight"]),
                 ptr(V[q + "norm1.bias"]), c_f(LN_EPS), c_ll(M), ptr(H1), ptr(stats[2 * l + 1, 0]),
                 ptr(stats[2 * l + 1, 1]), stream())
            U = self.buf(t + "U", (M, FF))
            G = self.buf(t + "G", (M, FF))
            # U receives d G / d(pre-activation) = gelu'(.) * dropout factor: the backward GEMM only multiplies by it
            ops.gemm(H1, W16[q + "linear1.weight"], G, M, FF, D, bias=V[q + "linear1.bias"],
                     act=L.PE_ACT_GELU_SAVE_GRAD, out2=U, p_drop=pdrop, seed=self._seed(site + 2))
            S2 = self.buf(t + "S2", (M, D), torch.float32)
            ops.gemm(G, W16[q + "linear2.weight"], S2, M, D, FF, bias=V[q + "linear2.bias"], p_drop=pdrop,
                     seed=self._seed(site + 3), aux=H1, aux_mode=L.PE_AUX_ADD)
            Hn = self.buf(t + "H2", (M, D))
            call("pe_layernorm_fwd", ptr(S2), None, None, c_int(T), c_int(D), ptr(V[q + "norm2.weight"]),
                 ptr(V[q + "norm2.bias"]), c_f(LN_EPS), c_ll(M), ptr(Hn), ptr(stats[2 * l + 2, 0]),
                 ptr(stats[2 * l + 2, 1]), stream())
            Hcur = Hn
        return Hcur


    # ------------------------------------------------------------------ BiLSTM sequence models (model.py:218-228)
    _LSTM_MODELS = ("sequence_classifier", "sequence_detector")

    @staticmethod
    def _ptrs(tensors, ctype=ctypes.c_void_p):
        return (ctype * len(tensors))(*[t.data_ptr() for t in tensors])

    def _lstm_names(self, layer):
        out = []
        for prefix in self._LSTM_MODELS:
            for sfx in ("", "_reverse"):
                out.append((prefix, "l%d%s" % (layer, sfx)))
        return out  # index = model * 2 + direction

    def _lstm_workspace(self, B, T, op=b"pe_lstm_seq_fwd"):
        fn = L.lib().pe_workspace_bytes
        fn.restype = ctypes.c_longlong
        need = int(fn(op, c_int(B), c_int(T), c_int(0)))
        return self.buf("lstm_ws_" + op.decode(), ((max(need, 4) + 3) // 4,), torch.int32)

    def _bilstm_fwd(self, Xc, Xd, B, T, training):
        V, W16 = self.view, self.bview
        M, Hh, G = B * T, 384, 1536
        drop = self._drop(self.p_lstm, training)
        # Inside the recurrent stacks every token tensor is TIME-MAJOR (row = t * B + b): a time step of the recurrence
        # then reads / writes one contiguous slab of B rows instead of B rows that lie T rows (megabytes) apart.
        X = []
        for mi, Xb in enumerate((Xc, Xd)):
            Xt = self.buf("lx_tm%d" % mi, (M, Xb.shape[1]))
            Xt.view(T, B, -1).copy_(Xb.view(B, T, -1).transpose(0, 1))
            X.append(Xt)
        Y = None
        for l in range(self.num_layers):
            In = 512 if l == 0 else 2 * Hh
            if l > 0:
                for mi in range(2):
                    if drop[0]:  # nn.LSTM inter-layer dropout (not after the last layer)
                        Xl = self.buf("lx%d_%d" % (mi, l), (M, 2 * Hh))
                        call("pe_dropout_bf16", ptr(Y[mi]), ptr(Xl), c_ll(M * 2 * Hh), c_u(drop[0]), c_f(drop[1]),
                             c_ull(self._seed(128 + 4 * l + mi)), stream())
                        X[mi] = Xl
                    else:
                        X[mi] = Y[mi]
            names = self._lstm_names(l)
            GX = [self.buf("lgx%d_%d" % (mi, l), (M, 2 * G), torch.float32) for mi in range(2)]
            for r, (prefix, sfx) in enumerate(names):
                mi, d = r >> 1, r & 1
                ops.gemm(X[mi], W16["%s.model.weight_ih_%s" % (prefix, sfx)], GX[mi][:, d * G:(d + 1) * G], M, G, In)
            self._bufs["lxin%d_0" % l], self._bufs["lxin%d_1" % l] = X[0], X[1]
            Y = [self.buf("ly%d_%d" % (mi, l), (M, 2 * Hh)) for mi in range(2)]
            C = [self.buf("lc%d_%d" % (mi, l), (M, 2 * Hh), torch.float32) for mi in range(2)]
            gx_a, c_a, y_a = self._ptrs(GX), self._ptrs(C), self._ptrs(Y)
            whh = self._ptrs([W16["%s.model.weight_hh_%s" % n] for n in names])
            bih = self._ptrs([V["%s.model.bias_ih_%s" % n] for n in names])
            bhh = self._ptrs([V["%s.model.bias_hh_%s" % n] for n in names])
            if self.lstm_persistent:  # one persistent launch per layer, W_hh resident in shared memory
                ws = self._lstm_workspace(B, T)
                call("pe_lstm_seq_fwd", c_int(B), c_int(T), c_int(Hh), gx_a, c_a, y_a, whh, bih, bhh, ptr(ws),
                     ctypes.c_size_t(ws.numel() * 4), stream())
            else:                     # one dependent launch per time step (kept as the cross-check)
                call("pe_lstm_steps_fwd", c_int(B), c_int(T), c_int(Hh), c_int(0), c_int(T), gx_a, c_a, y_a, whh, bih,
                     bhh, stream())
                L.launch_count += T - 1
        return Y[0], Y[1]

    def _bilstm_bwd(self, dHc, dHd, B, T):
        W16, g, bufs = self.bview, self.gview, self._bufs
        M, Hh, G = B * T, 384, 1536
        drop = self._drop(self.p_lstm, self._training)
        dY = [dHc, dHd]
        for l in reversed(range(self.num_layers)):
            In = 512 if l == 0 else 2 * Hh
            names = self._lstm_names(l)
            GX = [bufs["lgx%d_%d" % (mi, l)] for mi in range(2)]
            C = [bufs["lc%d_%d" % (mi, l)] for mi in range(2)]
            Y = [bufs["ly%d_%d" % (mi, l)] for mi in range(2)]
            X = [bufs["lxin%d_%d" % (l, mi)] for mi in range(2)]
            dG = [self.buf("ldg%d" % mi, (M, 2 * G)) for mi in range(2)]
            dc = [self.buf("ldc%d" % mi, (B, 2 * Hh), torch.float32) for mi in range(2)]
            gx_a, c_a, dy_a, dg_a, dc_a = (self._ptrs(GX), self._ptrs(C), self._ptrs(dY), self._ptrs(dG), self._ptrs(dc))
            whh = self._ptrs([W16["%s.model.weight_hh_%s" % n] for n in names])
            # measured on a B200 (profiles/r02_lstm_breakdown.txt): the persistent backward wins while the batch tile is
            # narrow (its epilogue is then latency-bound); at 128-column tiles the per-step launches are faster
            if self.lstm_persistent and B <= self.lstm_persistent_bwd_max_batch:
                ws = self._lstm_workspace(B, T, b"pe_lstm_seq_bwd")
                call("pe_lstm_seq_bwd", c_int(B), c_int(T), c_int(Hh), gx_a, c_a, dy_a, dg_a, whh, ptr(ws),
                     ctypes.c_size_t(ws.numel() * 4), stream())
            else:
                call("pe_lstm_steps_bwd", c_int(B), c_int(T), c_int(Hh), c_int(0), c_int(T), gx_a, c_a, dy_a, dg_a, dc_a,
                     whh, stream())
                L.launch_count += T - 1
            dX = [None, None]
            for r, (prefix, sfx) in enumerate(names):
                mi, d = r >> 1, r & 1
                dGd = dG[mi][:, d * G:(d + 1) * G]
                w_ih = "%s.model.weight_ih_%s" % (prefix, sfx)
                w_hh = "%s.model.weight_hh_%s" % (prefix, sfx)
                self._wgrad_linear(dGd, X[mi], w_ih, G, In, M)
                for bname in ("bias_ih", "bias_hh"):
                    call("pe_colsum_bf16", ptr(dGd), c_ll(M), c_int(G), c_ll(2 * G),
                         ptr(g["%s.model.%s_%s" % (prefix, bname, sfx)]), stream())
                # recurrent weights: dW_hh += sum_t dgates_t^T h_{t-1} (time-shifted token views)
                dg_off, y_off = (1, 0) if d == 0 else (0, 1)
                gw = self.mat(w_hh, G, Hh, "grad")
                for c0 in (0, 192):  # time-major: "images" = the T - 1 time steps, B token rows each, shifted by B rows
                    call("pe_wgrad_tokens",
                         ctypes.c_void_p(dG[mi].data_ptr() + 2 * (dg_off * B * 2 * G + d * G)), c_ll(2 * G),
                         c_ll(B * 2 * G),
                         ctypes.c_void_p(Y[mi].data_ptr() + 2 * (y_off * B * 2 * Hh + d * Hh + c0)), c_ll(2 * Hh),
                         c_ll(B * 2 * Hh), ctypes.c_void_p(gw.data_ptr() + 4 * c0), c_ll(Hh), c_int(T - 1), c_int(B),
                         c_int(192), c_int(G), c_int(0), stream())
                # input gradient (both directions accumulate into the same tensor)
                if d == 0:
                    dX[mi] = self.buf("ldx%d_%d" % (mi, l), (M, In))
                    ops.gemm(dGd, W16[w_ih], dX[mi], M, In, G, b_mn=True)
                else:
                    ops.gemm(dGd, W16[w_ih], dX[mi], M, In, G, b_mn=True, aux=dX[mi], aux_mode=L.PE_AUX_ADD)
            if l > 0 and drop[0]:
                for mi in range(2):
                    call("pe_dropout_bf16", ptr(dX[mi]), ptr(dX[mi]), c_ll(M * In), c_u(drop[0]), c_f(drop[1]),
                         c_ull(self._seed(128 + 4 * l + mi)), stream())
            dY = dX
        out = []
        for mi in range(2):  # back to batch-major tokens for the conv trunk
            d = self.buf("ldx_bm%d" % mi, (M, 512))
            d.view(B, T, -1).copy_(dY[mi].view(T, B, -1).transpose(0, 1))
            out.append(d)
        return out[0], out[1]

    # ------------------------------------------------------------------ heads
    def _tok(self, x):
        """[M] per-token vector, batch-major (b * T + t) <-> the order of the sequence models' outputs (time-major
        t * B + b for the BiLSTM stacks); an involution only in shape, so the direction is explicit."""
        if x is None or self.seq_type != "bilstm":
            return x
        return x.view(self._B, self._T).t().contiguous().view(-1)

    def _tok_back(self, x):
        if self.seq_type != "bilstm":
            return x
        return x.view(self._T, self._B).t().contiguous().view(-1)

    def _heads(self, f0, sil, lambda_f0, grad_scale, want_grad, gc_ext=None, gd_ext=None):
        V, g = self.view, self.gview
        M, D = self._B * self._T, self.seq_dim
        f0, sil, gc_ext, gd_ext = self._tok(f0), self._tok(sil), self._tok(gc_ext), self._tok(gd_ext)
        pred_f0 = self.buf("pred_f0", (M,), torch.float32)
        pred_sil = self.buf("pred_sil", (M,), torch.float32)
        dHc = self.buf("dHc", (M, D)) if want_grad else None
        dHd = self.buf("dHd", (M, D)) if want_grad else None
        self.loss_acc.zero_()
        gw = lambda n: ptr(g[n]) if want_grad else None
        if self.num_class > 1:  # the scalar classifier slot of the fused kernel is unused: row 0 in, scratch gradients out
            scratch = self.buf("cls_scalar_scratch", (D + 64,), torch.float32)
            gw_c = {"classifier.weight": ptr(scratch), "classifier.bias": ctypes.c_void_p(scratch.data_ptr() + 4 * D)}
            gw = lambda n, _gw=gw: (gw_c[n] if n in gw_c and want_grad else _gw(n))
        call("pe_heads_loss", ptr(self._Hc), ptr(self._Hd), c_ll(M), c_int(D), ptr(V["classifier.weight"]),
             ptr(V["classifier.bias"]), ptr(V["detector.weight"]), ptr(V["detector.bias"]), ptr(f0), ptr(sil),
             c_f(lambda_f0), c_f(grad_scale), ptr(pred_f0), ptr(pred_sil),
             ptr(self.loss_acc) if f0 is not None else None, ptr(self.loss_out) if f0 is not None else None,
             ptr(gc_ext), ptr(gd_ext), ptr(dHc), ptr(dHd), gw("classifier.weight"), gw("classifier.bias"),
             gw("detector.weight"), gw("detector.bias"), stream())
        self._pred_f0, self._pred_sil = self._tok_back(pred_f0), self._tok_back(pred_sil)   # batch-major for callers
        return dHc, dHd

    # ------------------------------------------------------------------ backward
    def _wgrad_linear(self, dY, X, name, M_out, N_in, tokens):
        """grad(weight[name]) [M_out, N_in] += dY^T X, contraction over tokens, split-K with fp32 atomics."""
        tiles = ((M_out + 127) // 128) * ((N_in + 255) // 256)
        splits = max(1, min((tokens + 63) // 64, 148 // tiles))  # one persistent round of <= 148 tiles
        ops.gemm(dY, X, self.mat(name, M_out, N_in, "grad"), M_out, N_in, tokens, a_mn=True, b_mn=True, splits=splits,
                 out_mode=L.PE_OUT_F32_ATOMIC)

    def _transformer_bwd(self, prefix, tag, X, dH, B, T, site0):
        """Backward of one encoder stack (the two stacks run on two streams, see backward_core).  Weight / bias gradients
        stay on the stack's stream: floating them on helper streams behind the data-gradient chain was measured and
        gave nothing (the phase is throughput-bound: 2.61 ms either way, profiles/r02_phase_times.txt)."""
        V, W16, g = self.view, self.bview, self.gview
        M, D, FF, H = B * T, 512, self.ff, self.nhead
        training = self._training
        drop = self._drop(self.p_seq, training)
        adrop = L.attn_drop_thresh(self.p_seq if drop[0] else 0.0)
        stats = self._bufs[tag + "lnstats"]
        sm = self.model.get_submodule(prefix)
        dS = self.buf(tag + "dS", (M, D))
        dH1 = self.buf(tag + "dH1", (M, D))
        dCTX = self.buf(tag + "dCTX", (M, D))
        delta = self.buf(tag + "delta", (B, H, T), torch.float32)
        for l in reversed(range(self.num_layers)):
            q = "%s.model.layers.%d." % (prefix, l)
            t = "%s%d" % (tag, l)
            site = site0 + 8 * l
            dSm2 = dSm1 = self.buf(tag + "dSm", (M, D))
            dU = self.buf(tag + "dU", (M, FF))
            dQKV = self.buf(tag + "dQKV", (M, 3 * D))
            Hin = self._bufs[tag + "H0"] if l == 0 else self._bufs["%s%dH2" % (tag, l - 1)]
            QKV, CTX, LSE = self._bufs[t + "QKV"], self._bufs[t + "CTX"], self._bufs[t + "LSE"]
            S1, H1, U, G, S2 = (self._bufs[t + k] for k in ("S1", "H1", "U", "G", "S2"))
            # norm2 backward (+ linear2 output dropout, linear2.bias gradient)
            call("pe_layernorm_bwd", ptr(dH), ptr(S2), None, None, c_int(T), c_int(D), ptr(V[q + "norm2.weight"]),
                 ptr(stats[2 * l + 2, 0]), ptr(stats[2 * l + 2, 1]), c_ll(M), ptr(dS), ptr(dSm2), c_u(drop[0]),
                 c_f(drop[1]), c_ull(self._seed(site + 3)), ptr(g[q + "norm2.weight"]), ptr(g[q + "norm2.bias"]),
                 ptr(g[q + "linear2.bias"]), stream())
            self._wgrad_linear(dSm2, G, q + "linear2.weight", D, FF, M)
            # linear2: dG -> through dropout and GELU' -> dU
            ops.gemm(dSm2, W16[q + "linear2.weight"], dU, M, FF, D, b_mn=True, aux=U, aux_mode=L.PE_AUX_MUL)
            self._wgrad_linear(dU, H1, q + "linear1.weight", FF, D, M)
            call("pe_colsum_bf16", ptr(dU), c_ll(M), c_int(FF), c_ll(FF), ptr(g[q + "linear1.bias"]), stream())
            # linear1: dH1 = dU W1 + dS (residual)
            ops.gemm(dU, W16[q + "linear1.weight"], dH1, M, D, FF, b_mn=True, aux=dS, aux_mode=L.PE_AUX_ADD)
            # norm1 backward (+ out_proj output dropout, out_proj.bias gradient)
            call("pe_layernorm_bwd", ptr(dH1), ptr(S1), None, None, c_int(T), c_int(D), ptr(V[q + "norm1.weight"]),
                 ptr(stats[2 * l + 1, 0]), ptr(stats[2 * l + 1, 1]), c_ll(M), ptr(dS), ptr(dSm1), c_u(drop[0]),
                 c_f(drop[1]), c_ull(self._seed(site + 1)), ptr(g[q + "norm1.weight"]), ptr(g[q + "norm1.bias"]),
                 ptr(g[q + "self_attn.out_proj.bias"]), stream())
            self._wgrad_linear(dSm1, CTX, q + "self_attn.out_proj.weight", D, D, M)
            # out_proj
            ops.gemm(dSm1, W16[q + "self_attn.out_proj.weight"], dCTX, M, D, D, b_mn=True)
            # attention
            call("pe_attn_bwd", ptr(QKV), ptr(CTX), ptr(dCTX), ptr(LSE), c_int(B), c_int(T), c_int(H), c_int(64),
                 c_u(adrop[0]), c_f(adrop[1]), c_ull(self._seed(site + 0)), ptr(dQKV), ptr(delta), c_int(0), stream())
            self._wgrad_linear(dQKV, Hin, q + "self_attn.in_proj_weight", 3 * D, D, M)
            call("pe_colsum_bf16", ptr(dQKV), c_ll(M), c_int(3 * D), c_ll(3 * D),
                 ptr(g[q + "self_attn.in_proj_bias"]), stream())
            # in_proj: dH = dQKV Wqkv + dS (residual)
            dHn = self.buf(tag + "dHin%d" % (l & 1), (M, D))
            ops.gemm(dQKV, W16[q + "self_attn.in_proj_weight"], dHn, M, D, 3 * D, b_mn=True, aux=dS,
                     aux_mode=L.PE_AUX_ADD)
            dH = dHn
        dX = self.buf(tag + "dX", (M, D))
        call("pe_layernorm_bwd", ptr(dH), None, ptr(X), ptr(sm.pos_encoding.pe), c_int(T), c_int(D),
             ptr(V[prefix + ".layer_norm.weight"]), ptr(stats[0, 0]), ptr(stats[0, 1]), c_ll(M), ptr(dX), None, c_u(0),
             c_f(1.0), c_ull(0), ptr(g[prefix + ".layer_norm.weight"]), ptr(g[prefix + ".layer_norm.bias"]), None,
             stream())
        return dX

    def backward_core(self, dHc, dHd):
        """Backward of everything below the heads; accumulates into the gradient arena."""
        B, T = self._B, self._T
        BT = B * T
        training = self._training
        W16, g, bufs = self.bview, self.gview, self._bufs
        notify = self.on_grads_ready or (self.reducer.ready if self.reducer is not None else (lambda tag: None))
        if self.seq_type == "transformer":
            with self._forked() as side:
                with torch.cuda.stream(side):
                    dSEQD = self._transformer_bwd("sequence_detector", "d", bufs["SEQD"], dHd, B, T, self._site_d)
                dSEQC = self._transformer_bwd("sequence_classifier", "c", bufs["SEQC"], dHc, B, T, self._site_c)
            notify("sequence_detector+heads")
            notify("sequence_classifier")
        else:
            dSEQC, dSEQD = self._bilstm_bwd(dHc, dHd, B, T)
            notify("sequence_detector+heads")
            notify("sequence_classifier")
        self._mark("encoder_bwd")
        tdrop = self._drop(self.p_trunk, training)
        # detector_conv
        dDD = self.buf("dDD", (BT * 2, 256))
        self._act_pool_bwd("detector_conv.1", bufs["DD"], BT, 2, 256, 1, dDD, dout_seq=dSEQD, drop=tdrop,
                           seed=self._seed(2))
        dCAT = self.buf("dCAT", (BT * 2, 640))
        ops.gemm(dDD, self.mat("detector_conv.0.weight", 256, 640), dCAT, BT * 2, 640, 256, b_mn=True)
        self._wgrad_linear(dDD, bufs["CAT"].view(BT * 2, 640), "detector_conv.0.weight", 256, 640, BT * 2)
        # pool_block
        dR = self.buf("dR3", (B, T, 10, 256))
        self._act_pool_bwd("pool_block.0", bufs["R3"], BT, 10, 256, 4, dR, dout=dCAT, ld_dout=640, c_off=384,
                           dout_seq=dSEQC, drop=tdrop, seed=self._seed(1))
        width = 10
        aux = {2: (10, 192), 1: (20, 64), 0: (40, 0)}  # aux max-pool window / concat channel offset per source R_k
        for i, (cin, cout) in reversed(list(enumerate(((64, 128), (128, 192), (192, 256)), 1))):
            r = "res_block%d" % i
            P, U, Vv = bufs["P%d" % i], bufs["U%d" % i], bufs["V%d" % i]
            gB = self.mat(r + ".conv.3.weight", cout, 9 * cout, "grad")
            gS = self.mat(r + ".conv1by1.weight", cout, cin, "grad")
            gA = self.mat(r + ".conv.0.weight", cout, 9 * cin, "grad")
            # conv B + shortcut
            dV = self.buf("dV%d" % i, (B, T, width, cout))
            # weight gradients are off the critical path (nothing reads them before the optimizer): they go to the
            # second stream and fill the SMs the data-gradient chain leaves idle
            with self._forked(join=False) as side:
                with torch.cuda.stream(side):
                    ops.conv_wgrad(dR, Vv, gB, taps=9)
                    ops.conv_wgrad(dR, P, gS, taps=1)
            ops.conv3x3(dR, self.wops[r + ".B.dgrad"], dV, **self._bn_bwd_fused(r + ".conv.1", U))
            dU = self.buf("dU%d" % i, (B, T, width, cout))
            self._act_pool_bwd(r + ".conv.1", U, BT, width, cout, 1, dU, dout=dV, ld_dout=cout, sums_ready=True)
            # conv A (+ shortcut data gradient fused as extra K columns)
            dP = self.buf("dP%d" % i, (B, T, width, cin))
            with self._forked(join=False) as side:
                with torch.cuda.stream(side):
                    ops.conv_wgrad(dU, P, gA, taps=9)
            Rin = bufs["R%d" % (i - 1)]
            ops.conv3x3(dU, self.wops[r + ".A.dgrad"], dP, x2=dR, **self._bn_bwd_fused(r + ".pre_conv.0", Rin, pooled=True))
            # pre_conv BN/LReLU/pool backward -> gradient of the block input
            dRin = self.buf("dR%d" % (i - 1), (B, T, width * 2, cin))
            k_aux, c_off = aux[i - 1]  # the auxiliary max-pool of R_{i-1} joins its gradient in the same pass
            self._act_pool_bwd(r + ".pre_conv.0", Rin, BT, width * 2, cin, 2, dRin, dout=dP, ld_dout=cin, sums_ready=True,
                               aux=(bufs["AUXIDX%d" % (i - 1)], dCAT, 640, c_off, k_aux))
            dR = dRin
            width *= 2
        # conv_block
        dZ1 = self.buf("dZ1", (B, T, 80, 64))
        with self._forked(join=False) as side:
            with torch.cuda.stream(side):
                ops.conv_wgrad(dR, bufs["Z1"], self.mat("conv_block.3.weight", 64, 576, "grad"), taps=9)
        ops.conv3x3(dR, self.wops["conv_block.3.dgrad"], dZ1, **self._bn_bwd_fused("conv_block.1", bufs["Y1"]))
        dY1 = self.buf("dY1", (B, T, 80, 64))
        self._act_pool_bwd("conv_block.1", bufs["Y1"], BT, 80, 64, 1, dY1, dout=dZ1, ld_dout=64, sums_ready=True)
        x = self._x
        call("pe_stem_conv_wgrad", ptr(x), c_ll(x.stride(0)), c_ll(x.stride(2)), c_ll(x.stride(3)), c_int(B), c_int(T),
             c_int(80), ptr(dY1), ptr(g["conv_block.0.weight"]), stream())
        self._join_side()
        notify("trunk")
        self._mark("trunk_bwd")

    # ------------------------------------------------------------------ public entry points
    def train_step(self, mel, f0, sil, lambda_f0=0.1, grad_scale=1.0):
        """mel [B,1,80,T] (reference batch layout) -> fills .grad, returns device tensor [loss, f0, sil]."""
        if self.num_class != 1:
            raise ValueError("the fused training step regresses F0 with num_class == 1 (Configs/config.yml:17; the "
                             "reference's Trainer.run cannot train num_class > 1 either: SmoothL1 of [B,T,C] against "
                             "[B,T]); use model(x) + your own loss + backward() for a classification head")
        x = mel.transpose(-1, -2)  # trainer.py:235
        if self.use_graph and ops.PROFILE is None and L.TIMING is None:
            return self._train_step_graphed(x, f0, sil, float(lambda_f0), float(grad_scale))
        self._set_salt(0)
        return self._train_step_eager(x, f0, sil, lambda_f0, grad_scale)

    def _train_step_eager(self, x, f0, sil, lambda_f0, grad_scale):
        if self._nvtx and not self._nvtx_open:
            torch.cuda.nvtx.range_push(self._PHASES[0])
            self._nvtx_open = True
        if self.reducer is not None:
            self.reducer.begin_step()
        self.zero_grad()
        self.forward_core(x, training=True)
        f0 = f0.to(self.device, torch.float32).contiguous().view(-1)
        sil = sil.to(self.device, torch.float32).contiguous().view(-1)
        dHc, dHd = self._heads(f0, sil, lambda_f0, grad_scale, want_grad=True)
        self._mark("heads_loss")
        self.backward_core(dHc, dHd)
        if self.reducer is not None:
            self.reducer.wait()  # the current stream waits for the bucket all-reduces (captured as graph edges)
        return self.loss_out

    # ------------------------------------------------------------------ CUDA-graph replay of the training step
    def _set_salt(self, salt):
        salt &= (1 << 56) - 1
        if salt != self._salt:  # this engine's own slot of the device-side salt table
            call("pe_set_step_salt", c_int(self.salt_slot), ctypes.c_ulonglong(salt), stream())
            self._salt = salt

    def _train_step_graphed(self, x, f0, sil, lambda_f0, grad_scale):
        x, B, T, _ = self._prep_input(x)
        key = (tuple(x.shape), tuple(x.stride()), lambda_f0, grad_scale, self.dropout_enabled, self.reducer is not None)
        ent = self._graphs.setdefault(key, {"warm": 0})
        if ent.get("failed") or ent["warm"] < 2:  # eager warm-up: allocates every buffer, sets kernel attributes
            ent["warm"] += 1
            self._set_salt(0)
            return self._train_step_eager(x, f0, sil, lambda_f0, grad_scale)
        if "segments" not in ent:
            self._capture(ent, x, B, T, lambda_f0, grad_scale)
            if ent.get("failed"):
                return self._train_step_eager(x, f0, sil, lambda_f0, grad_scale)
        ent["x"].copy_(x, non_blocking=True)
        ent["f0"].copy_(f0.reshape(-1), non_blocking=True)
        ent["sil"].copy_(sil.reshape(-1), non_blocking=True)
        self._x, self._B, self._T, self._training = ent["x"], B, T, True
        self._fwd_token += 1
        self.step_seed += 1
        # the graph's launch arguments carry the seeds of the capture step; the salt moves them to this step's
        self._set_salt((self.step_seed - ent["seed"]) << 8)
        self.cast_weights()  # not part of the graph: skipped when FusedAdamW.step already wrote the bf16 copy
        if self.reducer is not None:
            self.reducer.begin_step()
        if ent.get("phases"):  # profiling mode: the step was cut at the phase marks, time every segment
            cs, evs = torch.cuda.current_stream(), []
            for graph, tags in ent["segments"]:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cs)
                graph.replay()
                e1.record(cs)
                evs.append((tags[0] if tags else "tail", e0, e1))
            self.phase_events.append(evs)
            L.launch_count += ent["launches"]
            return self.loss_out
        for graph, tags in ent["segments"]:
            graph.replay()
            for tag in tags:  # gradient buckets completed by this segment: their all-reduces go out from the host
                self.reducer.ready(tag)
        if self.reducer is not None:
            self.reducer.wait()
        L.launch_count += ent["launches"]
        return self.loss_out

    def _capture(self, ent, x, B, T, lambda_f0, grad_scale):
        """Capture the step into CUDA graphs.  Single GPU: one graph.  Data parallel: the capture is cut at every
        point where the backward pass announces a finished gradient bucket, and the bucket's NCCL all-reduce is launched
        between the segments at replay time (collectives are not captured)."""
        ent["x"] = torch.empty_strided(x.shape, x.stride(), device=self.device, dtype=torch.float32)
        ent["f0"] = torch.zeros(B * T, device=self.device, dtype=torch.float32)
        ent["sil"] = torch.zeros(B * T, device=self.device, dtype=torch.float32)
        ent["x"].copy_(x)
        self._set_salt(0)
        seed_before, token_before, launches_before = self.step_seed, self._fwd_token, L.launch_count
        segments, seen, cur = [], set(), {}
        pool = torch.cuda.graph_pool_handle()

        def begin():
            cur["g"] = torch.cuda.CUDAGraph()
            cur["g"].capture_begin(pool=pool, capture_error_mode="thread_local")
            cur["launches"] = L.launch_count

        def end(tag):
            cur["g"].capture_end()
            segments.append((cur.pop("g"), [] if tag is None else [tag]))

        class _Splitter:  # stands in for the gradient reducer while capturing
            def begin_step(self_):
                pass

            def ready(self_, tag):
                if tag in seen:
                    return
                seen.add(tag)
                if segments and L.launch_count == cur["launches"]:
                    segments[-1][1].append(tag)  # nothing was launched since the last cut: same cut
                else:
                    end(tag)
                    begin()

            def wait(self_):
                pass

        real = self.reducer
        if real is None and os.environ.get("PE_PHASES") == "1":  # profiling: cut the graph at the phase marks
            ent["phases"] = True
            self.phase_events = []
            self._phase_cb = lambda tag: (end(tag), begin())
        side = torch.cuda.Stream(device=self.device)
        self.cast_weights()
        self._capturing = True
        torch.cuda.synchronize()
        side.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(side):
                self.reducer = _Splitter() if real is not None else None
                begin()
                self._train_step_eager(ent["x"], ent["f0"], ent["sil"], lambda_f0, grad_scale)
                if segments and L.launch_count == cur["launches"]:  # nothing after the last bucket: drop the empty tail
                    import warnings
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        cur.pop("g").capture_end()
                else:
                    end(None)
        except Exception as e:  # stay on the eager CUDA path (still no CPU fallback)
            self.reducer = real
            self._capturing = False
            self._phase_cb = None
            if "g" in cur:
                try:
                    cur["g"].capture_end()
                except Exception:
                    pass
            if os.environ.get("PE_CUDA_GRAPH") == "1":
                raise
            import logging
            logging.getLogger(__name__).warning("CUDA-graph capture of the training step failed (%s); running eagerly", e)
            ent["failed"] = True
            self.step_seed, self._fwd_token, L.launch_count = seed_before, token_before, launches_before
            return
        self.reducer = real
        self._capturing = False
        self._phase_cb = None
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ent["segments"] = segments
        ent["launches"] = L.launch_count - launches_before
        ent["seed"] = self.step_seed  # baked into the captured launch arguments
        self.step_seed, self._fwd_token, L.launch_count = seed_before, token_before, launches_before

    def eval_loss(self, mel, f0, sil, lambda_f0=0.1):
        if self.num_class != 1:
            raise ValueError("eval_loss needs the scalar regression head (num_class == 1); use model(x)")
        self.forward_core(mel.transpose(-1, -2), training=False)
        f0 = f0.to(self.device, torch.float32).contiguous().view(-1)
        sil = sil.to(self.device, torch.float32).contiguous().view(-1)
        self._heads(f0, sil, lambda_f0, 1.0, want_grad=False)
        return self.loss_out

    def autograd_forward(self, x):
        training = self.model.training
        self._set_salt(0)
        self.forward_core(x, training=training)
        B, T = self._B, self._T
        self._predict_only()
        if self.num_class > 1:
            logits = self._cls_logits[:, :self.num_class]
            if self.seq_type == "bilstm":  # rows are time-major there
                self._out_cls = logits.reshape(T, B, self.num_class).transpose(0, 1).contiguous()
            else:
                self._out_cls = logits.reshape(B, T, self.num_class)
        else:
            self._out_cls = self._pred_f0.view(B, T, 1)
        self._out_det = self._pred_sil.view(B, T)
        if torch.is_grad_enabled() and training:
            return _OutputGrad.apply(self, self._fwd_token, *self.params)
        return self._out_cls.clone(), self._out_det.clone()

    def _predict_only(self):
        M = self._B * self._T
        zeros = self.buf("zero_targets", (M,), torch.float32)
        zeros.zero_()
        self._heads(zeros, zeros, 0.0, 1.0, want_grad=False)
        if self.num_class > 1:
            self._classifier_fwd()

    # ------------------------------------------------------------------ classification head (num_class > 1)
    def _classifier_pad(self):
        """Zero-padded (to a multiple of 64 classes) bf16 weight / fp32 bias operands of Linear(D -> num_class)."""
        nc, D = self.num_class, self.seq_dim
        npad = _align(nc, 64)
        W = self.buf("cls_wpad", (npad, D))
        b = self.buf("cls_bpad", (npad,), torch.float32)
        if not getattr(self, "_cls_pad_zeroed", False):
            W.zero_()
            b.zero_()
            self._cls_pad_zeroed = True
        W[:nc].copy_(self.bview["classifier.weight"])
        b[:nc].copy_(self.view["classifier.bias"])
        return W, b, npad

    def _classifier_fwd(self):
        """model.py:96-98 with num_class > 1: logits [M, num_class] = Hc W^T + b on the tile engine."""
        M, D = self._B * self._T, self.seq_dim
        W, b, npad = self._classifier_pad()
        out = self.buf("cls_logits", (M, npad), torch.float32)
        ops.gemm(self._Hc, W, out, M, npad, D, bias=b)
        self._cls_logits = out

    def _classifier_bwd(self, dcls, dHc):
        """dHc = dcls W (overwrites the zero the scalar path left there); dW += dcls^T Hc; db += column sums."""
        M, D, nc = self._B * self._T, self.seq_dim, self.num_class
        W, _, npad = self._classifier_pad()
        g = self.buf("cls_gpad", (M, npad))
        g.zero_()
        if self.seq_type == "bilstm":
            dcls = dcls.reshape(self._B, self._T, nc).transpose(0, 1)
        g[:, :nc].copy_(dcls.reshape(M, nc))
        ops.gemm(g, W, dHc, M, D, npad, b_mn=True)
        dW = self.buf("cls_dw", (npad, D), torch.float32)
        db = self.buf("cls_db", (npad,), torch.float32)
        dW.zero_()
        db.zero_()
        tiles = ((npad + 127) // 128) * ((D + 255) // 256)
        splits = max(1, min((M + 63) // 64, 148 // tiles))
        ops.gemm(g, self._Hc, dW, npad, D, M, a_mn=True, b_mn=True, splits=splits, out_mode=L.PE_OUT_F32_ATOMIC)
        ops.colsum(g, db)
        self.gview["classifier.weight"].add_(dW[:nc])
        self.gview["classifier.bias"].add_(db[:nc])

    def backward_from_output_grads(self, dcls, ddet):
        M = self._B * self._T
        self._set_salt(0)
        gd = ddet.to(torch.float32).contiguous().view(M)
        if any(p.grad is None for p in self.params):
            self.flat_grad.zero_()  # grads were dropped by zero_grad(set_to_none=True): start from zero
        self.attach_grads()
        if self.num_class > 1:
            zero = self.buf("zero_targets", (M,), torch.float32)
            zero.zero_()
            dHc, dHd = self._heads(None, None, 0.0, 1.0, want_grad=True, gc_ext=zero, gd_ext=gd)
            self._classifier_bwd(dcls.to(torch.float32), dHc)
        else:
            gc = dcls.to(torch.float32).contiguous().view(M)
            dHc, dHd = self._heads(None, None, 0.0, 1.0, want_grad=True, gc_ext=gc, gd_ext=gd)
        self.backward_core(dHc, dHd)
