"""AdamW + OneCycleLR with the reference's ``build_optimizer`` contract (reference optimizers.py:50-76), the update
itself being one fused sm_100a kernel over JDCNet's flat parameter arena."""
import ctypes
import math

import torch
from torch.optim import Optimizer

from ._lib import call, ptr, stream

_ENGINES = {}  # id(first parameter) -> Engine; filled by Engine._pack


def register_engine(engine):
    for p in engine.params:
        _ENGINES[id(p)] = engine


class FusedAdamW(Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction) as a single pass over the flat fp32
    parameter / gradient arenas of a ``pitchextractor_b200.JDCNet``.  ``param_groups`` / ``state_dict`` keep torch's
    layout so ``OneCycleLR`` (lr and beta1 cycling) and reference checkpoints work unchanged."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=5e-4):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._engine = None
        self._steps = 0
        self.grad_scale = 1.0  # multiplied into the gradients (1/world_size for data parallel)

    def _bind(self):
        ps = [p for g in self.param_groups for p in g["params"]]
        eng = _ENGINES.get(id(ps[0]))
        if eng is None or len(ps) != len(eng.params) or any(a is not b for a, b in zip(ps, eng.params)):
            raise RuntimeError("FusedAdamW optimises exactly the parameters of one pitchextractor_b200.JDCNet that "
                               "lives on a CUDA device (call model.to('cuda') and run one step or model.engine first)")
        if len(self.param_groups) != 1:
            raise RuntimeError("FusedAdamW supports a single parameter group (as the reference uses)")
        self._engine = eng
        self.exp_avg = torch.zeros_like(eng.flat)
        self.exp_avg_sq = torch.zeros_like(eng.flat)
        for name, p in zip(eng.names, eng.params):
            off, n = eng.offset[name], p.numel()
            old = self.state.get(p, {})
            mk = (lambda buf: buf[off:off + n].view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)) \
                if p.dim() == 4 else (lambda buf: buf[off:off + n].view(p.shape))
            m, v = mk(self.exp_avg), mk(self.exp_avg_sq)
            if "exp_avg" in old:  # state restored by load_state_dict before binding
                m.copy_(old["exp_avg"])
                v.copy_(old["exp_avg_sq"])
                self._steps = int(old["step"])
            self.state[p] = {"step": torch.tensor(float(self._steps)), "exp_avg": m, "exp_avg_sq": v}

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("closures are not supported")
        if self._engine is None:
            self._bind()
        eng, g = self._engine, self.param_groups[0]
        self._steps += 1
        b1, b2 = g["betas"]
        call("pe_adamw", ptr(eng.flat), ptr(eng.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             ctypes.c_longlong(eng.total), ctypes.c_float(g["lr"]), ctypes.c_float(b1), ctypes.c_float(b2),
             ctypes.c_float(g["eps"]), ctypes.c_float(g["weight_decay"]), ctypes.c_longlong(self._steps),
             ctypes.c_float(self.grad_scale), ptr(eng.flat_bf16), stream())
        eng.bf16_fresh = True  # the same pass wrote the bf16 tensor-core copy of the updated weights
        for st in self.state.values():
            st["step"].fill_(float(self._steps))

    def zero_grad(self, set_to_none=True):
        if self._engine is None:
            return super().zero_grad(set_to_none)
        self._engine.zero_grad()  # one memset of the flat gradient arena; the .grad views stay attached

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._engine = None  # re-bind lazily and pull the restored moments into the flat arenas


def build_optimizer(parameters):
    """Same call contract as the reference: ``{'params', 'optimizer_params', 'scheduler_params'}`` ->
    ``(optimizer, scheduler)`` (optimizers.py:50-76)."""
    opt_params = parameters.get("optimizer_params", {}) or {}
    sch = parameters.get("scheduler_params", {}) or {}
    optimizer = FusedAdamW(parameters["params"], lr=opt_params.get("lr", 1e-4),
                           weight_decay=opt_params.get("weight_decay", 5e-4), betas=(0.9, 0.98), eps=1e-9)
    scheduler = torch.optim.lr_scheduler.OneCycleLR(
        optimizer, max_lr=sch.get("max_lr", 5e-4), epochs=sch.get("epochs", 200),
        steps_per_epoch=sch.get("steps_per_epoch", 1000), pct_start=sch.get("pct_start", 0.0), final_div_factor=5)
    return optimizer, scheduler
