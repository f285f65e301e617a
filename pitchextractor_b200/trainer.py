"""``Trainer`` with the reference's constructor and methods (reference trainer.py:32-291), driving the fused CUDA step.

Differences that follow from the B200 design (documented, not hidden):
  * mixed precision is always bf16-operands / fp32-accumulate inside the kernels; no GradScaler is needed (the
    reference autocasts to fp16 and scales, trainer.py:63-102).  ``use_mixed_precision`` is accepted and ignored.
  * ``gradient_checkpointing`` is accepted and ignored: activations are kept in HBM, so the reference's double
    BatchNorm running-stat update under checkpointing does not happen here.
  * a batch may carry waveforms instead of mels (``Collater(return_wave=True)``): the log-mel then runs on the GPU.
  * under torch.distributed (one process per GPU) gradients are all-reduced in overlapped buckets.
"""
import logging
import collections
import os
from collections import defaultdict

import numpy as np
import torch
import torch.distributed as dist
from tqdm import tqdm

from .mel import LogMel
from .parallel import GradReducer, broadcast_parameters, bucket_ranges

logger = logging.getLogger(__name__)
logger.setLevel(logging.DEBUG)


class Trainer(object):
    def __init__(self, model=None, criterion=None, optimizer=None, scheduler=None, config={}, loss_config={},
                 device=torch.device("cpu"), logger=logger, train_dataloader=None, val_dataloader=None,
                 initial_steps=0, initial_epochs=0, use_mixed_precision=False, gradient_checkpointing=False,
                 checkpoint_use_reentrant=None):
        self.steps, self.epochs = initial_steps, initial_epochs
        self.model, self.criterion, self.optimizer, self.scheduler = model, criterion, optimizer, scheduler
        self.train_dataloader, self.val_dataloader = train_dataloader, val_dataloader
        self.config, self.loss_config = config, loss_config
        self.device = torch.device(device)
        self.logger = logger
        self.finish_train = False
        if self.device.type != "cuda":
            raise RuntimeError("pitchextractor_b200.Trainer runs on a CUDA (sm_100a) device; there is no CPU fallback")
        # The fused step computes exactly the criteria train.py:104-106 builds: {'l1': SmoothL1Loss(), 'ce':
        # BCEWithLogitsLoss()} with default arguments.  Any other criterion dict still trains on the CUDA engine, through
        # JDCNet.forward + autograd (losses evaluated by the caller's modules on the GPU); anything else is an error.
        self._fused_losses = self._criterion_is_reference_default(criterion)
        self.use_amp = True  # bf16 tensor-core operands, fp32 accumulation / master weights
        self.gradient_checkpointing = False
        if gradient_checkpointing:
            self.logger.info("gradient_checkpointing ignored: activations stay resident in HBM")
        self._logmel = None
        self._copy_stream = None
        self._loss_ring = []   # pinned [3] buffers of run_pipelined
        self._reducer = None
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.sync_every_step = True  # reference semantics: run() returns python floats (3 .item() syncs there, 1 here)

    @staticmethod
    def _criterion_is_reference_default(criterion):
        """True when the fused heads + losses kernel computes exactly this criterion (None = the reference default)."""
        if criterion is None:
            return True
        if not isinstance(criterion, dict) or set(criterion) != {"l1", "ce"}:
            raise ValueError("criterion must be None or a dict {'l1': <F0 loss>, 'ce': <silence loss>} as built by the "
                             "reference (train.py:104-106); got %r" % (criterion,))
        l1, ce = criterion["l1"], criterion["ce"]
        if not (callable(l1) and callable(ce)):
            raise ValueError("criterion['l1'] and criterion['ce'] must be callable loss modules")
        return (type(l1) is torch.nn.SmoothL1Loss and l1.reduction == "mean" and float(l1.beta) == 1.0
                and type(ce) is torch.nn.BCEWithLogitsLoss and ce.reduction == "mean" and ce.weight is None
                and ce.pos_weight is None)

    # ------------------------------------------------------------------ checkpoints (trainer.py:138-195)
    def save_checkpoint(self, checkpoint_path):
        state = {"optimizer": self.optimizer.state_dict(), "scheduler": self.scheduler.state_dict(),
                 "steps": self.steps, "epochs": self.epochs,
                 "model": {k: v.detach().clone().contiguous() for k, v in self.model.state_dict().items()}}
        d = os.path.dirname(checkpoint_path)
        if d and not os.path.exists(d):
            os.makedirs(d)
        torch.save(state, checkpoint_path)

    def load_checkpoint(self, checkpoint_path, load_only_params=False):
        state = torch.load(checkpoint_path, map_location="cpu")
        self._load(state["model"], self.model)
        if not load_only_params:
            self.steps, self.epochs = state["steps"], state["epochs"]
            self.optimizer.load_state_dict(state["optimizer"])
            state["scheduler"].update(**self.config.get("scheduler_params", {}))
            self.scheduler.load_state_dict(state["scheduler"])

    def _load(self, states, model, force_load=True):
        """Shape-tolerant copy: missing keys are skipped, mismatching shapes copy the overlapping corner."""
        own = model.state_dict()
        for key, val in states.items():
            if key not in own:
                continue
            val = val.data if isinstance(val, torch.nn.Parameter) else val
            dst = own[key]
            if val.shape == dst.shape:
                dst.copy_(val)
                continue
            self.logger.info("%s does not have same shape" % key)
            if not force_load or val.dim() != dst.dim():
                continue
            corner = tuple(slice(0, min(a, b)) for a, b in zip(val.shape, dst.shape))
            dst[corner].copy_(val[corner])
        if getattr(model, "_engine", None) is not None:
            model._engine.invalidate_bf16()

    @staticmethod
    def get_gradient_norm(model):
        return float(np.sqrt(sum(p.grad.data.norm(2).item() ** 2 for p in model.parameters())))

    def _get_lr(self):
        return self.optimizer.param_groups[0]["lr"]

    # ------------------------------------------------------------------ one step (trainer.py:219-252)
    def _mel_from_batch(self, batch):
        """-> (mel-like model input, f0, sil).  batch = (mels [B,1,80,192], f0, sil) or (waves, f0, sil, crops[, lengths])."""
        if len(batch) >= 4:
            waves, f0, sil, crops = batch[:4]
            lengths = batch[4].to(self.device, non_blocking=True) if len(batch) > 4 else None
            if self._logmel is None:
                self._logmel = LogMel(self.device)
            x = self._logmel(waves.to(self.device, non_blocking=True), crop=crops.to(self.device, non_blocking=True),
                             T_out=192, layout="btm", lengths=lengths)  # [B, 192, 80] == the transposed model input
            return x[:, None].transpose(-1, -2), f0, sil  # present it as the reference's [B,1,80,192] view
        x, f0, sil = batch
        return x.to(self.device, non_blocking=True), f0, sil

    # ------------------------------------------------------------------ input pipeline
    class _DeviceBatch(tuple):
        """A batch whose tensors were copied to the device on the copy stream; `ready` orders consumers behind it."""
        ready = None

    def prefetched(self, batches):
        """Iterate host batches one ahead: the pinned-host -> device copy of batch i+1 is issued on a copy stream before
        batch i is handed out, so it overlaps batch i's training step (the reference copies synchronously,
        trainer.py:221-224)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)

        def to_device(batch):
            if batch is None:
                return None
            with torch.cuda.stream(self._copy_stream):
                dev = Trainer._DeviceBatch(t.to(self.device, non_blocking=True) if torch.is_tensor(t) else t for t in batch)
                dev.ready = torch.cuda.Event()
                dev.ready.record(self._copy_stream)
            return dev

        it = iter(batches)
        nxt = to_device(next(it, None))
        while nxt is not None:
            cur, nxt = nxt, to_device(next(it, None))
            yield cur

    def _await_batch(self, batch):
        ready = getattr(batch, "ready", None)
        if ready is not None:
            cur = torch.cuda.current_stream()
            cur.wait_event(ready)
            for t in batch:
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(cur)  # allocated on the copy stream, consumed here

    def _ensure_parallel(self):
        if self.world > 1 and self._reducer is None:
            eng = self.model.engine
            broadcast_parameters(eng.flat)
            eng.invalidate_bf16()
            for b in self.model.buffers():
                if b.dtype.is_floating_point:
                    dist.broadcast(b, src=0)
            offs = [eng.offset[n] for n in eng.names]
            self._reducer = GradReducer(eng.flat_grad, bucket_ranges(eng.names, offs, eng.total))
            eng.reducer = self._reducer  # the engine brackets each step with begin_step() / ready(tag) / wait()
            if hasattr(self.optimizer, "grad_scale"):
                self.optimizer.grad_scale = 1.0 / self.world

    def run_async(self, batch):
        """One optimisation step; returns the device tensor [loss, f0, sil] without synchronising.  The tensor is the
        engine's persistent loss buffer: the next step overwrites it, so ``.clone()`` (or copy to the host, as ``run`` and
        ``run_pipelined`` do) anything that must outlive the step."""
        self._ensure_parallel()
        self._await_batch(batch)
        x, f0, sil = self._mel_from_batch(batch)
        f0 = f0.to(self.device, non_blocking=True)
        sil = sil.to(self.device, non_blocking=True)
        if self._fused_losses:
            losses = self.model.train_step_loss(x, f0, sil, self.loss_config["lambda_f0"])
        else:
            losses = self._autograd_step(x, f0, sil)
        if self._reducer is not None and not hasattr(self.optimizer, "grad_scale"):
            self.model.engine.flat_grad.mul_(1.0 / self.world)
        self.optimizer.step()
        self.scheduler.step()
        self.steps += 1
        return losses

    def _autograd_step(self, x, f0, sil):
        """trainer.py:226-246 with a caller-supplied criterion: engine forward, torch losses, engine backward."""
        eng = self.model.engine
        if self._reducer is not None:
            self._reducer.begin_step()
        eng.zero_grad()
        f0_pred, sil_pred = self.model(x.transpose(-1, -2))
        loss_f0 = self.loss_config["lambda_f0"] * self.criterion["l1"](f0_pred.squeeze(), f0)
        loss_sil = self.criterion["ce"](sil_pred, sil)
        loss = loss_f0 + loss_sil
        loss.backward()
        if self._reducer is not None:
            self._reducer.wait()
        return torch.stack([loss.detach(), loss_f0.detach(), loss_sil.detach()]).float()

    def run(self, batch):
        losses = self.run_async(batch).tolist()  # one device->host read of 3 floats
        return {"loss": losses[0], "f0": losses[1], "sil": losses[2]}

    def run_pipelined(self, batches, depth=None):
        """``run`` over an iterable of host batches, yielding the same ``{'loss','f0','sil'}`` floats per batch, but up to
        ``depth`` steps behind the GPU (default 4, ``PE_PIPELINE_DEPTH``): steps i+1 .. i+depth (host->device copies,
        log-mel, graph replay, optimizer) are enqueued before the host waits for the three loss floats of step i, which
        travel through a small ring of pinned buffers.  The GPU does not idle while Python prepares the next step, and a
        host hiccup shorter than ``depth`` steps (a descheduled thread, a garbage collection) is absorbed by the queue;
        every batch is still copied in and every loss read back, in order."""
        if depth is None:
            depth = int(os.environ.get("PE_PIPELINE_DEPTH", "4"))
        depth = max(1, depth)
        if len(self._loss_ring) < depth + 2:  # pinned once: a pinned allocation inside the loop can stall the GPU queue
            self._loss_ring += [torch.empty(3, dtype=torch.float32).pin_memory()
                                for _ in range(depth + 2 - len(self._loss_ring))]
        ring = self._loss_ring[:depth + 2]
        pending, k = collections.deque(), 0

        def resolve(item):
            host, ev = item
            ev.synchronize()
            v = host.tolist()
            return {"loss": v[0], "f0": v[1], "sil": v[2]}

        for batch in self.prefetched(batches):
            dev_losses = self.run_async(batch)
            host = ring[k % len(ring)]
            k += 1
            host.copy_(dev_losses, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((host, ev))
            if len(pending) > depth:
                yield resolve(pending.popleft())
        while pending:
            yield resolve(pending.popleft())

    def _train_epoch(self):
        self.epochs += 1
        train_losses = defaultdict(list)
        self.model.train()
        total = len(self.train_dataloader) if hasattr(self.train_dataloader, "__len__") else None
        for _, losses in enumerate(tqdm(self.run_pipelined(self.train_dataloader), desc="[train]", total=total), 1):
            for key, value in losses.items():
                train_losses["train/%s" % key].append(value)
        train_losses = {key: np.mean(value) for key, value in train_losses.items()}
        train_losses["train/learning_rate"] = self._get_lr()
        return train_losses

    @torch.no_grad()
    def _eval_epoch(self):
        self.model.eval()
        eval_losses = defaultdict(list)
        for _, batch in enumerate(tqdm(self.val_dataloader, desc="[eval]"), 1):
            x, f0, sil = self._mel_from_batch(batch)
            if self._fused_losses:
                out = self.model.engine.eval_loss(x, f0, sil, self.loss_config["lambda_f0"]).tolist()
            else:
                f0_d, sil_d = f0.to(self.device), sil.to(self.device)
                f0_pred, sil_pred = self.model(x.transpose(-1, -2))
                lf = self.loss_config["lambda_f0"] * self.criterion["l1"](f0_pred.squeeze(), f0_d)
                ls = self.criterion["ce"](sil_pred, sil_d)
                out = [(lf + ls).item(), lf.item(), ls.item()]
            eval_losses["eval/loss"].append(out[0])
            eval_losses["eval/f0"].append(out[1])
            eval_losses["eval/sil"].append(out[2])
        return {key: np.mean(value) for key, value in eval_losses.items()}
