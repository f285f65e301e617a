"""Drive the UNMODIFIED reference staged under ``baseline/_ref`` (see install_ref.py) through its own public API:
``MelDataset._build_training_example`` (torchaudio MelSpectrogram per sample, meldataset.py:629-677) -> ``Collater``
(meldataset.py:804-826) -> ``Trainer.run`` (trainer.py:219-252) with ``JDCNet`` (model.py), AdamW + OneCycleLR from
``build_optimizer`` (optimizers.py:54-76) and the criteria of train.py:104-106.

Only import-time stubs are injected (no code edits): ``matplotlib.pyplot`` (imported, unused, trainer.py:16),
``soundfile`` (meldataset.py:10-11; never called, the segments are synthetic arrays) and an empty ``pyworld`` so that
``PyWorldBackend`` constructs (f0_backends.py:112-122; never called, F0 labels are supplied).
None of this repo's kernels, models or engine are on this path; the synthetic *inputs* come from
``pitchextractor_b200.synthetic`` (numpy only), the same generator the CUDA arm uses.
"""
import importlib
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "trainer.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_NS = None


def load():
    global _NS
    if _NS is not None:
        return _NS
    if not available():
        raise RuntimeError("reference not staged under baseline/_ref (run baseline/install_ref.py in the build container)")
    mpl, plt = _stub("matplotlib"), _stub("matplotlib.pyplot")
    mpl.pyplot = plt

    class LibsndfileError(Exception):
        pass

    _stub("soundfile", LibsndfileError=LibsndfileError, info=None, read=None, SoundFile=None)
    _stub("pyworld")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    ns = types.SimpleNamespace()
    for name in ("model", "optimizers", "f0_backends", "meldataset", "trainer"):
        mod = importlib.import_module(name)
        assert os.path.dirname(os.path.abspath(mod.__file__)) == REF, (name, mod.__file__)
        setattr(ns, name, mod)
    _NS = ns
    return ns


MEL_PARAMS = {"sample_rate": 24000, "win_length": 1024, "n_fft": 1024, "n_mels": 80, "hop_length": 300}


class ReferenceRunner:
    """One reference training setup (model + optimizer + scheduler + Trainer + dataset/collater) on ``device``."""

    def __init__(self, model_cfg, device="cpu", mixed_precision=False, gradient_checkpointing=False, seed=0,
                 state_dict=None):
        import logging
        ns = load()
        logging.getLogger("trainer").setLevel(logging.WARNING)
        logging.getLogger("meldataset").setLevel(logging.WARNING)
        torch.manual_seed(seed)
        self.device = torch.device(device)
        self.model = ns.model.JDCNet(num_class=1, sequence_model_config=dict(model_cfg))
        if state_dict is not None:
            self.model.load_state_dict(state_dict)
        self.model.to(self.device)
        sp = {"max_lr": 3e-4, "pct_start": 0.0, "epochs": 100, "steps_per_epoch": 1000}
        self.optimizer, self.scheduler = ns.optimizers.build_optimizer(
            {"params": self.model.parameters(), "optimizer_params": {}, "scheduler_params": sp})
        criterion = {"l1": torch.nn.SmoothL1Loss(), "ce": torch.nn.BCEWithLogitsLoss()}  # train.py:104-106
        self.trainer = ns.trainer.Trainer(
            model=self.model, criterion=criterion, optimizer=self.optimizer, scheduler=self.scheduler,
            device=self.device, loss_config={"lambda_f0": 0.1}, use_mixed_precision=mixed_precision,
            gradient_checkpointing=gradient_checkpointing, checkpoint_use_reentrant=False)
        self.model.train()
        self.dataset = ns.meldataset.MelDataset([], mel_params=dict(MEL_PARAMS), verbose=False)
        self.collate = ns.meldataset.Collater()

    def make_batch(self, waves, f0s):
        """Reference data path: one torchaudio mel per sample on the CPU, random 192-frame crop, collate."""
        items = [self.dataset._build_training_example(np.asarray(w), 24000, np.asarray(f, dtype=np.float64),
                                                      cache_key=None, allow_cache=False)
                 for w, f in zip(waves, f0s)]
        return self.collate(items)

    def step(self, waves, f0s):
        return self.trainer.run(self.make_batch(waves, f0s))

    def step_collated(self, batch):
        return self.trainer.run(batch)


def time_cpu(model_cfg, batch, steps, warmup, budget_s=200.0, seed_base=4321):
    """segments/s of the reference step (log-mel per sample + Trainer.run) on the host cores.  The per-step sample is
    the full batch unless (warmup + steps) steps of it would exceed budget_s, in which case it is halved until it fits."""
    from pitchextractor_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run = ReferenceRunner(model_cfg, "cpu")
    n = batch
    waves, f0 = synthetic.make_batch(batch, seed=seed_base)
    t0 = time.perf_counter()
    last = run.step(waves[:min(n, 8)], f0[:min(n, 8)])  # probe (also the first warm-up)
    probe = (time.perf_counter() - t0) / min(n, 8)
    while n > 8 and probe * n * (warmup + steps) > budget_s:
        n //= 2
    for _ in range(max(0, warmup - 1)):
        run.step(waves[:n], f0[:n])
    t0 = time.perf_counter()
    for s in range(steps):
        last = run.step(waves[:n], f0[:n])
    dt = time.perf_counter() - t0
    return {"segments_per_s": n * steps / dt, "s_per_step": dt / steps, "segments_per_step": n, "cores": cores,
            "loss_last": last}


def time_gpu(model_cfg, batch, steps, warmup, variants=None):
    """The reference on the same GPU through torch (SURVEY 8d "the real bar"): unmodified JDCNet + Trainer.run with
    cudnn.benchmark (train.py:28); mels are collated on the host beforehand (the reference computes them in DataLoader
    workers) and every step includes the reference's own host->device copy and three .item() read-backs."""
    from pitchextractor_b200 import synthetic
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda", torch.cuda.current_device())
    variants = variants or [("fp16_amp+checkpointing (config.yml default)", True, True, None),
                            ("fp16_amp", True, False, None),
                            ("bf16_autocast", True, False, torch.bfloat16),
                            ("fp32", False, False, None)]
    waves, f0 = synthetic.make_batch(batch, seed=4321)
    out = {}
    batches = None
    for name, amp, ckpt, dtype in variants:
        old = torch.get_autocast_dtype("cuda")
        try:
            if dtype is not None:
                torch.set_autocast_dtype("cuda", dtype)  # the reference's autocast takes torch's default dtype
            run = ReferenceRunner(model_cfg, dev, mixed_precision=amp, gradient_checkpointing=ckpt)
            if batches is None:
                b = run.make_batch(waves, f0)
                batches = [tuple(t.clone().pin_memory() for t in b) for _ in range(2)]
            for i in range(warmup):
                run.step_collated(batches[i % 2])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                last = run.step_collated(batches[i % 2])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"segments_per_s": batch / (ms / 1e3), "ms_per_step": ms, "loss_last": last["loss"]}
        except Exception as e:  # report, do not hide
            out[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        finally:
            torch.set_autocast_dtype("cuda", old)
            run = None
            torch.cuda.empty_cache()
    return out
