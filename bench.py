#!/usr/bin/env python
"""Benchmark of the hot path: JDCNet training segments/s (log-mel + forward + losses + backward + AdamW step).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU step (oracle port), rank 0 only

Workload (BASELINE.json configs[1], north_star): JDCNet with the Transformer sequence model (4 layers, d=512, 8 heads,
FFN 1536 -- Configs/config.yml:16-24), batch 64 segments per GPU, each a synthetic 58 624-sample 24 kHz segment
(pitchextractor_b200/synthetic.py) -> 192 log-mel frames, bf16 tensor-core operands with fp32 accumulation.
One JSON line is printed by rank 0 (contract: task statement).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SEG = 58624
FRAMES = 192
MODEL_CFG = dict(model_type="transformer", num_layers=4, dropout=0.1, nhead=8, dim_feedforward=1536, max_len=2048)
# algorithmic training flops per segment, 3 x forward: conv trunk 13.39 GFLOP + transformers 8.66
# (or BiLSTMs 10.27)
FLOPS_PER_SEGMENT = {"transformer": 3.0 * (13.39e9 + 8.66e9), "bilstm": 3.0 * (13.39e9 + 10.27e9)}
WORKLOAD = {"transformer": "JDCNet Transformer (4L d512 h8 ffn1536) train step incl. log-mel, batch %d/GPU, 24 kHz "
                           "58624-sample segments -> 192 frames (BASELINE configs[1])",
            "bilstm": "JDCNet BiLSTM (4L h384 bidirectional) train step incl. log-mel, batch %d/GPU, 24 kHz "
                      "58624-sample segments -> 192 frames (BASELINE configs[3] at 512/GPU, configs[0] at 16)"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_traffic(model, batch):
    """Mean DRAM bytes per tile-engine launch from the ncu capture of this (model, batch), if one is committed."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get("%s_b%d" % (model, batch), {}).get("bytes_per_launch")
    except Exception:
        return None


def make_pool(batch, n_batches, rank):
    """Host (pinned) pool of synthetic batches: waves, f0 [B,192], sil [B,192], crops."""
    from pitchextractor_b200 import synthetic
    from pitchextractor_b200.meldataset import align_length
    pool = []
    for i in range(n_batches):
        waves, f0_full = synthetic.make_batch(batch, seed=1234 + 1000 * rank + i)
        rng = np.random.default_rng([99, rank, i])
        crops = rng.integers(0, f0_full.shape[1] - FRAMES, size=batch).astype(np.int32)
        f0 = np.stack([align_length(f0_full[b], f0_full.shape[1])[c:c + FRAMES] for b, c in enumerate(crops)])
        sil = (f0 == 0).astype(np.float32)
        pool.append(tuple(torch.from_numpy(a).pin_memory() for a in (waves, f0.astype(np.float32), sil, crops)))
    return pool


def run_ours(args):
    from pitchextractor_b200 import JDCNet, Trainer, build_optimizer, ops, _lib
    from pitchextractor_b200.parallel import init_from_env
    import torch.distributed as dist
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line
    rank, world, local_rank = init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.batch
    torch.manual_seed(0)
    cfg = dict(MODEL_CFG, model_type=args.model)
    model = JDCNet(num_class=1, sequence_model_config=cfg).to(dev)
    opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {},
                                  "scheduler_params": {"max_lr": 3e-4, "pct_start": 0.0, "epochs": 100,
                                                       "steps_per_epoch": 1000}})
    trainer = Trainer(model=model, criterion=None, optimizer=opt, scheduler=sched, config={},
                      loss_config={"lambda_f0": 0.1}, device=dev)
    model.train()
    pool = make_pool(B, args.pool, rank)
    dpool = [tuple(t.to(dev) for t in b) for b in pool]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    resident = lambda i: trainer.run_async(dpool[i % len(dpool)])
    host = lambda i: trainer.run(pool[i % len(pool)])
    for i in range(args.warmup):
        resident(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = _lib.launch_count
    ms = timed(resident, args.steps)
    launches = _lib.launch_count - n0
    clocks = sampler.stop()
    last = trainer.run(pool[0])
    # end to end: host (pinned) buffers in, python floats out, every step
    # (Trainer.run_pipelined is the package's epoch loop, what Trainer._train_epoch iterates: pinned host batches in, the
    # three loss floats of every step out; the copy of batch i+1 and the enqueue of step i+1 overlap step i.  Every
    # host->device copy and every loss read-back lies inside the timed region.)
    def host_loop(steps):
        n = 0
        for out in trainer.run_pipelined(pool[i % len(pool)] for i in range(steps)):
            n += 1
        assert n == steps and out["loss"] == out["loss"]
    host_loop(args.warmup)  # the W warm-up steps of this path (pinned loss ring, copy stream, allocator steady state)
    gc.collect()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    host_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = ms_e2e.item()
    h2d = sum(t.numel() * t.element_size() for t in pool[0])
    # dominant kernel: the tcgen05 tile engine (implicit-GEMM conv / GEMM / weight-gradient launches)
    # (eager, one stream: with the two-stream / CUDA-graph step the per-launch events of concurrent kernels overlap)
    ops.PROFILE = []
    two_streams = os.environ.get("PE_TWO_STREAMS")
    os.environ["PE_TWO_STREAMS"] = "0"
    barrier()
    for i in range(min(3, args.steps)):
        resident(i)
    torch.cuda.synchronize()
    if two_streams is None:
        del os.environ["PE_TWO_STREAMS"]
    else:
        os.environ["PE_TWO_STREAMS"] = two_streams
    prof, ops.PROFILE = ops.PROFILE, None
    tc_ms = sum(a.elapsed_time(b) for _, a, b, _ in prof)
    tc_flops = sum(f for _, _, _, f in prof)
    n_prof_steps = min(3, args.steps)
    # log-mel alone (second headline metric): (a) as it runs inside the step -- the step's batch, 192 of 196 frames per
    # segment; (b) the largest segment-shaped cell of the BASELINE configs[4] sweep (batch 1024), where the launch
    # and pipeline-fill overheads of a 40-microsecond kernel no longer dominate
    from pitchextractor_b200.mel import LogMel
    lm = LogMel(dev)

    def time_logmel(w, crop, n):
        for _ in range(3):
            lm(w, crop=crop, T_out=FRAMES, layout="btm")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            lm(w, crop=crop, T_out=FRAMES, layout="btm")
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    lm_ms = time_logmel(dpool[0][0], dpool[0][3], 20)
    LM_BIG = 1024
    gen = torch.Generator(device=dev).manual_seed(7)
    wbig = torch.randn(LM_BIG, SEG, device=dev, generator=gen) * 0.1     # 240 MB: larger than the 126 MB L2
    lm_big_ms = time_logmel(wbig, None, 10)
    del wbig
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    pk = peaks()
    seg_s = world * B * args.steps / (ms / 1e3)
    line = {
        "metric": "train_segments_per_s", "value": seg_s, "unit": "segments/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD[args.model] % B,
                   "global_batch": world * B, "parallelism": "dp%d" % world,
                   "l2": "working set (>1 GB activations/step) exceeds the 126 MB L2; inputs cycle over %d batches" % len(pool)},
        "e2e": {"value": world * B * args.steps / (ms_e2e / 1e3), "unit": "segments/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 12},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "tc_tile_kernel (tcgen05 implicit-GEMM conv / GEMM / wgrad)",
                     "achieved": tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else None, "peak": pk["tf_sustained"],
                     "unit": "TFLOP/s",
                     "frac": (tc_flops / (tc_ms / 1e3) / 1e12 / pk["tf_sustained"]) if tc_ms > 0 else None,
                     # dram__bytes_read + write per launch from the committed ncu --set full capture of this
                     # configuration (profiles/traffic.json names the capture); null when there is none
                     "traffic": measured_traffic(args.model, args.batch),
                     "peak_source": pk["src"] + " (sustained bf16)",
                     "launches_per_step": len(prof) // max(1, n_prof_steps), "ms_per_step": tc_ms / max(1, n_prof_steps),
                     "step_frac_of_bf16_peak": seg_s / world * FLOPS_PER_SEGMENT[args.model] / (pk["tf_sustained"] * 1e12)},
        # algorithmic bytes per frame: hop * 4 B of waveform read once + 80 * 4 B written = 1520 B
        "logmel": {"workload": "log-mel of %d segments x 58624 samples -> %d frames each (BASELINE configs[4] cell)"
                               % (LM_BIG, FRAMES),
                   "frames_per_s": LM_BIG * FRAMES / (lm_big_ms / 1e3), "ms": lm_big_ms, "bound": "hbm",
                   "achieved_gbs": LM_BIG * FRAMES * 1520.0 / (lm_big_ms / 1e3) / 1e9, "peak_gbs": pk["hbm"],
                   "frac": LM_BIG * FRAMES * 1520.0 / (lm_big_ms / 1e3) / 1e9 / pk["hbm"],
                   "in_step": {"workload": "the step's own batch: %d segments -> %d frames each" % (B, FRAMES),
                               "frames_per_s": B * FRAMES / (lm_ms / 1e3), "ms": lm_ms,
                               "frac": B * FRAMES * 1520.0 / (lm_ms / 1e3) / 1e9 / pk["hbm"]}},
        "loss_last": last,
    }
    if world == 1 and not args.no_torch_gpu_baseline:
        line["torch_gpu_baseline"] = torch_gpu_baseline(args)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.batch, model=args.model)
    print(json.dumps(line), flush=True)


def _port_cpu(sample_batch, steps, warmup, model):
    """Fallback when baseline/_ref is not staged: the oracle port of reference Trainer.run on the host cores."""
    from oracle import jdcnet_torch as J, train_step as TS
    from pitchextractor_b200 import synthetic
    from pitchextractor_b200.model import JDCNet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = JDCNet(num_class=1, sequence_model_config=dict(MODEL_CFG, model_type=model)).state_dict()
    ref = TS.ReferenceStep(sd, J.default_config(model))
    waves, f0 = synthetic.make_batch(sample_batch, seed=4321)
    crops = np.arange(sample_batch) % 4
    for _ in range(warmup):
        ref.step(waves, f0, crops)
    t0 = time.perf_counter()
    for _ in range(steps):
        last = ref.step(waves, f0, crops)
    dt = (time.perf_counter() - t0) / steps
    return {"segments_per_s": sample_batch / dt, "s_per_step": dt, "segments_per_step": sample_batch, "cores": cores,
            "loss_last": last}


def reference_cpu(model, batch, steps, warmup, budget_s):
    """The reference's own CPU implementation of the step (fp32; the reference itself switches AMP and checkpointing
    off on CPU, trainer.py:64,103) on all host cores: unmodified code from baseline/_ref when staged (kind "reference"),
    else the oracle port (kind "port")."""
    import contextlib
    from baseline import ref_arm
    cfg = dict(MODEL_CFG, model_type=model)
    with contextlib.redirect_stdout(sys.stderr):  # the reference prints from build_optimizer
        if ref_arm.available():
            r, kind = ref_arm.time_cpu(cfg, batch, steps, warmup, budget_s=budget_s), "reference"
        else:
            n = min(batch, 8)
            r, kind = _port_cpu(n, steps, warmup, model), "port"
    r["kind"] = kind
    what = ("unmodified reference (baseline/_ref): MelDataset._build_training_example per sample (torchaudio, CPU) + "
            "Collater + Trainer.run fp32 + AdamW/OneCycleLR" if kind == "reference" else
            "oracle port of Trainer.run: per-sample CPU log-mel + fp32 fwd/bwd + AdamW")
    r["sample"] = "%d timed step(s) of %d segments (batch of the workload: %d), %s, %d host threads" % (
        steps, r["segments_per_step"], batch, what, r["cores"])
    return r


def cpu_baseline(batch, model="transformer"):
    r = reference_cpu(model, batch, steps=1, warmup=1, budget_s=25.0)
    return {"value": r["segments_per_s"], "unit": "segments/s", "cores": r["cores"], "kind": r["kind"],
            "sample": r["sample"], "s_per_step": r["s_per_step"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    r = reference_cpu(args.model, args.batch, args.steps, max(1, args.warmup), budget_s=240.0)
    v = r["segments_per_s"]
    print(json.dumps({
        "impl": "reference", "metric": "train_segments_per_s", "value": v, "unit": "segments/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD[args.model] % args.batch,
                   "global_batch": args.batch, "parallelism": "cpu", "segments_per_step": r["segments_per_step"]},
        "cpu_baseline": {"value": v, "unit": "segments/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": v, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss_last": r["loss_last"]}), flush=True)


def torch_gpu_baseline(args):
    """The unmodified reference model + Trainer.run on this GPU through torch (cuDNN / cuBLAS / SDPA), as shipped (fp16
    autocast + GradScaler + checkpointing) and in its faster settings -- SURVEY 8d: "that is the real bar"."""
    import contextlib
    from baseline import ref_arm
    if not ref_arm.available():
        return {"unavailable": "baseline/_ref not staged"}
    with contextlib.redirect_stdout(sys.stderr):
        return ref_arm.time_gpu(dict(MODEL_CFG, model_type=args.model), args.batch, steps=5, warmup=4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="segments per GPU per step")
    ap.add_argument("--model", default="transformer", choices=["transformer", "bilstm"],
                    help="sequence model (default: the configuration the metric is quoted on)")
    ap.add_argument("--pool", type=int, default=3, help="distinct synthetic batches cycled through")
    ap.add_argument("--cpu-sample", type=int, default=8, help="(unused; the CPU sample is sized from a time budget)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-gpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
