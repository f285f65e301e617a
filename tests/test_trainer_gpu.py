"""Trainer.run on the CUDA path vs the CPU port of the reference step (oracle/train_step.py): same init, same
waveform batches, dropout off on both sides, equal number of steps."""
import numpy as np
import pytest
import torch

import golden_inputs as GI

pytestmark = pytest.mark.gpu


def _setup(B=4, steps=3):
    from pitchextractor_b200 import synthetic
    batches = []
    for s in range(steps):
        waves, f0 = synthetic.make_batch(B, seed=100 + s)
        crops = (np.arange(B) + s) % 4
        batches.append((waves, f0, crops.astype(np.int32)))
    return batches


def test_equal_steps_loss_trajectory(built_lib):
    from oracle import jdcnet_torch as J, train_step as TS
    from pitchextractor_b200 import JDCNet, Trainer, build_optimizer
    from pitchextractor_b200.meldataset import align_length
    sd = GI.model_state_dict("transformer")
    batches = _setup()
    ref = TS.ReferenceStep(sd, J.default_config("transformer"), max_lr=3e-4, epochs=100, steps_per_epoch=1000)
    ref_losses = [ref.step(w, f, c, dropout=False) for w, f, c in batches]

    model = JDCNet(num_class=1, sequence_model_config=GI.model_config("transformer"))
    model.load_state_dict(sd)
    model = model.cuda()
    opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {},
                                  "scheduler_params": {"max_lr": 3e-4, "pct_start": 0.0, "epochs": 100,
                                                       "steps_per_epoch": 1000}})
    tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda")
    model.engine.dropout_enabled = False
    model.train()
    ours = []
    for w, f, c in batches:
        f0 = np.stack([align_length(f[b], f.shape[1])[c[b]:c[b] + 192] for b in range(len(c))]).astype(np.float32)
        sil = (f0 == 0).astype(np.float32)
        batch = tuple(torch.from_numpy(a) for a in (w, f0, sil, c))
        ours.append(tr.run(batch))
    for a, b in zip(ours, ref_losses):
        print("cuda", a, "cpu-port", b)
        # bf16 tensor-core mode vs fp32 reference: per-step loss within 2e-2 relative
        assert abs(a["loss"] - b["loss"]) <= 2e-2 * abs(b["loss"])
        assert abs(a["sil"] - b["sil"]) <= 3e-2 * abs(b["sil"]) + 5e-3
    # parameters after equal steps: the update direction is sign-like for AdamW's first steps, compare drift
    new_sd = model.state_dict()
    for k in ("classifier.weight", "sequence_classifier.model.layers.3.linear2.weight", "detector_conv.0.weight"):
        d_ours = (new_sd[k].cpu().float() - sd[k]).flatten()
        d_ref = (ref.sd[k].detach() - sd[k]).flatten()
        cos = torch.nn.functional.cosine_similarity(d_ours, d_ref, dim=0).item()
        print(k, "update cosine", cos)
        assert cos > 0.9, (k, cos)
    assert tr.steps == len(batches)
    assert model.conv_block._modules["1"].num_batches_tracked.item() == len(batches)


def test_checkpoint_roundtrip_and_mel_batches(built_lib, tmp_path):
    from pitchextractor_b200 import JDCNet, Trainer, build_optimizer, LogMel
    torch.manual_seed(0)
    mk = lambda: JDCNet(num_class=1, sequence_model_config=GI.model_config("transformer")).cuda()
    model = mk()
    opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {}, "scheduler_params": {}})
    tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda",
                 config={})
    (w, f, c), = _setup(B=2, steps=1)
    mel = LogMel("cuda")(torch.from_numpy(w).cuda())[:, :, :192][:, None].contiguous()  # reference batch layout
    f0 = torch.from_numpy(f[:, :192].copy())
    sil = (f0 == 0).float()
    out = tr.run((mel.cpu(), f0, sil))
    assert np.isfinite(out["loss"])
    path = str(tmp_path / "ckpt" / "epoch_1.pth")
    tr.save_checkpoint(path)
    state = torch.load(path, map_location="cpu")
    assert set(state) == {"optimizer", "scheduler", "steps", "epochs", "model"}
    assert state["model"]["conv_block.0.weight"].shape == (64, 1, 3, 3)
    model2 = mk()
    opt2, sched2 = build_optimizer({"params": model2.parameters(), "optimizer_params": {}, "scheduler_params": {}})
    tr2 = Trainer(model=model2, optimizer=opt2, scheduler=sched2, loss_config={"lambda_f0": 0.1}, device="cuda",
                  config={})
    tr2.load_checkpoint(path)
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
    model.engine.dropout_enabled = False
    model2.engine.dropout_enabled = False
    o1 = tr.run((mel.cpu(), f0, sil))
    o2 = tr2.run((mel.cpu(), f0, sil))
    assert abs(o1["loss"] - o2["loss"]) <= 1e-3 * abs(o1["loss"]), (o1, o2)
    # the restored optimizer state itself: moments, step count and schedule position equal the originals' after the
    # extra step both trainers just made (a broken restore of exp_avg / exp_avg_sq / step would show here, not in a loss
    # that is computed before the update)
    assert opt2._steps == opt._steps == 2 and sched2.last_epoch == sched.last_epoch
    for key in ("exp_avg", "exp_avg_sq"):
        a, b = getattr(opt, key), getattr(opt2, key)
        # (the two trainers' second-step gradients differ by atomic-summation order, amplified by BatchNorm at batch 2)
        assert (a - b).abs().max().item() <= 2e-2 * a.abs().max().item() + 1e-12, key
    upd = (model.engine.flat - model2.engine.flat).norm().item() / model.engine.flat.norm().item()
    assert upd <= 2e-4, upd
    o1b, o2b = tr.run((mel.cpu(), f0, sil)), tr2.run((mel.cpu(), f0, sil))
    assert abs(o1b["loss"] - o2b["loss"]) <= 2e-3 * abs(o1b["loss"]), (o1b, o2b)
    model.eval()
    ev = model.engine.eval_loss(mel, f0, sil).tolist()
    assert np.isfinite(ev).all()


def test_predict_f0_matches_reference_chunking(built_lib):
    """Batched chunked inference vs the notebook's per-chunk loop restated on the fp32 oracle (eval mode)."""
    from oracle import jdcnet_torch as J, logmel_np
    from pitchextractor_b200 import JDCNet, predict_f0
    sd = GI.model_state_dict("transformer")
    rng = np.random.default_rng(3)
    audio = (0.1 * rng.standard_normal(24000 * 4 + 137)).astype(np.float32)  # 321 frames -> chunks at 0, 144, 288
    mel = torch.from_numpy(logmel_np.log_mel(audio)).float()
    total, preds = mel.shape[-1], []
    for start in range(0, total, 144):
        end = min(start + 192, total)
        chunk = torch.nn.functional.pad(mel[:, start:end], (0, 192 - (end - start)))[None, None].transpose(-1, -2)
        cls, _ = J.jdcnet_forward(sd, chunk, J.default_config("transformer"), training=False)
        preds.append(cls.squeeze().numpy()[:end - start])
    ref = np.concatenate(preds)
    model = JDCNet(num_class=1, sequence_model_config=GI.model_config("transformer"))
    model.load_state_dict(sd)
    got = predict_f0(model.cuda(), audio)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-2 * np.abs(ref).max() + 1e-2


@pytest.mark.parametrize("model_type", ["transformer", "bilstm"])
def test_cuda_graph_replay_matches_eager(built_lib, model_type):
    """The training step is captured into a CUDA graph after two eager steps.  With dropout ON, graph replay must
    (a) follow the eagerly launched trajectory step for step (same per-step seeds via the device-side salt) and
    (b) draw fresh masks on every replay (loss on an identical batch changes from replay to replay)."""
    from pitchextractor_b200 import JDCNet, Trainer, build_optimizer
    sd = GI.model_state_dict(model_type)
    B = 4
    g = torch.Generator().manual_seed(5)
    mel = (torch.randn(B, 1, 80, 192, generator=g) * 2 - 4).cuda()
    f0 = (torch.rand(B, 192, generator=g) * 300 + 80) * (torch.rand(B, 192, generator=g) > 0.3)
    sil = (f0 == 0).float()

    def trajectory(use_graph, lr):
        model = JDCNet(num_class=1, sequence_model_config=GI.model_config(model_type))
        model.load_state_dict(sd)
        model = model.cuda()
        opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {"lr": lr},
                                      "scheduler_params": {"max_lr": lr, "epochs": 100, "steps_per_epoch": 1000}})
        tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda")
        model.engine.use_graph = use_graph
        model.train()
        out = [tr.run((mel, f0, sil))["loss"] for _ in range(6)]
        captured = any("segments" in e for e in model.engine._graphs.values())
        return out, captured, model

    eager, cap_e, _ = trajectory(False, 2e-4)
    graphed, cap_g, model = trajectory(True, 2e-4)
    print("eager  ", eager)
    print("graphed", graphed)
    assert cap_g and not cap_e
    for a, b in zip(eager, graphed):  # fp32 atomics reorder sums from run to run; nothing else may differ
        assert abs(a - b) <= 5e-3 * abs(a), (eager, graphed)
    assert model.conv_block._modules["1"].num_batches_tracked.item() == 6
    # frozen weights (lr = 0): any change between replays is the dropout mask alone
    frozen, _, _ = trajectory(True, 0.0)
    print("frozen ", frozen)
    assert len(set(frozen[2:])) == len(frozen[2:]), frozen  # every replay draws its own masks


def test_cuda_graph_segments_with_reducer(built_lib):
    """Data parallel replay: the capture is cut where the backward pass announces a finished gradient bucket and the
    reducer is called from the host between the segments, in bucket order, exactly once per step."""
    from pitchextractor_b200 import JDCNet, Trainer, build_optimizer
    sd = GI.model_state_dict("transformer")
    g = torch.Generator().manual_seed(6)
    mel = (torch.randn(2, 1, 80, 192, generator=g) * 2 - 4).cuda()
    f0 = torch.rand(2, 192, generator=g) * 300
    sil = (f0 < 60).float()

    class Recorder:
        def __init__(self):
            self.calls = []

        def begin_step(self):
            self.calls.append("begin")

        def ready(self, tag):
            self.calls.append(tag)

        def wait(self):
            self.calls.append("wait")

    def run(use_graph):
        model = JDCNet(num_class=1, sequence_model_config=GI.model_config("transformer"))
        model.load_state_dict(sd)
        model = model.cuda()
        opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {}, "scheduler_params": {}})
        tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda")
        eng = model.engine
        eng.use_graph = use_graph
        eng.dropout_enabled = False
        eng.reducer = Recorder()
        model.train()
        losses = [tr.run((mel, f0, sil))["loss"] for _ in range(5)]
        return losses, eng

    eager, _ = run(False)
    graphed, eng = run(True)
    ent = [e for e in eng._graphs.values() if "segments" in e]
    # the two encoder stacks run concurrently, so their buckets complete at the same cut; the trunk's is the last
    assert len(ent) == 1 and len(ent[0]["segments"]) == 2
    per_step = ["begin", "sequence_detector+heads", "sequence_classifier", "trunk", "wait"]
    assert eng.reducer.calls == per_step * 5, eng.reducer.calls
    for a, b in zip(eager, graphed):
        assert abs(a - b) <= 5e-3 * abs(a), (eager, graphed)


def test_prefetched_batches_give_the_same_steps(built_lib):
    """Trainer.prefetched copies batch i+1 on a copy stream while step i runs; the steps themselves are unchanged."""
    from pitchextractor_b200 import JDCNet, Trainer, build_optimizer
    sd = GI.model_state_dict("transformer")
    g = torch.Generator().manual_seed(9)
    batches = []
    for _ in range(4):
        mel = (torch.randn(2, 1, 80, 192, generator=g) * 2 - 4).pin_memory()
        f0 = (torch.rand(2, 192, generator=g) * 300).pin_memory()
        batches.append((mel, f0, (f0 < 60).float().pin_memory()))

    def run(prefetch):
        model = JDCNet(num_class=1, sequence_model_config=GI.model_config("transformer"))
        model.load_state_dict(sd)
        model = model.cuda()
        opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {}, "scheduler_params": {}})
        tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda")
        model.engine.dropout_enabled = False
        model.train()
        src = tr.prefetched(batches) if prefetch else batches
        return [tr.run(b)["loss"] for b in src]

    a, b = run(False), run(True)
    assert len(a) == len(b) == 4
    for x, y in zip(a, b):
        assert abs(x - y) <= 5e-3 * abs(x), (a, b)


def test_cuda_graph_survives_a_batch_size_change(built_lib):
    """The short last batch of an epoch re-allocates the activation buffers a captured step points at: the graphs are
    dropped and re-captured, and the trajectory equals the eagerly launched one."""
    from pitchextractor_b200 import JDCNet, Trainer, build_optimizer
    sd = GI.model_state_dict("transformer")
    g = torch.Generator().manual_seed(12)
    sizes = [4, 4, 4, 4, 2, 4, 4, 4, 4]
    batches = []
    for B in sizes:
        mel = (torch.randn(B, 1, 80, 192, generator=g) * 2 - 4).cuda()
        f0 = (torch.rand(B, 192, generator=g) * 300).cuda()
        batches.append((mel, f0, (f0 < 60).float()))

    def run(use_graph):
        model = JDCNet(num_class=1, sequence_model_config=GI.model_config("transformer"))
        model.load_state_dict(sd)
        model = model.cuda()
        opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {}, "scheduler_params": {}})
        tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda")
        model.engine.use_graph = use_graph
        model.engine.dropout_enabled = False
        model.train()
        out = [tr.run(b)["loss"] for b in batches]
        return out, any("segments" in e for e in model.engine._graphs.values())

    eager, _ = run(False)
    graphed, captured = run(True)
    assert captured  # re-captured after the size change
    for a, b in zip(eager, graphed):
        assert abs(a - b) <= 5e-3 * abs(a), (eager, graphed)


def test_run_pipelined_yields_every_steps_losses(built_lib):
    """Trainer.run_pipelined == Trainer.run per batch (same values, same order), one step behind the GPU."""
    from pitchextractor_b200 import JDCNet, Trainer, build_optimizer
    sd = GI.model_state_dict("transformer")
    g = torch.Generator().manual_seed(21)
    batches = []
    for _ in range(5):
        mel = (torch.randn(2, 1, 80, 192, generator=g) * 2 - 4).pin_memory()
        f0 = (torch.rand(2, 192, generator=g) * 300).pin_memory()
        batches.append((mel, f0, (f0 < 60).float().pin_memory()))

    def make():
        model = JDCNet(num_class=1, sequence_model_config=GI.model_config("transformer"))
        model.load_state_dict(sd)
        model = model.cuda()
        opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {}, "scheduler_params": {}})
        tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda")
        model.engine.dropout_enabled = False
        model.train()
        return tr

    a = [make_out["loss"] for make_out in (lambda tr: [tr.run(b) for b in batches])(make())]
    b = [o["loss"] for o in make().run_pipelined(batches)]
    assert len(a) == len(b) == 5
    for x, y in zip(a, b):
        assert abs(x - y) <= 5e-3 * abs(x), (a, b)


def test_train_and_eval_epoch_return_the_reference_keys(built_lib):
    """Trainer._train_epoch / _eval_epoch (reference trainer.py:254-291) over a list of batches."""
    from pitchextractor_b200 import JDCNet, Trainer, build_optimizer
    sd = GI.model_state_dict("transformer")
    g = torch.Generator().manual_seed(31)
    batches = []
    for _ in range(3):
        mel = torch.randn(2, 1, 80, 192, generator=g) * 2 - 4
        f0 = torch.rand(2, 192, generator=g) * 300
        batches.append((mel, f0, (f0 < 60).float()))
    model = JDCNet(num_class=1, sequence_model_config=GI.model_config("transformer"))
    model.load_state_dict(sd)
    model = model.cuda()
    opt, sched = build_optimizer({"params": model.parameters(), "optimizer_params": {}, "scheduler_params": {}})
    tr = Trainer(model=model, optimizer=opt, scheduler=sched, loss_config={"lambda_f0": 0.1}, device="cuda",
                 train_dataloader=batches, val_dataloader=batches)
    out = tr._train_epoch()
    assert set(out) == {"train/loss", "train/f0", "train/sil", "train/learning_rate"}
    assert tr.epochs == 1 and tr.steps == 3 and all(v == v for v in out.values())
    ev = tr._eval_epoch()
    assert set(ev) == {"eval/loss", "eval/f0", "eval/sil"} and all(v == v for v in ev.values())
