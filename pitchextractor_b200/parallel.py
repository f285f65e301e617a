"""Data parallelism: one process per GPU, gradients all-reduced with NCCL over NVLink in arena-contiguous buckets that
are launched on a side stream as soon as the backward pass has finished writing them (overlap with the rest of the
backward), summed and divided by the world size inside the fused optimizer kernel.

The reference has no distributed code; BatchNorm statistics stay per rank, which is the
reference's single-process semantics at the per-GPU batch size.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style environment -> (rank, world, local_rank); no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


def bucket_ranges(names, offsets, total):
    """Arena ranges in the order the backward pass completes them: heads + sequence_detector, sequence_classifier,
    then the conv trunk + detector_conv.  Returns [(lo, hi, tag)] covering [0, total)."""
    first = {}
    for n, off in zip(names, offsets):
        key = n.split(".")[0]
        first.setdefault(key, off)
    lo_cls = first.get("sequence_classifier", total)
    lo_det = first.get("sequence_detector", total)
    out = []
    if lo_det < total:
        out.append((lo_det, total, "sequence_detector+heads"))
    if lo_cls < lo_det:
        out.append((lo_cls, lo_det, "sequence_classifier"))
    out.append((0, min(lo_cls, lo_det), "trunk"))
    return out


class GradReducer:
    """Bucketed asynchronous all-reduce of a flat gradient tensor.

    ``ready(tag)`` is called by the engine right after the kernels that complete a bucket were enqueued; the
    all-reduce is issued on ``comm_stream`` behind an event, so it overlaps with the remaining backward kernels.
    ``wait()`` makes the current stream wait for every outstanding bucket (call before the optimizer step).
    """

    def __init__(self, flat_grad, buckets, group=None):
        self.flat = flat_grad
        self.buckets = {tag: (lo, hi) for lo, hi, tag in buckets}
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat_grad.is_cuda
        self.comm_stream = torch.cuda.Stream(device=flat_grad.device) if self.cuda else None
        self._pending = []
        self._done = set()

    def begin_step(self):
        self._pending = []
        self._done = set()

    def ready(self, tag):
        if self.world == 1 or tag in self._done or tag not in self.buckets:
            return
        self._done.add(tag)
        lo, hi = self.buckets[tag]
        view = self.flat[lo:hi]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append(work)
        else:
            self._pending.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        for tag in self.buckets:  # anything the engine did not announce goes now
            self.ready(tag)
        for w in self._pending:
            w.wait()  # on CUDA: makes the current stream wait for the collective
        self._pending = []


def broadcast_parameters(flat, src=0, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
