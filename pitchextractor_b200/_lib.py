"""ctypes binding of the C-ABI library (``include/pitchextractor_b200.h``).

The product path is CUDA-only: if ``libpe_b200.so`` is missing or a call fails, a ``RuntimeError`` is raised --
there is no CPU / eager fallback.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpe_b200.so")

PE_OUT_F32, PE_OUT_BF16, PE_OUT_F32_ATOMIC = 0, 1, 2
PE_ACT_NONE, PE_ACT_GELU, PE_ACT_GELU_SAVE_GRAD = 0, 1, 2
PE_AUX_NONE, PE_AUX_ADD, PE_AUX_GELU_GRAD, PE_AUX_MUL = 0, 1, 2, 3

_ERRORS = {-1: "bad shape / argument", -2: "workspace too small", -3: "device is not sm_100 (B200)",
           -4: "CUDA driver entry point unavailable", -5: "kernel launch failed"}


class Epilogue(ctypes.Structure):
    _fields_ = [
        ("out", ctypes.c_void_p), ("ldc", ctypes.c_longlong), ("out_mode", ctypes.c_int), ("act", ctypes.c_int),
        ("out2", ctypes.c_void_p), ("ld2", ctypes.c_longlong), ("bias", ctypes.c_void_p),
        ("aux", ctypes.c_void_p), ("ld_aux", ctypes.c_longlong), ("aux_mode", ctypes.c_int),
        ("drop_thresh", ctypes.c_uint), ("drop_scale", ctypes.c_float), ("drop_seed", ctypes.c_ulonglong),
        ("alpha", ctypes.c_float), ("stats", ctypes.c_void_p), ("stats_mode", ctypes.c_int),
        ("stats_x", ctypes.c_void_p), ("stats_scale", ctypes.c_void_p), ("stats_shift", ctypes.c_void_p),
        ("stats_slope", ctypes.c_float), ("debug", ctypes.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "pitchextractor_b200: %s not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.pe_version.restype = ctypes.c_int
    return _lib


def ptr(t):
    """Device pointer of a tensor (or None)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc, what):
    if rc != 0:
        raise RuntimeError("pitchextractor_b200.%s failed: %s (rc=%d)" % (what, _ERRORS.get(rc, "unknown"), rc))


# kernels launched per C-ABI call (for bench.py's gpu_launches count)
_LAUNCHES_PER_CALL = {"pe_logmel_f32": 2, "pe_bn_act_pool_bwd": 3, "pe_attn_bwd": 2, "pe_heads_loss": 2}
launch_count = 0
_next_salt_slot = 0


def new_salt_slot():
    """A dropout-salt slot (0..255) for one engine: its seeds carry the slot in their top 8 bits (pe_set_step_salt)."""
    global _next_salt_slot
    slot = _next_salt_slot % 256
    _next_salt_slot += 1
    return slot


# When set to a list, every C-ABI call appends (name, start_event, end_event): per-entry-point GPU time with warm
# caches (tools/step_breakdown.py).  Adds event records only, no synchronisation.
TIMING = None


def call(name, *args):
    global launch_count
    fn = getattr(lib(), name)
    fn.restype = ctypes.c_int
    if TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(*args), name)
        e1.record()
        TIMING.append((name, e0, e1))
    else:
        check(fn(*args), name)
    launch_count += _LAUNCHES_PER_CALL.get(name, 1)


def drop_thresh(p_drop):
    """16-bit keep threshold (Philox lanes) and the exactly matching scale for dropout probability p_drop."""
    if p_drop <= 0.0:
        return 0, 1.0
    t = max(1, min(65535, int(round((1.0 - p_drop) * 65536.0))))
    return t, 65536.0 / t


def attn_drop_thresh(p_drop):
    """32-bit keep threshold (hash) and scale for the attention-probability dropout."""
    if p_drop <= 0.0:
        return 0, 1.0
    keep = 1.0 - p_drop
    return min(int(keep * 4294967296.0), 4294967295), 1.0 / keep
