"""log-mel throughput sweep (BASELINE configs[4]): clip length x batch, frames/s and fraction of the HBM roofline
(1520 algorithmic bytes per frame), plus the worker-phase cycle breakdown of the tcgen05 kernel (tuning aid)."""
import ctypes, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchextractor_b200 import _lib
from pitchextractor_b200.mel import LogMel
peak = 6541.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
lm = LogMel("cuda")
rows = []
for secs, batches in ((1, (1, 64, 1024)), (2.4427, (16, 64, 512)), (10, (1, 64, 256)), (60, (1, 16, 64))):
    L = int(round(secs * 24000)) // 4 * 4
    for B in batches:
        w = torch.randn(B, L, device="cuda") * 0.1
        for _ in range(3):
            lm(w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            lm(w)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        frames = B * (1 + L // 300)
        fps = frames / (ms * 1e-3)
        rows.append(dict(seconds=secs, batch=B, frames=frames, ms=ms, frames_per_s=fps, hbm_frac=fps * 1520 / 1e9 / peak))
        print("%6.2f s x %4d : %8.3f ms  %8.1f M frames/s  %.4f of HBM roofline" % (secs, B, ms, fps / 1e6, fps * 1520 / 1e9 / peak), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/logmel_sweep.json", "w"), indent=1)
dbg = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
_lib.lib().pe_logmel_set_debug(ctypes.c_void_p(dbg.data_ptr()))
w = torch.randn(512, 58624, device="cuda") * 0.1
lm(w); torch.cuda.synchronize()
_lib.lib().pe_logmel_set_debug(None)
d = dbg.float().mean(0).tolist()
tiles = 512 * 25 / 148
names = ["wait raw", "pre-pass", "wait mma1", "twiddle", "wait mma2", "unpack", "mel+store"]
print("cycles per tile (8 frames):", {n: round(v / tiles) for n, v in zip(names, d)}, "total", round(sum(d) / tiles))
