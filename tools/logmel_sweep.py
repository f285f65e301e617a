"""log-mel throughput sweep (BASELINE configs[4]: 1 s - 60 s clips, batch 1 - 1024, vs torchaudio): frames/s and the
fraction of the HBM roofline (1520 algorithmic bytes per frame: hop * 4 B read + 80 * 4 B written) of this repo's
tcgen05 kernel, next to torchaudio's MelSpectrogram + log normalisation (the reference's arithmetic, meldataset.py:77,
644,650) batched on the same GPU and on the host cores.  Cells larger than ~12 GB of waveform are skipped.

    python tools/logmel_sweep.py [out.json]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pitchextractor_b200.mel import LogMel  # noqa: E402

peak = 6541.8
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

import torchaudio  # noqa: E402

lm = LogMel("cuda")
ta_gpu = torchaudio.transforms.MelSpectrogram(sample_rate=24000, n_mels=80, n_fft=1024, win_length=1024,
                                              hop_length=300).cuda()
ta_cpu = torchaudio.transforms.MelSpectrogram(sample_rate=24000, n_mels=80, n_fft=1024, win_length=1024, hop_length=300)
torch.set_num_threads(os.cpu_count() or 1)


def ref_mel(tr, w):
    return (torch.log(1e-5 + tr(w)) + 4.0) / 4.0


def gpu_ms(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
for secs in (1, 2, 2.4427, 5, 10, 30, 60):
    L = int(round(secs * 24000)) // 4 * 4
    for B in (1, 4, 16, 64, 256, 1024):
        if B * L * 4 > 12e9:
            continue
        w = torch.randn(B, L, device="cuda") * 0.1
        frames = B * (1 + L // 300)
        n = 20 if frames < 2e6 else 5
        ms = gpu_ms(lambda: lm(w), n)
        row = dict(seconds=secs, batch=B, frames=frames, ms=ms, frames_per_s=frames / (ms * 1e-3),
                   hbm_frac=frames / (ms * 1e-3) * 1520 / 1e9 / peak)
        try:
            if B * L * 4 <= 3e9:  # torch.stft materialises the complex spectrogram: 8 KB per frame
                ms_ta = gpu_ms(lambda: ref_mel(ta_gpu, w), max(2, n // 4))
                row["torchaudio_cuda_ms"] = ms_ta
                row["torchaudio_cuda_frames_per_s"] = frames / (ms_ta * 1e-3)
        except Exception as e:  # out of memory on the largest cells
            row["torchaudio_cuda_error"] = type(e).__name__
            torch.cuda.empty_cache()
        if frames <= 120000:
            wc = w.cpu()
            ref_mel(ta_cpu, wc)
            t0 = time.perf_counter()
            ref_mel(ta_cpu, wc)
            dt = time.perf_counter() - t0
            row["torchaudio_cpu_ms"] = dt * 1e3
            row["torchaudio_cpu_frames_per_s"] = frames / dt
            row["cpu_threads"] = os.cpu_count()
        rows.append(row)
        print("%6.2f s x %4d : %9.3f ms %8.1f M frames/s %.4f of HBM roofline | torchaudio cuda %s ms | cpu %s ms" % (
            secs, B, ms, row["frames_per_s"] / 1e6, row["hbm_frac"],
            "%.3f" % row["torchaudio_cuda_ms"] if "torchaudio_cuda_ms" in row else "-",
            "%.1f" % row["torchaudio_cpu_ms"] if "torchaudio_cpu_ms" in row else "-"), flush=True)
        del w
        torch.cuda.empty_cache()
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "logmel_sweep.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump({"peak_hbm_gbs": peak, "bytes_per_frame": 1520, "rows": rows}, open(out, "w"), indent=1)
