"""Host-side constant tables and the batched GPU log-mel transform (mirror of the reference's
``MelDataset.to_melspec`` + normalisation, meldataset.py:77,644,650)."""
import math

import numpy as np
import torch

from . import ops

DEFAULT_MEL_PARAMS = {  # meldataset.py:34-40
    "sample_rate": 24000,
    "n_mels": 80,
    "n_fft": 1024,
    "win_length": 1024,
    "hop_length": 300,
}


def mel_filterbank(sample_rate, n_fft, n_mels, f_min=0.0, f_max=None):
    """HTK triangular filterbank [n_fft//2+1, n_mels], computed with the same fp32 torch ops (and in the same
    order) as torchaudio.functional.melscale_fbanks (norm=None, mel_scale='htk') so the table is bit-identical
    to the one the reference builds at meldataset.py:77."""
    f_max = float(sample_rate // 2) if f_max is None else f_max
    n_freqs = n_fft // 2 + 1
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).contiguous()


def windowed_dft_basis(n_fft, win_length=None):
    """fp32 [n_fft, ld] with basis[n, 2k] = w[n] cos(2 pi k n / N), basis[n, 2k+1] = w[n] sin(2 pi k n / N)
    (periodic Hann, torch.hann_window default), computed in fp64; ld = 2*(N/2+1) rounded up to 64."""
    win_length = n_fft if win_length is None else win_length
    n_bins = n_fft // 2 + 1
    w = np.zeros(n_fft)
    left = (n_fft - win_length) // 2  # torch.stft centres a short window inside n_fft
    w[left:left + win_length] = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)
    n = np.arange(n_fft)[:, None]
    k = np.arange(n_bins)[None, :]
    # reduce the phase index mod N in integers so the fp64 angle stays small and exact
    ang = 2.0 * np.pi * ((n * k) % n_fft) / n_fft
    ld = ((2 * n_bins + 63) // 64) * 64
    basis = np.zeros((n_fft, ld), dtype=np.float32)
    basis[:, 0:2 * n_bins:2] = (w[:, None] * np.cos(ang)).astype(np.float32)
    basis[:, 1:2 * n_bins:2] = (w[:, None] * np.sin(ang)).astype(np.float32)
    return torch.from_numpy(basis)


def dft32_operand_images():
    """fp16 operand images (hi, lo) of the real 64 x 64 form of the 32-point complex DFT matrix
    F[(c', k), (c, n)] (c, c' in {re, im}), laid out as a 128-byte-swizzled K-major [64 rows x 128 B] tcgen05 operand."""
    k = np.arange(32)[:, None]
    n = np.arange(32)[None, :]
    ang = 2.0 * np.pi * ((k * n) % 32) / 32.0
    F = np.zeros((64, 64))
    F[:32, :32], F[:32, 32:] = np.cos(ang), np.sin(ang)
    F[32:, :32], F[32:, 32:] = -np.sin(ang), np.cos(ang)
    hi = F.astype(np.float16)
    lo = (F - hi.astype(np.float64)).astype(np.float16)
    out = np.zeros((2, 64 * 64), dtype=np.float16)
    o = np.arange(64)[:, None]
    kk = np.arange(64)[None, :]
    elem = (o * 128 + (((kk >> 3) ^ (o & 7)) << 4) + (kk & 7) * 2) // 2
    for i, mat in enumerate((hi, lo)):
        out[i, elem.reshape(-1)] = mat.reshape(-1)
    return torch.from_numpy(out)


def four_step_twiddles():
    """[2, 32, 32] fp32: cos / sin of 2 pi k1 n2 / 1024, indexed [k1][n2]."""
    k1 = np.arange(32)[:, None]
    n2 = np.arange(32)[None, :]
    ang = 2.0 * np.pi * (k1 * n2) / 1024.0
    return torch.from_numpy(np.stack([np.cos(ang), np.sin(ang)]).astype(np.float32))


def banded_filterbank(fb):
    """fb [n_bins, n_mels] -> (start, count, offset int32 [n_mels], weights fp32 [nnz]): each mel filter is one
    contiguous band of bins."""
    fb = fb.cpu().numpy()
    starts, counts, offs, weights = [], [], [], []
    for m in range(fb.shape[1]):
        nz = np.nonzero(fb[:, m])[0]
        lo, hi = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, 1)
        starts.append(lo)
        counts.append(hi - lo)
        offs.append(len(weights))
        weights.extend(fb[lo:hi, m].tolist())
    to_i = lambda a: torch.tensor(a, dtype=torch.int32)
    return to_i(starts), to_i(counts), to_i(offs), torch.tensor(weights, dtype=torch.float32)


ITEM_COMBINE, ITEM_WRITER = 1, 2


def mel_work_items(starts, counts, offsets, lanes=128):
    """Work items of the banded mel product for the tcgen05 log-mel kernel, one per worker lane: int32 [lanes][4] =
    {filter (-1: idle), first bin, number of bins, offset into the weight array | flags << 24}.  The longest filters are
    split in two halves that sit on adjacent lanes (l, l ^ 1; flag ITEM_COMBINE on both, ITEM_WRITER on the even one)
    until every lane has work.  Items are packed into groups of 8 lanes (the lanes one shared-memory wavefront of a
    16-byte load serves) such that (a) the 8 first bins are distinct modulo 8 -- the power-spectrum rows are 48 bytes
    apart, so rows k that differ modulo 8 lie in different 16-byte bank groups and the loads of a group never
    conflict while it walks its bands in lockstep -- and (b) lengths are similar, longest groups first, so the lanes of a
    warp run similar trip counts."""
    n = len(counts)
    if n > lanes:
        raise ValueError("the tcgen05 log-mel kernel handles at most %d mel filters" % lanes)
    order = sorted(range(n), key=lambda m: -int(counts[m]))
    n_split = min(lanes - n, sum(1 for m in order if counts[m] >= 2))
    units = []  # (length, [items]) -- a split filter is one unit of two adjacent items
    for rank, m in enumerate(order):
        st, c, off = int(starts[m]), int(counts[m]), int(offsets[m])
        if rank < n_split:
            c0 = (c + 1) // 2
            if c0 % 8 == 0 and c - c0 > 1:
                c0 += 1  # keep the two halves' first bins distinct modulo 8
            units.append((c0, [(m, st, c0, off, ITEM_COMBINE | ITEM_WRITER), (m, st + c0, c - c0, off + c0, ITEM_COMBINE)]))
        else:
            units.append((c, [(m, st, c, off, ITEM_WRITER)]))
    units.sort(key=lambda u: -u[0])
    groups, remaining = [], units
    while remaining:
        used, group, rest = set(), [], []
        for length, items in remaining:
            res = [it[1] % 8 for it in items]
            fits = len(group) + len(items) <= 8 and len(set(res)) == len(res) and not (set(res) & used)
            if fits:
                group.extend(items)
                used.update(res)
            else:
                rest.append((length, items))
        if len(group) < 8:  # residues could not all be distinct: fill up with the longest leftovers
            keep = []
            for length, items in rest:
                if len(group) + len(items) <= 8 and (len(group) % 2 == 0 or len(items) == 1):
                    group.extend(items)
                else:
                    keep.append((length, items))
            rest = keep
        groups.append(group)
        remaining = rest
    items = []
    for g in groups:
        assert len(g) <= 8
        items.extend(g + [(-1, 0, 0, 0, 0)] * ((8 - len(g)) if len(items) + 8 <= lanes else 0))
    items = items[:lanes] if len(items) <= lanes else None
    if items is None:  # could not pack into 8-lane groups within the lane budget: plain longest-first order
        items = [it for _, its in units for it in its]
    out = np.full((lanes, 4), 0, dtype=np.int64)
    out[:, 0] = -1
    for i, (m, st, c, off, fl) in enumerate(items):
        out[i] = (m, st, c, off | (fl << 24))
    # split pairs must sit on lanes (l, l ^ 1)
    for i in range(0, lanes, 2):
        a, b = out[i], out[i + 1]
        if ((a[3] >> 24) & ITEM_COMBINE) or ((b[3] >> 24) & ITEM_COMBINE):
            assert a[0] == b[0] and ((a[3] >> 24) & ITEM_COMBINE) and ((b[3] >> 24) & ITEM_COMBINE), (i, a, b)
    return torch.from_numpy(out.astype(np.int32))


_TABLE_CACHE = {}


def logmel_tables(device, sample_rate=24000, n_fft=1024, win_length=1024, hop_length=300, n_mels=80, **_):
    """Immutable constant tables, built once per (device, params)."""
    key = (str(device), sample_rate, n_fft, win_length, hop_length, n_mels)
    if key not in _TABLE_CACHE:
        fb = mel_filterbank(sample_rate, n_fft, n_mels)
        tab = {
            "n_fft": n_fft, "hop": hop_length, "n_mels": n_mels, "sr": sample_rate,
            "basis": windowed_dft_basis(n_fft, win_length).to(device),
            "fb": fb.to(device),
            "tc": n_fft == 1024 and hop_length % 4 == 0 and hop_length <= 320 and n_mels <= 128,
        }
        if tab["tc"]:
            win = np.zeros(n_fft)
            left = (n_fft - win_length) // 2
            win[left:left + win_length] = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)
            st, cnt, off, w = banded_filterbank(fb)
            # power-of-two scales of the fp16-split pipeline (exact): window x 2^11 keeps the split remainders of quiet
            # samples in fp16's normal range, twiddles x 2^-5 keep the stage-2 operand inside fp16's range for |x| <= 16;
            # the power spectra then carry (2 * 2^6)^2 = 2^14 (the factor 2 is the two-frames-per-FFT unpacking)
            tab.update(win=torch.from_numpy((win * 2048.0).astype(np.float32)).to(device),
                       fmat=dft32_operand_images().to(device), tw=(four_step_twiddles() / 32.0).to(device),
                       mel_start=st.to(device), mel_count=cnt.to(device),
                       mel_off=off.to(device), mel_w=(w * 2.0 ** -14).to(device), mel_nnz=int(w.numel()),
                       mel_items=mel_work_items(st.tolist(), cnt.tolist(), off.tolist()).to(device))
            tab["tc"] = tab["mel_nnz"] <= 1536
        _TABLE_CACHE[key] = tab
    return _TABLE_CACHE[key]


class LogMel:
    """Batched replacement of ``(log(1e-5 + MelSpectrogram(wave)) + 4) / 4``: wave [B, L] fp32 (cuda) ->
    [B, n_mels, T] with T = 1 + L // hop.  CUDA-only."""

    def __init__(self, device="cuda", impl="auto", **mel_params):
        """impl: "auto" (tcgen05 four-step kernel when n_fft == 1024, else the fp32 SIMT kernels), "tc" or "simt"."""
        self.impl = impl
        params = dict(DEFAULT_MEL_PARAMS)
        params.update(mel_params)
        if "win_len" in params and "win_length" not in mel_params:
            params["win_length"] = params.pop("win_len")
        params.pop("win_len", None)
        self.params = params
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pitchextractor_b200.LogMel runs on CUDA (sm_100) only; there is no CPU path")
        self.tables = logmel_tables(self.device, **params)
        self._ws = None

    def num_frames(self, num_samples):
        return 1 + num_samples // self.params["hop_length"]

    def __call__(self, wave, crop=None, T_out=0, layout="bmt", lengths=None):
        """lengths: int32 [B] valid samples per item of a zero-padded mixed-length batch (None = all L): each item is
        reflect-padded at its own end and its frames t >= 1 + length // hop are 0.0, exactly what the reference's
        per-item mel + Collater give (meldataset.py:644,804-816)."""
        if wave.dim() == 1:
            wave = wave[None]
        wave = wave.to(self.device, torch.float32).contiguous()
        B, Lw = wave.shape
        T = self.num_frames(Lw)
        To = T_out if T_out > 0 else T
        n_mels = self.params["n_mels"]
        shape = (B, n_mels, To) if layout == "bmt" else (B, To, n_mels)
        out = torch.empty(shape, device=self.device, dtype=torch.float32)
        if crop is not None:
            crop = crop.to(self.device, torch.int32).contiguous()
        if lengths is not None:
            lengths = lengths.to(self.device, torch.int32).contiguous()
        use_tc = self.tables["tc"] if self.impl == "auto" else self.impl == "tc"
        if use_tc and not self.tables["tc"]:
            raise RuntimeError("the tcgen05 log-mel kernel needs n_fft == 1024, hop % 4 == 0, hop <= 320, n_mels <= 128")
        if use_tc:
            need = B * ((Lw + 3) // 4 * 4)
        else:
            need = B * T * (self.params["n_fft"] // 2 + 1)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, device=self.device, dtype=torch.float32)
        kw = dict(out_bmt=out if layout == "bmt" else None, out_btm=out if layout == "btm" else None, crop=crop, T_out=To,
                  lengths=lengths)
        if use_tc:  # writes every output row itself (zeros past the end of an item: Collater padding)
            ops.logmel_tc(wave, self.tables, ws=self._ws, **kw)
        else:
            ops.logmel(wave, self.tables, power_ws=self._ws, **kw)
        return out
