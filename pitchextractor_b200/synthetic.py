"""Seeded synthetic training segments with exact F0 labels (host side, numpy).

Restates the *distribution* of the reference's WORLD-vocoder vowel generator (Utils/synthetic.py:23-48,122-147,
155-191,203-218 with the ``world_vocoder`` block of Configs/config.yml:169-180): random piecewise-linear F0 curve of at
most 4 segments in [110, 320] Hz, vibrato with probability 0.5 (0.4 semitones, 4-6 Hz), three Gaussian-formant vowel
envelopes, gain U(-18, -6) dB and additive N(0, -60 dB) noise.  ``pyworld.synthesize`` itself is replaced by additive
harmonic synthesis (phase = cumulative sum of the per-sample F0, harmonic amplitudes sampled from the formant envelope
up to Nyquist); 10-30 % of the frames are made unvoiced (F0 = 0, harmonics muted) so the voicing loss sees both classes.
"""
import numpy as np

VOWELS = (  # (centre Hz, bandwidth Hz, amplitude) -- Utils/synthetic.py:23-48
    ((730.0, 90.0, 1.0), (1090.0, 110.0, 0.6), (2440.0, 150.0, 0.4)),
    ((390.0, 80.0, 1.0), (1990.0, 120.0, 0.6), (2550.0, 160.0, 0.4)),
    ((440.0, 70.0, 1.0), (1020.0, 90.0, 0.6), (2240.0, 150.0, 0.4)),
)
SEGMENT_SAMPLES = 58624  # ceil((192*300 + 1024) / 24000 * 24000), meldataset.py:191-195


def _envelope(freqs, formants):
    env = np.zeros_like(freqs)
    for f, bw, a in formants:
        env += a * np.exp(-0.5 * ((freqs - f) / (bw / 2.0)) ** 2)
    return np.maximum(env, 1e-3)


def f0_curve(rng, n_frames, frame_period_s, pitch=(110.0, 320.0), max_segments=4, vib_p=0.5, vib_semitones=0.4,
             vib_rate=(4.0, 6.0)):
    curve = np.full(n_frames, rng.uniform(*pitch))
    nseg = int(rng.integers(1, max_segments + 1))
    if nseg > 1 and n_frames > 2:
        cuts = np.sort(rng.choice(np.arange(1, n_frames - 1), size=min(nseg - 1, n_frames - 2), replace=False))
        pos = np.concatenate(([0], cuts, [n_frames - 1]))
        vals = rng.uniform(pitch[0], pitch[1], size=len(pos))
        curve = np.interp(np.arange(n_frames), pos, vals)
    if rng.random() < vib_p and vib_semitones > 0:
        t = np.arange(n_frames) * frame_period_s
        curve = curve * 2.0 ** (np.sin(2.0 * np.pi * rng.uniform(*vib_rate) * t) * vib_semitones / 12.0)
    return curve


def make_segment(rng, num_samples=SEGMENT_SAMPLES, sr=24000, hop=300, unvoiced=(0.1, 0.3), gain_db=(-18.0, -6.0),
                 noise_db=-60.0):
    """-> (wave float32 [num_samples], f0 float32 [1 + num_samples // hop] in Hz with 0 = unvoiced)."""
    n_frames = 1 + num_samples // hop
    f0 = f0_curve(rng, n_frames, hop / sr)
    voiced = np.ones(n_frames, bool)
    n_unv = int(round(rng.uniform(*unvoiced) * n_frames))
    if n_unv > 0:
        start = int(rng.integers(0, n_frames - n_unv + 1))
        voiced[start:start + n_unv] = False
    t_frames = np.arange(n_frames) * hop
    t = np.arange(num_samples)
    f0_s = np.interp(t, t_frames, f0)
    gate = np.interp(t, t_frames, voiced.astype(np.float64))
    phase = 2.0 * np.pi * np.cumsum(f0_s) / sr
    formants = VOWELS[int(rng.integers(0, len(VOWELS)))]
    wave = np.zeros(num_samples)
    # harmonics up to 5 kHz (the formant envelopes sit below 2.6 kHz and are at their 1e-3 floor beyond)
    n_harm = int(min(sr / 2, 5000.0) // f0.max())
    env_frames = np.stack([_envelope(h * f0, formants) for h in range(1, n_harm + 1)])  # [H, frames]
    for h in range(1, n_harm + 1):
        wave += np.interp(t, t_frames, env_frames[h - 1]) * np.sin(h * phase)
    wave *= gate / max(np.abs(wave).max(), 1e-9)
    wave *= 10.0 ** (rng.uniform(*gain_db) / 20.0)
    wave += rng.normal(scale=10.0 ** (noise_db / 20.0), size=num_samples)
    return wave.astype(np.float32), (f0 * voiced).astype(np.float32)


def make_batch(batch_size, seed, num_samples=SEGMENT_SAMPLES, sr=24000, hop=300):
    """-> waves [B, num_samples] f32, f0 [B, T] f32 (Hz, 0 = unvoiced); counter-based seeding (seed, item)."""
    waves = np.empty((batch_size, num_samples), np.float32)
    f0s = np.empty((batch_size, 1 + num_samples // hop), np.float32)
    for i in range(batch_size):
        waves[i], f0s[i] = make_segment(np.random.default_rng([seed, i]), num_samples, sr, hop)
    return waves, f0s
