"""Deterministic synthetic workload: clips are sums of 8
log-uniform tones plus Gaussian noise, peak <= 0.9, from default_rng(1000 + i);
watermarks from default_rng(7).  Used by bench.py and the tests (there is no
dataset offline)."""
import numpy as np

N_BITS = 20


def synth_clip(i: int, seconds: float, sr: int) -> np.ndarray:
    rng = np.random.default_rng(1000 + i)
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    f = np.exp(rng.uniform(np.log(100.0), np.log(8000.0), 8))
    f = np.minimum(f, 0.45 * sr)
    a = rng.uniform(0.05, 0.3, 8)
    ph = rng.uniform(0, 2 * np.pi, 8)
    x = (a[:, None] * np.sin(2 * np.pi * f[:, None] * t[None, :] + ph[:, None])).sum(0)
    x = x + 0.02 * rng.standard_normal(n)
    peak = np.max(np.abs(x))
    if peak > 0.9:
        x = x * (0.9 / peak)
    return x.astype(np.float32)


def synth_bits(n_clips: int) -> np.ndarray:
    return np.random.default_rng(7).integers(0, 2, (n_clips, N_BITS), dtype=np.int32)


def synth_batch(n_clips: int, seconds: float, sr: int, unique: int = 16, first: int = 0) -> np.ndarray:
    """Clips [first, first + n_clips) of the global synthetic set, as [n_clips, N].  Generating
    thousands of distinct clips on the host is slow, so only `unique` distinct clips are
    synthesised; the rest are those clips circularly shifted and re-scaled (clip-dependent),
    which keeps every clip's content distinct.  `first` lets a rank build only its own shard."""
    base = np.stack([synth_clip(i, seconds, sr) for i in range(min(unique, first + n_clips))])
    out = np.empty((n_clips, base.shape[1]), dtype=np.float32)
    for j in range(n_clips):
        i = first + j
        b = base[i % len(base)]
        k = i // len(base)
        out[j] = np.roll(b, 997 * k) * np.float32(1.0 - 0.01 * (k % 7)) if k else b
    return out
