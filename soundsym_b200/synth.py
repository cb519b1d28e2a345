"""Seeded synthetic inputs of the shapes BASELINE.json's configs name.

Host-side numpy only; used by bench.py and the tests to build dictionaries / queries / audio. No GPU code here.
"""
import numpy as np


def segments(nseg, ncoeffs=13, lmin=4, lmax=32, seed=1234):
    """Ragged MFCC-like segments: lengths ~ U{lmin..lmax}; features AR(1) along time,
    x_t = 0.9 x_{t-1} + eps, eps_k ~ N(0, s_k^2), s_k = 20/(1+k), plus a per-segment offset ~ N(0, s_k^2).
    Returns (flat f64 [total_frames, ncoeffs], offsets u64 [nseg+1] in FRAMES)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lmin, lmax + 1, size=nseg)
    off = np.zeros(nseg + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    sig = 20.0 / (1.0 + np.arange(ncoeffs))
    eps = rng.normal(size=(total, ncoeffs)) * sig
    x = np.empty((total, ncoeffs), dtype=np.float64)
    # AR(1) restarted at every segment start, vectorised over segments by stepping through time positions
    starts = off[:-1].astype(np.int64)
    prev = np.zeros((nseg, ncoeffs))
    for t in range(int(lens.max())):
        alive = lens > t
        idx = starts[alive] + t
        cur = 0.9 * prev[alive] + eps[idx]
        x[idx] = cur
        prev[alive] = cur
    offs = rng.normal(size=(nseg, ncoeffs)) * sig
    x += np.repeat(offs, lens, axis=0)
    return x, off


def audio(seconds, sr=44100, seed=42):
    """Mono f64 audio in [-1, 1]: three sinusoids whose frequencies (100-8000 Hz, log-uniform) and amplitudes re-draw
    every U(50, 300) ms, plus white noise at -40 dBFS (no all-zero frames)."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    out = rng.normal(size=n) * 0.01
    pos = 0
    phase = np.zeros(3)
    while pos < n:
        ln = min(n - pos, int(rng.uniform(0.05, 0.3) * sr))
        freqs = np.exp(rng.uniform(np.log(100.0), np.log(8000.0), size=3))
        amps = rng.uniform(0.05, 0.3, size=3)
        t = np.arange(ln)
        for k in range(3):
            w = 2 * np.pi * freqs[k] / sr
            out[pos:pos + ln] += amps[k] * np.sin(phase[k] + w * t)
            phase[k] = (phase[k] + w * ln) % (2 * np.pi)
        pos += ln
    return np.clip(out, -1.0, 1.0)
