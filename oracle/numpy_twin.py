"""Independent numpy / pure-Python restatement of the same spec as soundsym_oracle.cpp.

TEST INFRASTRUCTURE ONLY. Its job is to cross-check the C++ oracle: two restatements written separately (different
FFT, different linear algebra, different n-gram bookkeeping) must agree before either is trusted. Citations are to
/root/reference; assumption labels A1..A9 are defined in oracle/ASSUMPTIONS.h.
"""
import math
from collections import Counter, defaultdict

import numpy as np


def read_wav(path):
    """hound-style decode (src/sound.rs:117-120): integer PCM -> f64 / (i32::MAX >> (32 - bits)). Returns
    (samples f64, sample_rate, bits, pcm int32). Mono only (the reference reads interleaved samples as one stream)."""
    import struct

    with open(path, "rb") as f:
        data = f.read()
    assert data[:4] == b"RIFF" and data[8:12] == b"WAVE"
    pos = 12
    fmt = None
    pcm = None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            bits = fmt[3]
            if bits == 16:
                pcm = np.frombuffer(body, dtype="<i2").astype(np.int32)
            elif bits == 24:
                b = np.frombuffer(body[: len(body) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
                v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
                pcm = np.where(v & 0x800000, v - (1 << 24), v).astype(np.int32)
            elif bits == 32:
                pcm = np.frombuffer(body, dtype="<i4").astype(np.int32)
            else:
                raise ValueError("unsupported bits %d" % bits)
        pos += 8 + size + (size & 1)
    denom = float((2 ** 31 - 1) >> (32 - fmt[3]))
    return pcm.astype(np.float64) / denom, float(fmt[2]), fmt[3], pcm


def frame_count(n, bin=1024, hop=256):  # A1
    return (n - bin) // hop + 1 if n >= bin else 0


def hann(bin):  # A2
    i = np.arange(bin, dtype=np.float64)
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * i / (bin - 1)))


def mel_bins(ncoeffs, sr, bin=1024, f_lo=100.0, f_hi=8000.0):  # A3b-c
    lo = 1125.0 * math.log1p(f_lo / 700.0)
    hi = 1125.0 * math.log1p(f_hi / 700.0)
    pts = [(i / ncoeffs) * (hi - lo) + lo for i in range(ncoeffs + 2)]
    return [int(math.floor((bin + 1) * (700.0 * (math.exp(p / 1125.0) - 1.0)) / sr)) for p in pts]


def mfcc(samples, sr=44100.0, ncoeffs=12, bin=1024, hop=256, f_lo=100.0, f_hi=8000.0, floor=1e-10):
    """analyze_mfccs (src/sound.rs:215-242) via numpy's FFT and a dense filterbank / DCT matrix."""
    samples = np.asarray(samples, dtype=np.float64)
    frames = frame_count(samples.shape[0], bin, hop)
    if frames == 0:
        return np.zeros((0, ncoeffs))
    idx = np.arange(bin)[None, :] + hop * np.arange(frames)[:, None]
    x = samples[idx] * hann(bin)[None, :]
    spec = np.fft.fft(x, axis=1)
    power = spec.real ** 2 + spec.imag ** 2
    bins = mel_bins(ncoeffs, sr, bin, f_lo, f_hi)
    fb = np.zeros((ncoeffs, bin))
    for b in range(ncoeffs):
        b0, b1, b2 = bins[b], bins[b + 1], bins[b + 2]
        for i, k in enumerate(range(b0, b1)):
            fb[b, k % bin] += i / (b1 - b0)
        for i, k in enumerate(range(b1, b2)):
            fb[b, k % bin] += 1.0 - i / (b2 - b1)
    e = power @ fb.T
    le = np.log10(np.maximum(e, floor))
    n = np.arange(ncoeffs)
    dctm = 2.0 * np.cos(np.pi * n[:, None] * (2.0 * n[None, :] + 1.0) / (2.0 * ncoeffs))  # [k, n]
    return le @ dctm.T


def max_power(samples):  # src/sound.rs:244-256
    samples = np.asarray(samples, dtype=np.float64)
    frames = frame_count(samples.shape[0], 128, 64)
    if frames == 0:
        return 0.0
    idx = np.arange(128)[None, :] + 64 * np.arange(frames)[:, None]
    rms = np.sqrt((samples[idx] ** 2).sum(axis=1) / 128.0)
    return float(max(0.0, np.nanmax(rms)))


def cosine_sim(me, you):  # src/sound.rs:23-38 (numpy summation order; agrees with the oracle to ~1e-15 relative)
    me, you = np.asarray(me, dtype=np.float64).ravel(), np.asarray(you, dtype=np.float64).ravel()
    n = min(me.shape[0], you.shape[0])
    return float(np.dot(me[:n], you[:n]) / (np.sum(me * me) * np.sum(you * you)))


def dtw(a, b):  # A8
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    la, lb = a.shape[0], b.shape[0]
    if la == 0 or lb == 0:
        return math.inf
    cost = ((a[:, None, :] - b[None, :, :]) ** 2).sum(axis=2)
    D = np.full((la, lb), np.inf)
    for i in range(la):
        for j in range(lb):
            if i == 0 and j == 0:
                m = 0.0
            else:
                m = min(D[i - 1, j] if i else np.inf, D[i, j - 1] if j else np.inf, D[i - 1, j - 1] if i and j else np.inf)
            D[i, j] = cost[i, j] + m
    return float(D[-1, -1] / (la + lb))


def standardize(x):  # A5
    x = np.asarray(x, dtype=np.float64)
    return (x - x.mean(axis=0)) / np.sqrt(x.var(axis=0, ddof=1))


def gmm_posteriors(z, means, covs, weights):  # A6
    z = np.asarray(z, dtype=np.float64)
    out = np.empty((z.shape[0], means.shape[0]))
    for j in range(means.shape[0]):
        P = np.linalg.inv(covs[j])
        d = z - means[j]
        quad = np.einsum("ri,ij,rj->r", d, P, d)
        out[:, j] = weights[j] * np.exp(-0.5 * quad) / math.sqrt(np.linalg.det(covs[j]))
    with np.errstate(all="ignore"):
        return out / out.sum(axis=1, keepdims=True)


def max_index(row):  # src/sound.rs:486-495
    best, idx = 0.0, 0
    for i, v in enumerate(row):
        if v > best:
            best, idx = v, i
    return idx


def cast_votes(sym, depth):  # A7, pure-Python bookkeeping with Counters
    text = bytes(bytearray(int(s) for s in sym))
    n = len(text)
    votes = [0] * (n + 1)
    if depth < 1 or n < depth:
        return np.array(votes, dtype=np.uint32)
    zf, zh = {}, {}
    for ln in range(1, depth + 2):
        cnt = Counter(text[s:s + ln] for s in range(n - ln + 1))
        nxt = defaultdict(Counter)
        for s in range(n - ln):
            nxt[text[s:s + ln]][text[s + ln]] += 1
        keys = sorted(cnt)
        f = np.array([cnt[k] for k in keys], dtype=np.float64)
        h = []
        for k in keys:
            tot = sum(nxt[k].values())
            h.append(-sum((c / tot) * math.log(c / tot) for _, c in sorted(nxt[k].items())) if tot else 0.0)
        h = np.array(h, dtype=np.float64)
        m = float(len(keys))
        s1, s2 = sum(cnt[k] for k in keys), sum(cnt[k] * cnt[k] for k in keys)  # exact integers
        mf = s1 / m
        vf = s2 / m - mf * mf
        sf = math.sqrt(vf) if vf > 0 else 0.0
        for k, v in zip(keys, f):
            zf[k] = (float(v) - mf) / sf if sf > 0 else 0.0
        mh = float(sum(h.tolist())) / m if False else float(np.add.reduce(h)) / m
        sd = math.sqrt(float(((h - mh) ** 2).sum()) / m)
        for k, v in zip(keys, h):
            zh[k] = (float(v) - mh) / sd if sd > 0 else 0.0
    for s in range(n - depth + 1):
        w = text[s:s + depth]
        best_p, best = 1, zh[w[:1]]
        for p in range(2, depth + 1):
            if zh[w[:p]] > best:
                best_p, best = p, zh[w[:p]]
        votes[s + best_p] += 1
        if depth >= 2:
            bp, bv = 1, zf[w[:1]] + zf[w[1:]]
            for p in range(2, depth):
                v = zf[w[:p]] + zf[w[p:]]
                if v > bv:
                    bp, bv = p, v
            votes[s + bp] += 1
    return np.array(votes, dtype=np.uint32)


def split(votes, n, threshold):  # A7
    if n == 0:
        return np.zeros(0, dtype=np.uint64)
    lens, start = [], 0
    for i in range(1, n):
        if votes[i] > votes[i - 1] and votes[i] >= votes[i + 1] and votes[i] >= threshold:
            lens.append(i - start)
            start = i
    lens.append(n - start)
    return np.array(lens, dtype=np.uint64)
