"""Small end-to-end pass over EVERY kernel of libsoundsym_b200.so, sized for compute-sanitizer:

    compute-sanitizer --tool {memcheck,racecheck,synccheck,initcheck} python tools/sanitize_driver.py

No torch, no oracle: ctypes + numpy only, so that the sanitizer instruments this library's kernels and nothing else.
Every stage prints the kernels' launch count; tools/sanitize.sh collects the four reports under profiles/."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from soundsym_b200 import api, synth  # noqa: E402
from soundsym_b200._lib import SS_COSINE_REF, SS_DTW  # noqa: E402


def main():
    ctx = api.Context(0)
    rng = np.random.default_rng(0)
    # ---- Sound: decode, MFCC, max power, mean; batch; chunk-free ingest ------------------------------------------------
    audio = synth.audio(0.6, seed=3)
    pcm16 = (audio * 32767).astype(np.int16)
    samples = ctx.decode_pcm(pcm16.astype(np.int32), 16)
    m, mp, mean = ctx.analyze(samples, 44100.0, 12)
    s2, m2, mp2, mean2 = ctx.analyze_pcm(pcm16, 16, 44100.0, 12)
    assert np.array_equal(m, m2) and mp == mp2
    off = np.array([0, 3000, 3000, 9000, len(samples)], dtype=np.uint64)
    ctx.analyze_batch(samples, off, 44100.0, 12)
    ctx.max_power_batch(samples, off)
    ctx.analyze(samples[:5000], 44100.0, 13)
    big = np.tile(pcm16, 160)[: (4 << 20) + 70000]  # > 4 Mi samples: the chunked ingest (copy of chunk k+1 over compute of chunk k)
    ctx.analyze_pcm(big, 16, 44100.0, 12)
    print("sound ok", ctx.launches, flush=True)
    # ---- Partitioner: GMM training, symbols, Voting Experts ------------------------------------------------------------
    long_m, _, _ = ctx.analyze(synth.audio(4.0, seed=5), 44100.0, 12)
    model = None
    for seed in range(8):
        try:
            model = ctx.gmm_train(long_m, 26, 5, 0.1, seed)
            break
        except Exception:
            continue
    assert model is not None
    sym = ctx.symbols(long_m, model)
    ctx.vote_split(sym, 3, 4)
    splits = ctx.partition(long_m, model, 4, 3)
    print("segment ok", ctx.launches, len(splits), flush=True)
    # ---- matcher: cosine-ref, tensor-core DTW (single + paired tiles), fp32 DTW with strips, exhaustive stage, merge -----
    d, doff = synth.segments(700, 13, seed=1)
    q, qoff = synth.segments(150, 13, seed=2)
    dev = api.DeviceDictionary(ctx, d, doff)
    for k in (1, 4):
        dev.match(q, qoff, SS_DTW, k)
    assert dev.last_uncertified == 0
    dev.match(q, qoff, SS_COSINE_REF, 1)
    dev.match(q, qoff, SS_COSINE_REF, 1, np.linspace(-1e-4, 1e-4, 150))
    dev.debug_tc_scan(q[: int(qoff[20])], qoff[:21])
    lens = np.array([1, 0, 33, 64, 65, 200, 2, 31, 32], dtype=np.uint64)
    loff = np.zeros(len(lens) + 1, dtype=np.uint64)
    loff[1:] = np.cumsum(lens)
    ld = rng.normal(size=(int(loff[-1]), 13)) * 5
    long_dev = api.DeviceDictionary(ctx, ld, loff)  # > 32 frames: fp32 scan, strips, thread-per-pair rescore
    long_dev.match(ld, loff, SS_DTW, 8)
    seg = d[int(doff[3]):int(doff[4])]
    copies = [seg + rng.normal(size=seg.shape) * 1e-4 for _ in range(40)]
    dd = np.concatenate([d[: int(doff[64])]] + copies)
    ddoff = np.concatenate([doff[:65], doff[64] + np.uint64(len(seg)) * np.arange(1, 41, dtype=np.uint64)]).astype(np.uint64)
    dup = api.DeviceDictionary(ctx, dd, ddoff)  # near-duplicates: tensor-core -> fp32 -> exhaustive f64 stage
    dup.match(seg, np.array([0, len(seg)], dtype=np.uint64), SS_DTW, 4)
    assert dup.last_exhaustive >= 1
    print("match ok", ctx.launches, flush=True)
    # ---- resynthesis + sequence distances -------------------------------------------------------------------------------
    ds = rng.normal(size=5000)
    dso = np.array([0, 1000, 2500, 5000], dtype=np.uint64)
    ctx.resynth(ds, dso, np.array([2, 0, 1, 1], dtype=np.uint32), np.array([700, 1800, 0, 1500], dtype=np.uint64))
    ctx.sequence_distances(rng.normal(size=(6, 12)))
    print("all ok", ctx.launches, flush=True)


if __name__ == "__main__":
    main()
