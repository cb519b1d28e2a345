"""Latency of the nq = 1 path: SoundDictionary::at_distance / match_sound issue ONE query per call
(examples/matcher.rs:39-48, SoundSequence::from_distances src/sound.rs:405-417). Wall clock of ss_dict_match(nq = 1) with host
buffers against the 100k-segment dictionary of config 4 (and the 10k one of config 3), both matchers; also nq = 8 and 128.
  python tools/bench_latency.py [--out profiles/bench/latency.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soundsym_b200 import api, synth  # noqa: E402
from soundsym_b200._lib import SS_COSINE_REF, SS_DTW  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--reps", type=int, default=50)
args = ap.parse_args()
ctx = api.Context(0)
res = {}
q, qoff = synth.segments(256, 13, seed=5678)
for nd in (10000, 100000):
    d, doff = synth.segments(nd, 13, seed=1234)
    dev = api.DeviceDictionary(ctx, d, doff)
    for mode, name in ((SS_DTW, "dtw"), (SS_COSINE_REF, "cosine_ref")):
        for nq in (1, 8, 128):
            subs = []
            for r in range(args.reps + 5):
                i0 = (r * nq) % (256 - nq + 1)
                off = (qoff[i0:i0 + nq + 1] - qoff[i0]).astype(np.uint64)
                subs.append((np.ascontiguousarray(q[int(qoff[i0]):int(qoff[i0 + nq])]), off))
            for s, o in subs[:5]:
                dev.match(s, o, mode, 1)
            ts = []
            for s, o in subs[5:]:
                t0 = time.perf_counter()
                dev.match(s, o, mode, 1)
                ts.append((time.perf_counter() - t0) * 1e6)
            ts = np.array(ts)
            res["%s_nd%d_nq%d" % (name, nd, nq)] = {"median_us": float(np.median(ts)), "p10_us": float(np.percentile(ts, 10)), "p90_us": float(np.percentile(ts, 90)),
                                                    "calls_per_s": float(1e6 / np.median(ts))}
            print("%-10s nd=%6d nq=%3d: median %8.1f us  (p10 %8.1f, p90 %8.1f)" % (name, nd, nq, np.median(ts), np.percentile(ts, 10), np.percentile(ts, 90)), flush=True)
    dev.close()
if args.out:
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
