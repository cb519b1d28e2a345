"""One warm + one measured launch of a single stage kernel at its bench size, for `ncu -k regex:<kernel> --launch-skip 1 -c 1`
(tools/bench_stages.py launches every stage several times; under `--set full` that is minutes of replays).

  python tools/prof_kernels.py mfcc      # k_mfcc, 1 h of synthetic audio resident in HBM (620 153 frames)
  python tools/prof_kernels.py cosine    # k_cosine_scan at config 4 (100 000 x 10 000)
Prints the CUDA-event time of the second launch (never a bench value under a profiler)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from soundsym_b200 import api, synth

what = sys.argv[1] if len(sys.argv) > 1 else "mfcc"
ctx = api.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)


def timed(fn):
    fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn()
    e1.record(stream)
    ctx.sync()
    return e0.elapsed_time(e1)


if what == "mfcc":
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 3600.0
    # (profiling only needs the right shape of data: white noise instead of synth.audio's 6 s of host synthesis)
    audio = np.random.default_rng(1).uniform(-0.5, 0.5, size=int(seconds * 44100))
    n = len(audio)
    frames = (n - 1024) // 256 + 1
    d_audio = torch.from_numpy(audio).cuda()
    d_mfcc = torch.empty((frames, 12), dtype=torch.float64, device="cuda")
    ms = timed(lambda: ctx.check(ctx.lib.ss_mfcc_dev(ctx.h, d_audio.data_ptr(), n, 44100.0, 12, d_mfcc.data_ptr())))
    print("k_mfcc %d frames: %.3f ms" % (frames, ms))
else:
    nd, nq = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (100000, 10000)
    d, doff = synth.segments(nd, 13, seed=1234)
    q, qoff = synth.segments(nq, 13, seed=5678)
    dev = api.DeviceDictionary(ctx, d, doff)
    qs = api.DeviceQueries(ctx, q, qoff)
    oi = torch.empty((nq, 1), dtype=torch.int32, device="cuda")
    od = torch.empty((nq, 1), dtype=torch.float64, device="cuda")
    ms = timed(lambda: ctx.check(ctx.lib.ss_dict_match_dev(dev.h, qs.h, api.SS_COSINE_REF, None, 1, oi.data_ptr(), od.data_ptr())))
    print("cosine-ref %d x %d: %.3f ms (scan %.3f)" % (nd, nq, ms, float(ctx.lib.ss_dict_last_scan_ms(dev.h))))
