"""Per-stage throughput of the non-headline kernels (configs 2 and 5 of BASELINE.json), with the CPU
oracle timed beside them on a bounded sample. Writes one JSON document to stdout / --out.

  MFCC        synthetic 44.1 kHz audio resident in HBM -> ss_mfcc_dev; algorithmic 2 144 B and ~28 kflop (f64) per frame
  analyze     the host-buffer call ss_sound_analyze (H2D of the samples + MFCC + max_power + mean + D2H)
  partition   ss_partition (Standardizer + GMM symbols + Voting Experts) on the MFCC rows, host buffers
  cosine-ref  the reference's matcher on the synthetic 10k x 1k and 100k x 10k dictionaries (device-resident)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402  (baseline timing only)
from soundsym_b200 import api, synth  # noqa: E402
from soundsym_b200._lib import SS_COSINE_REF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=3600.0)
ap.add_argument("--out", default="")
ap.add_argument("--skip-big-cosine", action="store_true")
args = ap.parse_args()

ctx = api.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
res = {"gpu": torch.cuda.get_device_name(0), "host_threads": O.hardware_threads()}


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync()
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
    ctx.sync()
    return e0.elapsed_time(e1) / reps


# ---- MFCC ----------------------------------------------------------------------------------------------------------
t0 = time.time()
audio = synth.audio(args.seconds, seed=42)
res["audio_seconds"] = args.seconds
res["audio_gen_s"] = time.time() - t0
n = len(audio)
frames = O.frame_count(n)
d_audio = torch.from_numpy(audio).cuda()
d_mfcc = torch.empty((frames, 12), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
ms = ev_time(lambda: ctx.check(ctx.lib.ss_mfcc_dev(ctx.h, d_audio.data_ptr(), n, 44100.0, 12, d_mfcc.data_ptr())))
res["mfcc_dev"] = {"frames": frames, "ms": ms, "frames_per_s": frames / (ms * 1e-3), "algorithmic_GBps": frames * 2144 / (ms * 1e-3) / 1e9,
                   "f64_gflops": frames * 28e3 / (ms * 1e-3) / 1e9, "audio_seconds_per_s": args.seconds / (ms * 1e-3)}
# CPU oracle on a bounded sample (single thread, like the reference)
sample_n = min(n, 44100 * 20)
t0 = time.perf_counter()
om = O.mfcc(audio[:sample_n])
dt = time.perf_counter() - t0
res["mfcc_cpu_oracle_1thread"] = {"frames": int(om.shape[0]), "s": dt, "frames_per_s": om.shape[0] / dt, "sample": "first 20 s"}
got = d_mfcc[: om.shape[0]].cpu().numpy()
res["mfcc_max_abs_err_vs_oracle"] = float(np.max(np.abs(got - om)))

# ---- analyze end to end (host buffers, pinned) ------------------------------------------------------------------------
h_audio = torch.from_numpy(audio).pin_memory()
h_mfcc = torch.empty((frames, 12), dtype=torch.float64).pin_memory()
import ctypes as C
fr, mp = C.c_size_t(), C.c_double()
mean = np.empty(12)
reps = 3
ctx.check(ctx.lib.ss_sound_analyze(ctx.h, h_audio.data_ptr(), n, 44100.0, 12, h_mfcc.data_ptr(), C.byref(fr), C.byref(mp), mean.ctypes.data))  # warm-up
t0 = time.perf_counter()
for _ in range(reps):
    ctx.check(ctx.lib.ss_sound_analyze(ctx.h, h_audio.data_ptr(), n, 44100.0, 12, h_mfcc.data_ptr(), C.byref(fr), C.byref(mp), mean.ctypes.data))
dt = (time.perf_counter() - t0) / reps
res["analyze_e2e"] = {"ms": dt * 1e3, "h2d_bytes": int(n * 8), "d2h_bytes": int(frames * 96), "samples_per_s": n / dt,
                      "audio_seconds_per_s": args.seconds / dt}

# ---- PCM ingest: int16 over PCIe, conversion on the device ---------------------------------------------------------------
pcm16 = torch.from_numpy(np.clip(np.round(audio * 32767.0), -32768, 32767).astype(np.int16)).pin_memory()
h_samples = torch.empty(n, dtype=torch.float64).pin_memory()
ctx.check(ctx.lib.ss_sound_analyze_pcm(ctx.h, pcm16.data_ptr(), n, 16, 44100.0, 12, None, h_mfcc.data_ptr(), C.byref(fr), C.byref(mp), mean.ctypes.data))
t0 = time.perf_counter()
for _ in range(reps):
    ctx.check(ctx.lib.ss_sound_analyze_pcm(ctx.h, pcm16.data_ptr(), n, 16, 44100.0, 12, None, h_mfcc.data_ptr(), C.byref(fr), C.byref(mp), mean.ctypes.data))
dt = (time.perf_counter() - t0) / reps
res["analyze_pcm16_e2e"] = {"ms": dt * 1e3, "h2d_bytes": int(n * 2), "audio_seconds_per_s": args.seconds / dt, "note": "samples not copied back (out_samples = NULL)"}

# ---- GMM training on the device -------------------------------------------------------------------------------------------
m_train = np.ascontiguousarray(h_mfcc.numpy()[: frames // 2])
ctx.gmm_train(m_train[:4096], 26, 1, 0.1, 0)
t0 = time.perf_counter()
gm = ctx.gmm_train(m_train, 26, 5, 0.1, 0)
res["gmm_train_e2e"] = {"frames": int(m_train.shape[0]), "ms": (time.perf_counter() - t0) * 1e3, "iters": 5, "ncomp": 26}

# ---- partition ------------------------------------------------------------------------------------------------------
m_host = h_mfcc.numpy()
half = frames // 2
z, _, _ = O.standardize(m_host[: min(half, 20000)])
model = O.gmm_train(z, seed=0)
src = np.ascontiguousarray(m_host[:half])
ctx.partition(src, model, 3, 4)  # warm-up (first call loads the CUB kernels and sizes the workspaces)
t0 = time.perf_counter()
for _ in range(reps):
    lens = ctx.partition(src, model, 3, 4)
dt = (time.perf_counter() - t0) / reps
res["partition_e2e"] = {"frames": int(half), "ms": dt * 1e3, "frames_per_s": half / dt, "segments": int(len(lens)), "depth": 3, "threshold": 4}
sub = src[:40000]
t0 = time.perf_counter()
olens, _, _ = O.partition(sub, model, 3, 4)
dt = time.perf_counter() - t0
res["partition_cpu_oracle_1thread"] = {"frames": 40000, "s": dt, "frames_per_s": 40000 / dt}
glens = ctx.partition(sub, model, 3, 4)
res["partition_equal_to_oracle_on_sample"] = bool(np.array_equal(glens, olens))

# ---- cosine-ref matcher ------------------------------------------------------------------------------------------------
for nd, nq in ((10000, 1000),) + (() if args.skip_big_cosine else ((100000, 10000),)):
    d, doff = synth.segments(nd, 13, seed=1234)
    q, qoff = synth.segments(nq, 13, seed=5678)
    dev = api.DeviceDictionary(ctx, d, doff)
    qs = api.DeviceQueries(ctx, q, qoff)
    oi = torch.empty((nq, 1), dtype=torch.int32, device="cuda")
    od = torch.empty((nq, 1), dtype=torch.float64, device="cuda")
    ms = ev_time(lambda: ctx.check(ctx.lib.ss_dict_match_dev(dev.h, qs.h, SS_COSINE_REF, None, 1, oi.data_ptr(), od.data_ptr())), reps=3)
    work = dev.last_work
    res["cosine_ref_%dx%d" % (nd, nq)] = {"ms": ms, "products": work, "products_per_s": work / (ms * 1e-3), "queries_per_s": nq / (ms * 1e-3),
                                          "scan_ms": float(ctx.lib.ss_dict_last_scan_ms(dev.h))}
    if nd == 10000:
        O.set_threads(1)
        t0 = time.perf_counter()
        ci, cd = O.cosine_match(d, doff, q[: int(qoff[100])], qoff[:101], 13)
        dt = time.perf_counter() - t0
        res["cosine_ref_cpu_oracle_1thread"] = {"queries": 100, "s": dt, "queries_per_s": 100 / dt}
        res["cosine_ref_equal_to_oracle_on_sample"] = bool(np.array_equal(ci, oi[:100, 0].cpu().numpy().astype(np.uint32)))
    dev.close()
    qs.close()

txt = json.dumps(res, indent=1)
print(txt)
if args.out:
    open(args.out, "w").write(txt)
