"""bench.py --config 2 / --config 5: the non-headline configs of BASELINE.json as bench lines (same JSON contract, other workloads).

config 2  examples/partition.rs on tests/Section_7_1.wav (the committed fixture tests/golden/section71.npz holds its PCM and an
          oracle-trained model): Sound::from_path + Partitioner::partition (depth 4, threshold 3), end to end from integer PCM.
          The fixture is 11.5 s of audio (1 978 frames: microseconds of kernel time), so the line also carries the MFCC kernel on
          `--seconds` (default 3600) of synthetic audio resident in HBM, held against BOTH of its roofs: the algorithmic bytes
          (2 144 B / frame) over the measured HBM peak, and ~28 kflop (f64) / frame over the FP64 pipe.
config 5  examples/reconstruction.rs on `--seconds` of synthetic 44.1 kHz audio (source = first half, target = second half) on
          N GPUs: analysis, training, both partitions and the dictionary build are replicated per rank (tens of ms: cheaper than
          broadcasting the MFCC matrix), the matcher runs sharded inside the library (ss_dict_match_sharded), rank 0
          resynthesises. value = target segments matched per second over the whole WARM flow.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP64_PEAK_TFLOPS = 37.0  # B200 FP64 vector peak (non-tensor), NVIDIA datasheet; there is no measured figure in MEASURED_PEAKS.json


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run(args):
    if args.config == 2:
        return run_config2(args)
    return run_config5(args)


def run_config2(args):
    import ctypes as CT
    import torch
    from soundsym_b200 import api, synth
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "section71.npz")))
    ctx = api.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    pcm = np.ascontiguousarray(g["pcm"], dtype=np.int16)
    model = (g["gmm_means"], g["gmm_covs"], g["gmm_weights"])

    def step():
        _, m, _, _ = ctx.analyze_pcm(pcm, 16, 44100.0, 12)
        return m, ctx.partition(m, model, 4, 3)

    for _ in range(max(args.warmup, 3)):
        m, splits = step()
    assert np.array_equal(splits, g["splits_d4t3"]), "partition differs from the golden fixture"
    l0 = ctx.launches
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    launches = ctx.launches - l0
    frames = m.shape[0]

    # the MFCC kernel alone on device-resident audio
    audio = synth.audio(args.seconds, seed=42)
    n = len(audio)
    nfr = (n - 1024) // 256 + 1
    d_audio = torch.from_numpy(audio).cuda()
    d_mfcc = torch.empty((nfr, 12), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(stream):
        for _ in range(3):
            ctx.check(ctx.lib.ss_mfcc_dev(ctx.h, d_audio.data_ptr(), n, 44100.0, 12, d_mfcc.data_ptr()))
        times = []
        for _ in range(max(args.steps, 5)):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.check(ctx.lib.ss_mfcc_dev(ctx.h, d_audio.data_ptr(), n, 44100.0, 12, d_mfcc.data_ptr()))
            e1.record(stream)
            ctx.sync()
            times.append(e0.elapsed_time(e1))
    kms = float(np.mean(times))
    hbm, src = _peaks()
    alg = nfr * 2144
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        O.set_threads(1)
        s16 = O.decode_pcm(pcm.astype(np.int32), 16)
        t0 = time.perf_counter()
        om = O.mfcc(s16)
        O.partition(om, model, 4, 3)
        dt = time.perf_counter() - t0
        cpu = {"value": frames / dt, "unit": "frames/s", "cores": 1, "kind": "port", "sample": "the whole fixture (1 978 frames), MFCC + partition, scalar f64 oracle port, %.2f s" % dt}
        assert np.max(np.abs(om - m)) < 1e-9
    line = {"metric": "partition_frames_per_s", "value": frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "fixture (tests/golden/section71.npz)",
            "config": {"workload": "examples/partition.rs on tests/Section_7_1.wav: Sound::from_path + Partitioner::partition, depth 4, threshold 3 (config 2)",
                       "frames": int(frames), "segments": int(len(splits))},
            "e2e": {"value": frames / (ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(pcm.nbytes + frames * 96), "d2h_bytes_per_step": int(len(pcm) * 8 + frames * 96 + len(splits) * 8),
                    "note": "value IS the end-to-end number here: integer PCM in host memory -> segment lengths in host memory, wall clock"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "kernel": "k_mfcc on %.0f s of synthetic audio resident in HBM (%d frames)" % (args.seconds, nfr), "kernel_ms": kms,
                         "achieved": nfr * 28e3 / (kms * 1e-3) / 1e12, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": nfr * 28e3 / (kms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                         "peak_source": "nominal B200 FP64 vector peak (no measured figure available); ~28 kflop (f64) per frame, SURVEY.md 8d",
                         "hbm": {"achieved_gbs": alg / (kms * 1e-3) / 1e9, "peak_gbs": hbm, "frac": alg / (kms * 1e-3) / 1e9 / hbm, "peak_source": src,
                                 "algorithmic_bytes_per_launch": alg, "note": "2 144 B per frame (256 new f64 samples in, 12 coefficients out)"},
                         "frames_per_s": nfr / (kms * 1e-3), "audio_seconds_per_s": args.seconds / (kms * 1e-3), "traffic": None,
                         "l2": "flushed before every timed launch (256 MB fill, outside the events)"},
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def run_config5(args):
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("CONFIG5_WATCHDOG_S", "900")), exit=True)  # a hung collective must not hold the box
    import torch
    import torch.distributed as dist
    from soundsym_b200 import api, synth
    from soundsym_b200._lib import SS_COSINE_REF, SS_DTW

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    mode = SS_DTW if args.mode == "dtw" else SS_COSINE_REF
    C = api.NCOEFFS
    comm = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(api.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = api.Comm(ctx, world, rank, idt.cpu().numpy().tobytes())

    audio = synth.audio(args.seconds, seed=42)
    half = (len(audio) // 2 // 256) * 256
    pcm = np.clip(np.round(audio * 32767.0), -32768, 32767).astype(np.int16)  # Sound::from_path reads integer PCM

    def flow(times):
        def stage(name, fn):
            ctx.sync()
            t = time.perf_counter()
            r = fn()
            ctx.sync()
            times[name] = (time.perf_counter() - t) * 1e3
            return r
        # Sound::from_path (integer PCM over PCIe, conversion + MFCC + max power + mean on the device)
        def load(p):
            s, m, mp, mean = ctx.analyze_pcm(p, 16, 44100.0, C)
            return api.Sound(s, 44100.0, m, mp, mean, None, ctx)
        src = stage("analyze_source_ms", lambda: load(pcm[:half]))
        part = api.Partitioner(src, ctx).set_threshold(4).set_depth(3)  # examples/reconstruction.rs:43-45
        stage("train_ms", lambda: part.train(seed=3))
        splits = stage("partition_source_ms", part.partition)
        tgt = stage("analyze_target_ms", lambda: load(pcm[half:]))
        part.sound = tgt
        tsplits = stage("partition_target_ms", part.partition)
        doff = np.zeros(len(splits) + 1, dtype=np.uint64)
        doff[1:] = np.cumsum(np.asarray(splits, dtype=np.uint64) // np.uint64(256))
        qoff = np.zeros(len(tsplits) + 1, dtype=np.uint64)
        qoff[1:] = np.cumsum(np.asarray(tsplits, dtype=np.uint64) // np.uint64(256))
        dm, qm = src.mfcc_arrays()[: int(doff[-1])], tgt.mfcc_arrays()[: int(qoff[-1])]
        cuts = api.shard_bounds(doff, world)
        s0, s1 = cuts[rank], cuts[rank + 1]
        shard = stage("dictionary_build_ms", lambda: api.DeviceDictionary(ctx, dm[int(doff[s0]): int(doff[s1])], doff[s0:s1 + 1] - doff[s0], C, index_base=s0))
        if world > 1:
            idx, dst = stage("match_ms", lambda: comm.match(shard, qm, qoff, mode, 1))
        else:
            idx, dst = stage("match_ms", lambda: shard.match(qm, qoff, mode, 1))
        info = {"tc_fallback": shard.last_tc_fallback, "exhaustive": shard.last_exhaustive} if mode == SS_DTW else {}
        out = None
        if rank == 0:
            soff = np.zeros(len(splits) + 1, dtype=np.uint64)
            soff[1:] = np.cumsum(np.asarray(splits, dtype=np.uint64))
            out = stage("resynth_ms", lambda: ctx.resynth(src.samples()[: int(soff[-1])], soff, idx[:, 0], np.asarray(tsplits, dtype=np.uint64)))
            stage("analyze_result_ms", lambda: ctx.analyze(out, 44100.0, C))
        return dict(src_frames=src.num_frames(), tgt_frames=tgt.num_frames(), nd=len(splits), nq=len(tsplits), maxlen=int((doff[1:] - doff[:-1]).max()),
                    idx=idx[:, 0], dst=dst[:, 0], out=out, info=info)

    keys = ["analyze_source_ms", "train_ms", "partition_source_ms", "analyze_target_ms", "partition_target_ms", "dictionary_build_ms", "match_ms",
            "resynth_ms", "analyze_result_ms"]
    cold = {}
    r = flow(cold)  # first pass: cold (module load, workspace allocation, NCCL channels)
    runs = []
    for _ in range(max(args.steps, 2)):
        t = {}
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        r = flow(t)
        t["total_ms"] = (time.perf_counter() - t0) * 1e3
        runs.append(t)
    warm = {k: float(np.median([t.get(k, 0.0) for t in runs])) for k in keys + ["total_ms"]}  # (median: a host hiccup in one run is not the flow)
    vals = [warm[k] for k in keys + ["total_ms"]]
    if world > 1:
        tt = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        vals = tt.tolist()
    if rank == 0:
        tm = dict(zip(keys + ["total_ms"], vals))
        nq = r["nq"]
        line = {"metric": "reconstruction_target_segments_per_s", "value": nq / (tm["total_ms"] * 1e-3), "unit": "segments/s", "n_gpus": world,
                "steps": max(args.steps, 2), "warmup": 1, "ms_per_step": tm["total_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64 (analysis, cosine-ref) / f32+f64 (DTW filter + refine)", "data": "synthetic",
                "config": {"workload": "examples/reconstruction.rs on %.0f s of synthetic 44.1 kHz audio, %s matcher (config 5)" % (args.seconds, args.mode),
                           "seconds": args.seconds, "mode": args.mode},
                "e2e": {"value": nq / (tm["total_ms"] * 1e-3), "unit": "segments/s", "h2d_bytes_per_step": int(pcm.nbytes), "d2h_bytes_per_step": int(len(audio) * 8),
                        "note": "the flow is host-buffer to host-buffer throughout (integer PCM in, resynthesised f64 samples out): value is the e2e number"},
                "stage_ms_warm": tm, "stage_ms_cold_rank0": cold, "match_ms_runs_rank0": [t.get("match_ms") for t in runs], "source_frames": r["src_frames"], "target_frames": r["tgt_frames"],
                "dictionary_segments": r["nd"], "target_segments": nq, "max_segment_frames": r["maxlen"],
                "match_pairs_per_s": r["nd"] * nq / (tm["match_ms"] * 1e-3), "match_info": r["info"],
                "idx_sha1": hashlib.sha1(r["idx"].tobytes()).hexdigest()[:16], "dist_sha1": hashlib.sha1(r["dst"].tobytes()).hexdigest()[:16],
                "out_sum": float(r["out"].sum())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        comm.close()
        dist.destroy_process_group()
