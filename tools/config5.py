#!/usr/bin/env python
"""Config 5 of BASELINE.json: examples/reconstruction.rs on synthetic 44.1 kHz audio (default 1 h: source = first half,
target = second half) on N B200s of one box.

  python tools/config5.py [--seconds S] [--mode cosine|dtw]                                    # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/config5.py ...

Stages (examples/reconstruction.rs:43-86): Sound::from_samples(source) -> Partitioner(threshold 4, depth 3).train ->
partition -> SoundDictionary::from_segments -> Sound::from_samples(target) -> partition with the source's model ->
clone_from_dictionary (one nearest-match query per target segment) -> to_sound.

What shards: the matcher. Every rank runs the cheap front end (MFCC ~7 ms, training ~30 ms, partition ~2 ms
per half hour of audio) on its own GPU - replicating it costs less than one broadcast of the MFCC matrix - then holds
one contiguous shard of the dictionary (balanced by frames), matches ALL target segments against it, and the per-rank
winners are all-gathered over NCCL and merged lexicographically by (distance, index) on every rank, so the result equals
the 1-GPU run bit for bit. Rank 0 resynthesises. Prints one JSON line with per-stage device times (max over ranks)."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3600.0)
    ap.add_argument("--mode", default="cosine", choices=["cosine", "dtw"])
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()

    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("CONFIG5_WATCHDOG_S", "600")), exit=True)  # a hung collective must not hold the box
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
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    mode = SS_DTW if args.mode == "dtw" else SS_COSINE_REF
    C = api.NCOEFFS

    t0 = time.perf_counter()
    audio = synth.audio(args.seconds, seed=args.seed)
    half = (len(audio) // 2 // 256) * 256
    gen_s = time.perf_counter() - t0
    times = {}

    def stage(name, fn):
        ctx.sync()
        t = time.perf_counter()
        r = fn()
        ctx.sync()
        times[name] = (time.perf_counter() - t) * 1e3
        return r

    # ---- front end, replicated on every rank ----------------------------------------------------------------------
    src = stage("analyze_source_ms", lambda: api.Sound.from_samples(audio[:half], 44100.0, ctx=ctx))
    part = api.Partitioner(src, ctx).set_threshold(4).set_depth(3)  # examples/reconstruction.rs:43-45
    stage("train_ms", lambda: part.train(seed=3))
    splits = stage("partition_source_ms", part.partition)
    tgt = stage("analyze_target_ms", lambda: api.Sound.from_samples(audio[half:], 44100.0, ctx=ctx))
    part.sound = tgt
    tsplits = stage("partition_target_ms", part.partition)

    # SoundDictionary::add_segments (src/sound.rs:330-343): segment i owns (seg_i / HOP) MFCC rows, in order
    doff = np.zeros(len(splits) + 1, dtype=np.uint64)
    doff[1:] = np.cumsum(np.asarray(splits, dtype=np.uint64) // np.uint64(256))
    qoff = np.zeros(len(tsplits) + 1, dtype=np.uint64)
    qoff[1:] = np.cumsum(np.asarray(tsplits, dtype=np.uint64) // np.uint64(256))
    dm, qm = src.mfcc_arrays()[: int(doff[-1])], tgt.mfcc_arrays()[: int(qoff[-1])]
    nq = len(tsplits)

    # ---- dictionary shard of this rank ------------------------------------------------------------------------------
    total = int(doff[-1])
    cuts = [0] + [int(np.searchsorted(doff, total * r // world, side="left")) for r in range(1, world)] + [len(splits)]
    s0, s1 = cuts[rank], cuts[rank + 1]
    shard = stage("dictionary_build_ms", lambda: api.DeviceDictionary(ctx, dm[int(doff[s0]): int(doff[s1])], doff[s0:s1 + 1] - doff[s0], C,
                                                                     index_base=s0))
    qdev = api.DeviceQueries(ctx, qm, qoff, C)
    o_idx = torch.empty((nq, 1), dtype=torch.int32, device="cuda")
    o_dist = torch.empty((nq, 1), dtype=torch.float64, device="cuda")
    if world > 1:
        g_idx = torch.empty((world, nq, 1), dtype=torch.int32, device="cuda")
        g_dist = torch.empty((world, nq, 1), dtype=torch.float64, device="cuda")
        m_idx = torch.empty((nq, 1), dtype=torch.int32, device="cuda")
        m_dist = torch.empty((nq, 1), dtype=torch.float64, device="cuda")

    def match():
        ctx.check(ctx.lib.ss_dict_match_dev(shard.h, qdev.h, mode, None, 1, o_idx.data_ptr(), o_dist.data_ptr()))
        if world > 1:
            dist.all_gather_into_tensor(g_idx, o_idx)
            dist.all_gather_into_tensor(g_dist, o_dist)
            ctx.check(ctx.lib.ss_topk_merge_dev(ctx.h, g_idx.data_ptr(), g_dist.data_ptr(), world, nq, 1, m_idx.data_ptr(), m_dist.data_ptr()))

    with torch.cuda.stream(stream):
        match()  # warm-up (layout kernels, NCCL channels)
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        match()
        e1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        match_ms = e0.elapsed_time(e1)
    idx = (m_idx if world > 1 else o_idx).cpu().numpy()[:, 0].astype(np.uint32)
    dst = (m_dist if world > 1 else o_dist).cpu().numpy()[:, 0]

    # ---- resynthesis on rank 0: clone_from_dictionary + to_sound ---------------------------------------------------
    out_sum, out_frames = None, None
    if rank == 0:
        soff = np.zeros(len(splits) + 1, dtype=np.uint64)
        soff[1:] = np.cumsum(np.asarray(splits, dtype=np.uint64))
        tlens = np.asarray(tsplits, dtype=np.uint64)
        out = stage("resynth_ms", lambda: ctx.resynth(src.samples()[: int(soff[-1])], soff, idx, tlens))
        res = stage("analyze_result_ms", lambda: api.Sound.from_samples(out, 44100.0, ctx=ctx))
        out_sum, out_frames = float(out.sum()), res.num_frames()

    keys = ["analyze_source_ms", "train_ms", "partition_source_ms", "analyze_target_ms", "partition_target_ms", "dictionary_build_ms",
            "resynth_ms", "analyze_result_ms"]  # the same list on every rank (the last two only run on rank 0)
    vals = [times.get(k, 0.0) for k in keys] + [match_ms]
    if world > 1:
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = t.tolist()
    if rank == 0:
        tm = dict(zip(keys, vals[:-1]))
        pairs = len(splits) * nq
        line = {"config": "config 5: examples/reconstruction.rs on %.0f s synthetic 44.1 kHz audio, %s matcher" % (args.seconds, args.mode),
                "n_gpus": world, "source_frames": src.num_frames(), "target_frames": tgt.num_frames(), "dictionary_segments": len(splits),
                "target_segments": nq, "max_segment_frames": int((doff[1:] - doff[:-1]).max()), "match_ms": vals[-1],
                "pairs_per_s": pairs / (vals[-1] * 1e-3), "queries_per_s": nq / (vals[-1] * 1e-3), "stage_ms": tm,
                "idx_sha1": hashlib.sha1(idx.tobytes()).hexdigest()[:16], "dist_sha1": hashlib.sha1(dst.tobytes()).hexdigest()[:16],
                "out_sum": out_sum, "out_frames": out_frames, "synth_s": gen_s}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
