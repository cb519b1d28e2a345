#!/usr/bin/env python
"""bench.py — headline benchmark of the soundsym hot path on B200 (contract in the build brief, section 4).

Metric (BASELINE.json): DTW cell-updates/s and queries/s, 100k-segment dictionary, on 1/2/4/8 B200.
Workload (config 4 of BASELINE.json): synthetic dictionary of 100 000 segments x 10 000 query segments,
13-coefficient MFCC frames, lengths ~ U{4..32}; a STEP = one match of the whole query batch against the whole dictionary
(top-1, SS_DTW). With N GPUs the dictionary is partitioned into N contiguous shards (balanced by frames), every rank
sees all queries, per-rank top-k are all-gathered over NCCL and merged on every rank  ->  "scaling": "strong".

  value      cells/s with the queries' f64 MFCCs already in HBM: layout kernel + tensor-core scan + merge + f64 refine
             (+ all-gather + merge for N > 1), CUDA events on the library's stream, max over ranks.
  e2e        the same through the reference-facing call ss_dict_match with pinned HOST buffers: H2D of the queries and
             D2H of the results inside the timed region, every step.
  roofline   the dominant kernel (k_dtw_scan_tc), timed live with CUDA events around its launch (ss_dict_last_scan_ms);
             algorithmic bytes = sum over pairs of (Lq + Ld) * C * 4 B, peak = MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference   the f64 CPU oracle ("port": the Rust reference cannot be built here) on a bounded
             sample of the same workload, on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ND, NQ, C, K = 100000, 10000, 13, 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nd", type=int, default=ND)
    ap.add_argument("--nq", type=int, default=NQ)
    ap.add_argument("--cpu-sample-queries", type=int, default=0, help="queries in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """samples SM clock, power and throttle reasons DURING the timed region through NVML (the same counters as the
    `nvidia-smi --query-gpu=clocks.sm,...,clocks_event_reasons.*` line of B200_PROFILING.md). An in-process NVML thread
    is used instead of an `nvidia-smi -lms` child because starting nvidia-smi next to a ~1 s timed region stalled the
    driver and inflated the measured step by 30-60 % (measured; see DESIGN.md "Measurement")."""

    def __init__(self, gpu_index, period_s=0.01):
        self.gpu, self.period = gpu_index, period_s
        self.rows, self.on, self.t, self.h = [], False, None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                try:
                    phys = int(vis.split(",")[gpu_index])
                except Exception:
                    phys = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def _sample(self):
        nv = self.nv
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
            try:
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.rows.append((sm, pw, rs))
        except Exception:
            pass

    def _loop(self):
        # the timed region of an 8-GPU run is ~40 ms: sample every 10 ms while it is on
        while self.alive:
            if self.on:
                self._sample()
            time.sleep(self.period)

    def prepare(self):
        """start the thread (idle) before warm-up so that no start-up cost lands in the timed region"""
        if self.h is None:
            return
        self.alive = True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def start(self):
        self.on = True

    def stop(self):
        if self.h is not None and self.on and not self.rows:
            self._sample()  # region shorter than one period: one reading while the last step is still in flight
        self.on = False
        self.alive = False
        if self.h is None or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, bit in names.items() if any(r[2] & bit for r in self.rows))
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(self.max_sm), "reasons": reasons,
                "power_w_max": max(r[1] for r in self.rows), "samples": len(self.rows)}


def workload(nd, nq):
    from soundsym_b200 import synth
    d, doff = synth.segments(nd, C, seed=1234)
    q, qoff = synth.segments(nq, C, seed=5678)
    return d, doff, q, qoff


def shard_bounds(doff, n):
    """contiguous segment ranges balanced by frames."""
    total = int(doff[-1])
    cuts = [0]
    for r in range(1, n):
        cuts.append(int(np.searchsorted(doff, total * r // n, side="left")))
    cuts.append(len(doff) - 1)
    return cuts


def algorithmic_bytes(doff, qoff, s0, s1):
    """sum over (query, dict segment in [s0, s1)) of (Lq + Ld) * C * 4 B."""
    ld = int(doff[s1] - doff[s0])
    lq = int(qoff[-1])
    nq, nd = len(qoff) - 1, s1 - s0
    return (nd * lq + nq * ld) * C * 4


def cpu_sample(d, doff, q, qoff, nsample, threads):
    from oracle import oracle as O
    O.set_threads(threads)
    sub_off = qoff[: nsample + 1]
    sub = q[: int(sub_off[-1])]
    t0 = time.perf_counter()
    O.dtw_topk(d, doff, sub, sub_off, C, K)
    dt = time.perf_counter() - t0
    cells = int(doff[-1]) * int(sub_off[-1])
    return cells / dt, dt, cells


def run_reference(args):
    """the reference arm: the CPU path (oracle port; the Rust crate cannot be built in this image) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    d, doff, q, qoff = workload(args.nd, args.nq)
    threads = O.hardware_threads()
    nsample = args.cpu_sample_queries or max(threads, 32)
    # calibrate so one step is a few seconds
    rate, dt, _ = cpu_sample(d, doff, q, qoff, min(nsample, 8), threads)
    for _ in range(args.warmup):
        pass  # CPU path has no warm-up state beyond the calibration pass above
    times, cells = [], 0
    for _ in range(args.steps):
        r, dt, cells = cpu_sample(d, doff, q, qoff, nsample, threads)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = cells / (ms * 1e-3)
    sample = "%d of %d queries x full %d-segment dictionary per step (%.3e cells), f64 oracle port, %d threads" % (
        nsample, args.nq, args.nd, cells, threads)
    line = {"impl": "reference", "metric": "dtw_cell_updates_per_s", "value": value, "unit": "cells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic 100k-segment dictionary x 10k queries, C=13, L~U{4..32}, DTW top-1 (config 4)",
                       "nd": args.nd, "nq": args.nq, "ncoeffs": C, "k": K},
            "queries_per_s": nsample / (ms * 1e-3),
            "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from soundsym_b200 import api
    from soundsym_b200._lib import SS_DTW

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    n_gpus = world

    d, doff, q, qoff = workload(args.nd, args.nq)
    nq = len(qoff) - 1
    cuts = shard_bounds(doff, world)
    s0, s1 = cuts[rank], cuts[rank + 1]
    ctx = api.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    shard = api.DeviceDictionary(ctx, d, doff[s0:s1 + 1], C, index_base=s0)
    total_cells = int(doff[-1]) * int(qoff[-1])

    # device-resident queries + outputs
    qdev = api.DeviceQueries(ctx, q, qoff, C)
    o_idx = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    o_dist = torch.empty((nq, K), dtype=torch.float64, device="cuda")
    if world > 1:
        g_idx = torch.empty((world, nq, K), dtype=torch.int32, device="cuda")
        g_dist = torch.empty((world, nq, K), dtype=torch.float64, device="cuda")
        m_idx = torch.empty((nq, K), dtype=torch.int32, device="cuda")
        m_dist = torch.empty((nq, K), dtype=torch.float64, device="cuda")
    # pinned host buffers for the e2e leg
    hq = torch.from_numpy(q).pin_memory()
    hqoff = torch.from_numpy(qoff.astype(np.int64)).pin_memory()
    h_idx = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    h_dist = torch.empty((nq, K), dtype=torch.float64).pin_memory()
    h_gidx = torch.empty((world, nq, K), dtype=torch.int32).pin_memory() if world > 1 else None

    def exchange():
        """all-gather of the per-shard top-k over NCCL + lexicographic merge on every rank."""
        dist.all_gather_into_tensor(g_idx, o_idx)
        dist.all_gather_into_tensor(g_dist, o_dist)
        ctx.check(ctx.lib.ss_topk_merge_dev(ctx.h, g_idx.data_ptr(), g_dist.data_ptr(), world, nq, K, m_idx.data_ptr(), m_dist.data_ptr()))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step_resident():
        flush.zero_()
        ctx.check(ctx.lib.ss_queries_invalidate(qdev.h))  # layout kernels run inside the step
        ctx.check(ctx.lib.ss_dict_match_dev(shard.h, qdev.h, SS_DTW, None, K, o_idx.data_ptr(), o_dist.data_ptr()))
        if world > 1:
            exchange()

    def step_e2e():
        flush.zero_()
        # the reference-facing call: HOST buffers in, HOST buffers out (H2D + D2H inside)
        ctx.check(ctx.lib.ss_dict_match(shard.h, hq.data_ptr(), hqoff.data_ptr(), nq, SS_DTW, None, K, h_idx.data_ptr(), h_dist.data_ptr()))
        if world > 1:
            o_idx.copy_(h_idx, non_blocking=True)
            o_dist.copy_(h_dist, non_blocking=True)
            exchange()
            h_idx.copy_(m_idx, non_blocking=True)
            h_dist.copy_(m_dist, non_blocking=True)
            stream.synchronize()

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        return ms / steps, wall / steps

    with torch.cuda.stream(stream):
        sampler = ClockSampler(local)
        if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
            sampler.prepare()
        for _ in range(max(args.warmup, 3)):
            step_resident()
        sampler.start()
        launches0 = ctx.launches
        ms_step, _ = timed(step_resident, args.steps)
        launches = ctx.launches - launches0
        scan_ms = float(ctx.lib.ss_dict_last_scan_ms(shard.h))
        tc_fallback = shard.last_tc_fallback
        clocks = sampler.stop() if rank == 0 else None
        uncert = shard.last_uncertified
        for _ in range(2):
            step_e2e()
        # e2e blocks on the host every step (the call returns results in host memory): wall clock == device time here
        _, ms_e2e = timed(step_e2e, args.steps)

    # correctness guard on a few queries against the oracle (outside every timed region)
    ok = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.set_threads(O.hardware_threads())
        res_idx = (m_idx if world > 1 else o_idx).cpu().numpy().astype(np.uint32)
        res_dist = (m_dist if world > 1 else o_dist).cpu().numpy()
        probe = 4
        oi, od = O.dtw_topk(d, doff, q[: int(qoff[probe])], qoff[: probe + 1], C, K)
        ok = bool(np.array_equal(oi, res_idx[:probe]) and np.allclose(od, res_dist[:probe], rtol=1e-12, atol=0))
        assert np.array_equal(res_idx, h_idx.numpy().astype(np.uint32)), "e2e and resident paths disagree"

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        alg = algorithmic_bytes(doff, qoff, s0, s1)
        achieved = alg / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else None
        scan_kernel = "k_dtw_scan_tc (tcgen05 fp16 cost + CUDA-core DP)" if os.environ.get("SS_DTW_TC", "1") != "0" else "k_dtw_scan (fp32)"
        traffic = None
        tp = os.path.join(ROOT, "profiles", "dtw_scan_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("tc" if "tc" in scan_kernel else "fp32", {}).get("dram_bytes_per_launch") if world == 1 else None
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import oracle as O
            threads = O.hardware_threads()
            nsample = args.cpu_sample_queries or max(4 * threads, 64)
            rate, dt, cells = cpu_sample(d, doff, q, qoff, nsample, threads)
            cpu = {"value": rate, "unit": "cells/s", "cores": threads, "kind": "port",
                   "sample": "%d of %d queries x full %d-segment dictionary (%.3e cells, %.1f s), f64 oracle port" % (nsample, nq, args.nd, cells, dt)}
        line = {
            "metric": "dtw_cell_updates_per_s", "value": total_cells / (ms_step * 1e-3), "unit": "cells/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 (DP scan; local costs from f16 tensor-core products accumulated in f32; winners refined in f64)", "data": "synthetic",
            "config": {"workload": "synthetic 100k-segment dictionary x 10k queries, C=13, L~U{4..32}, DTW top-1 (config 4)",
                       "nd": args.nd, "nq": args.nq, "ncoeffs": C, "k": K, "parallelism": "dictionary sharded x%d, NCCL all-gather top-k merge" % world,
                       "l2": "flushed: a 256 MB buffer is overwritten at the start of every timed step (dictionary resident: %.0f MB fp32 stream + %.0f MB f64)" % (
                           int(doff[s1] - doff[s0]) * 64 / 1e6, int(doff[s1] - doff[s0]) * C * 8 / 1e6)},
            "queries_per_s": nq / (ms_step * 1e-3),
            "e2e": {"value": total_cells / (ms_e2e * 1e-3), "unit": "cells/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(q.nbytes + qoff.nbytes), "d2h_bytes_per_step": int(nq * K * 12),
                    "queries_per_s": nq / (ms_e2e * 1e-3)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": scan_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": scan_ms, "algorithmic_bytes_per_launch": alg,
                         "note": "EFFECTIVE bandwidth of the pairwise-streaming model of SURVEY.md §8d (bytes the CPU path streams per pair); "
                                 "operands stay on chip (one dictionary tile serves 128 queries), so compulsory DRAM traffic is ~0.1 GB and "
                                 "frac can exceed 1. The kernel is bound by CUDA-core issue of the DP recurrence (DESIGN.md §3.1), not by HBM "
                                 "or the tensor pipe",
                         "issue_bound": {"cells_per_clk_per_sm": (int(doff[s1] - doff[s0]) * int(qoff[-1])) / (scan_ms * 1e-3) / 148.0
                                         / (((clocks or {}).get("sm_mhz") or 1965.0) * 1e6) if scan_ms > 0 else None,
                                         "ceiling_cells_per_clk_per_sm": 45.3,
                                         "ceiling_note": "the register-resident band alone (FMNMX3 + FADD per cell, no TMEM, no barriers) measured on "
                                                         "B200 with 16 warps / SM: tools/microbench_band.cu, profiles/r1_microbench_band.log"},
                         "cells_per_s_kernel": (int(doff[s1] - doff[s0]) * int(qoff[-1])) / (scan_ms * 1e-3) if scan_ms > 0 else None},
            "cpu_baseline": cpu,
            "uncertified_queries": int(uncert),
            "tc_fallback_queries": int(tc_fallback),
            "oracle_probe_ok": ok,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
