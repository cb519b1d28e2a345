#!/bin/bash
# The ncu-free part of tools/measure_head.sh (tests, bench lines, stage timings, latency, one per-CTA timeline of a 1/8 shard)
# on one GPU at the current commit. Writes into gpurun_out/head/.
set -u
O=gpurun_out/head
mkdir -p $O
(time timeout 600 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py > $O/r2_bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/r2_bench_n1.json; tail -3 $O/bench_n1.err
python bench.py --config 3 > $O/r2_config3.json 2> $O/config3.err
python bench.py --config 5 > $O/r2_config5_dtw_n1.json 2> $O/config5.err
python bench.py --config 5 --mode cosine > $O/r2_config5_cosine_n1.json 2> $O/config5c.err
python tools/bench_stages.py --out $O/r2_stage_timings.json > $O/stages.log 2>&1
python tools/bench_latency.py --out $O/r2_latency.json > $O/latency.log 2>&1
SS_DTW_H2_TIMELINE=$O/tl_12500.txt python tools/dtw_sweep.py 12500 10000 > $O/sweep_12500_tl.log 2>&1
python tools/h2_timeline.py $O/tl_12500.txt > $O/r2_h2_timeline_eighth_shard.txt 2>&1; tail -5 $O/r2_h2_timeline_eighth_shard.txt
python tools/dtw_sweep.py 12500 10000 2>&1 | grep scan_ms
ls -la $O | tail -20
