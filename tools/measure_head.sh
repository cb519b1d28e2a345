#!/bin/bash
# Everything DESIGN.md section 6 quotes, on one GPU, at the current commit: tests, bench lines, stage timings, launch list and the
# ncu --set full captures (each after its command has exited 0 without ncu). Writes into gpurun_out/head/.
set -u
O=gpurun_out/head
mkdir -p $O
(time timeout 600 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py > $O/r2_bench_n1.json 2> $O/bench_n1.err; tail -c 400 $O/r2_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/bench_ref.err
python bench.py --config 3 > $O/r2_config3.json 2> $O/config3.err
python bench.py --config 2 > $O/r2_config2.json 2> $O/config2.err
python bench.py --config 5 > $O/r2_config5_dtw_n1.json 2> $O/config5.err
python bench.py --config 5 --mode cosine > $O/r2_config5_cosine_n1.json 2> $O/config5c.err
python tools/bench_stages.py --out $O/r2_stage_timings.json > $O/stages.log 2>&1
python tools/bench_latency.py --out $O/r2_latency.json > $O/latency.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_n1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --probe-queries 8 > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mfcc --launch-skip 1 -c 1 -o $O/prof_mfcc_head -f python tools/prof_kernels.py mfcc > $O/ncu_mfcc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cosine_scan --launch-skip 1 -c 1 -o $O/prof_cos_head -f python tools/prof_kernels.py cosine > $O/ncu_cos.log 2>&1
ncu --set full --clock-control none -k regex:k_dtw_scan_h2 --launch-skip 6 -c 3 -o $O/prof_h2_head -f python tools/dtw_sweep.py 100000 10000 > $O/ncu_h2.log 2>&1
ls -la $O
