#!/bin/bash
# compute-sanitizer over every kernel of the library (tools/sanitize_driver.py), one tool at a time; summaries land in
# gpurun_out/sanitizer_<tool>.log (copy them to profiles/ to keep them). Run on the GPU box:  bash tools/sanitize.sh
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
rc=0
for tool in memcheck racecheck synccheck initcheck; do
    extra=""
    [ "$tool" = memcheck ] && extra="--leak-check full"
    timeout 900 $CS --tool $tool $extra --print-limit 50 --error-exitcode 9 python tools/sanitize_driver.py > "$OUT/sanitizer_$tool.log" 2>&1
    r=$?
    echo "$tool exit $r: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY' "$OUT/sanitizer_$tool.log" | tr '\n' ' ')"
    [ $r -ne 0 ] && rc=$r
done
exit $rc
