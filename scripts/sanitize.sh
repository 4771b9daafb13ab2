#!/bin/bash
# compute-sanitizer over every kernel of the path on a 96x64x32 rig. Run on the GPU box:
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
# Summaries land in gpurun_out/sanitize_<tool>.log; copy them to profiles/ to commit.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck initcheck synccheck; do
    extra=""
    [ "$tool" = memcheck ] && extra="--leak-check no"
    timeout 900 $CS --tool $tool $extra --print-limit 20 --error-exitcode 9 python scripts/sanitize_rig.py --bands > gpurun_out/sanitize_$tool.log 2>&1
    echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
    tail -n 4 gpurun_out/sanitize_$tool.log
done
