#!/bin/bash
# Round-end evidence on one B200 (gpurun --timeout 2400 -- 'bash scripts/final_round.sh'): GPU tests, compute-sanitizer (default
# build), the ncu full-set capture of one frame's kernels, the launch list of a short bench run, and the bench line itself.
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final_pytest_gpu.txt
bash scripts/sanitize.sh 2>&1 | grep -E "rc=" | tee gpurun_out/final_sanitize_rc.txt
# second frame of two: launches 9.. are the warm frame (8 kernels per frame)
ncu --set full --clock-control none --import-source on -s 8 -c 8 -f -o gpurun_out/r02_final python scripts/profile_rig.py --rigs 2 > gpurun_out/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/bench_final.err
tail -c 300 gpurun_out/r02_bench_final.json
