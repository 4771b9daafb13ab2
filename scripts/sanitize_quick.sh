set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in synccheck racecheck; do
    timeout 900 $CS --tool $tool --print-limit 3 --error-exitcode 9 python scripts/sanitize_rig.py --bands > gpurun_out/sanitize_$tool.log 2>&1
    echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
    grep -E "SUMMARY|Barrier error|Race reported" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -8
done
