#!/usr/bin/env bash
# Round-2 first GPU call (1 GPU): parity suite after the ADVICE fixes, fp64 peaks / library comparators, sanitizer passes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest1.log | cut -c1-300
timeout 600 python tools/fp64_peaks.py gpurun_out/r02_fp64_peaks.json > gpurun_out/r2_fp64.log 2>&1; echo "fp64 rc=$?"; tail -2 gpurun_out/r2_fp64.log | cut -c1-2500
for tool in memcheck racecheck; do
  : > gpurun_out/r02_sanitizer_${tool}.txt
  for c in batch fused flush lu small rank1; do
    timeout 200 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_cases.py $c > gpurun_out/san_tmp.txt 2>&1; rc=$?
    echo "== $tool $c rc=$rc" >> gpurun_out/r02_sanitizer_${tool}.txt
    grep -v "^$" gpurun_out/san_tmp.txt | tail -12 | cut -c1-300 >> gpurun_out/r02_sanitizer_${tool}.txt
    echo "$tool $c rc=$rc"
  done
done
