#!/usr/bin/env bash
# Round-2 first GPU call (1 GPU): parity suite after the ADVICE fixes, fp64 peaks / library comparators, sanitizer passes.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest1.log | cut -c1-300
timeout 600 python tools/fp64_peaks.py gpurun_out/r02_fp64_peaks.json > gpurun_out/r2_fp64.log 2>&1; echo "fp64 rc=$?"; tail -2 gpurun_out/r2_fp64.log | cut -c1-2500
CASES="batch fused flush lu small rank1"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py $CASES > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02_sanitizer_memcheck.txt | cut -c1-300
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py $CASES > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r02_sanitizer_racecheck.txt | cut -c1-300
