#!/usr/bin/env bash
mkdir -p gpurun_out
CMD="python bench.py --workload batch_small_lps_tiny --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_batch_primal -s 1 -c 1 -o gpurun_out/prof_k6 $CMD > gpurun_out/ncu_k6.log 2>&1
tail -3 gpurun_out/ncu_k6.log
