#!/usr/bin/env bash
mkdir -p gpurun_out
CMD="python bench.py --workload batch_small_lps_tiny --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_k6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_batch_primal -s 1 -c 1 -o gpurun_out/prof_k6_v2 $CMD > gpurun_out/ncu_k6_v2.log 2>&1
echo "ncu k6 rc=$?"; cut -c1-300 gpurun_out/plain_k6.log | tail -1
timeout 600 python bench.py --workload batch_small_lps_65536x64x128 --steps 3 --no-cpu > gpurun_out/bench_batch_final.json 2> gpurun_out/bench_batch_final.err; echo "batch rc=$?"; cut -c1-400 gpurun_out/bench_batch_final.json
