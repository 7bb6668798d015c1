#!/usr/bin/env bash
# final 1-GPU validation of round 2 + ncu evidence of the default bench command
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_final.log | head -20; grep -n "^E  " gpurun_out/r2_pytest_final.log | head -10 | cut -c1-600
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r2_smoke.log | cut -c1-160
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_default_g1.json 2> gpurun_out/r02_bench_default_g1.err; echo "bench default rc=$?"; cut -c1-200 gpurun_out/r02_bench_default_g1.json
timeout 600 python bench.py --workload batch_small_lps_65536x64x128 --steps 5 --warmup 3 > gpurun_out/r02_bench_batch_small_lps_65536x64x128_g1.json 2>/dev/null; echo "batch rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_default_32768x65536.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-other-configs > gpurun_out/ncu_default_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_blk_flush4 -s 2 -c 1 -o gpurun_out/r02_ncu_full_blk_flush4 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-other-configs > gpurun_out/ncu_flush_full.log 2>&1; echo "ncu full rc=$?"
