#!/usr/bin/env bash
# Round 2, last 1-GPU validation: full parity suite, smoke, the driver's bench command, the K3b kernel sweeps of the final build, batch /
# dual bench lines.  gpurun -- 'bash tools/gpu_r2_final2.sh [nosuite]'
mkdir -p gpurun_out
if [[ "$1" != "nosuite" ]]; then
  timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_final2.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_final2.log | head -20; grep -n "^E  " gpurun_out/r2_pytest_final2.log | head -10 | cut -c1-400
  timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke2.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke2.log | cut -c1-160
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_default_g1.json 2> gpurun_out/r02_bench_default_g1.err; echo "bench default rc=$?"; cut -c1-200 gpurun_out/r02_bench_default_g1.json
timeout 300 python bench.py --workload batch_small_lps_65536x64x128 --steps 5 --warmup 3 > gpurun_out/r02_bench_batch_small_lps_65536x64x128_g1.json 2>/dev/null; echo "batch rc=$?"; cut -c1-160 gpurun_out/r02_bench_batch_small_lps_65536x64x128_g1.json
timeout 300 python bench.py --workload dense_tableau_dual_4096x12288 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_dense_tableau_dual_4096x12288_g1.json 2>/dev/null; echo "dual rc=$?"; cut -c1-160 gpurun_out/r02_bench_dense_tableau_dual_4096x12288_g1.json
timeout 300 python tools/flush5_sweep.py > gpurun_out/r02_flush5_sweep.jsonl 2> gpurun_out/r02_flush5_sweep.err; echo "sweep rc=$?"
timeout 300 python tools/flush_lowk_sweep.py > gpurun_out/r02_flush_lowk_sweep.jsonl 2> gpurun_out/r02_flush_lowk_sweep.err; echo "lowk sweep rc=$?"
python -c "
import json
for f in ('gpurun_out/r02_flush5_sweep.jsonl', 'gpurun_out/r02_flush_lowk_sweep.jsonl'):
    for l in open(f):
        d = json.loads(l); print({k: d[k] for k in d if k in ('R', 'C', 'k', 'ms', 'TFLOPs', 'GBs', 'flush_kernel', 'block_k', 'pivots_per_s')})
"
# ncu evidence of the default command with the final kernels (each only after the plain run above exited 0)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_default_k64_32768x65536.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-other-configs > gpurun_out/ncu_default_launches2.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_blk_flush4r -s 2 -c 1 -o gpurun_out/r02_ncu_full_blk_flush4r python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-other-configs > gpurun_out/ncu_flush4r_full.log 2>&1; echo "ncu full rc=$?"
