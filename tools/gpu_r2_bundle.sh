#!/usr/bin/env bash
# Round 2, 1 GPU bundle: full suite, smoke, every 1-GPU bench line of the round, ncu launch list + full capture of the dual kernel.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest7.log | head -20; grep -n "^E  " gpurun_out/r2_pytest7.log | head -10 | cut -c1-600
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
for wl in dense_tableau_32768x65536 dense_tableau_16384x32768 dense_tableau_4096x12288 dense_tableau_dual_4096x12288 dense_tableau_dual_devex_4096x12288 dense_tableau_dual_32768x65536 dense_revised_dual_dse_4096x12288 batch_small_lps_65536x64x128 netlib_afiro; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r02_bench_${wl}_g1.json 2> gpurun_out/r02_bench_${wl}_g1.err; echo "bench $wl rc=$? $(cut -c1-110 gpurun_out/r02_bench_${wl}_g1.json)"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo "ref arm rc=$?"
# ncu: launch list of the dual bench command, then one full capture of the dual pivot kernel (only after the plain run exited 0)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_ncu_launches_dual_tableau_4096x12288.csv python bench.py --workload dense_tableau_dual_4096x12288 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_dual_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_blk_dual_pivots_fused -c 1 -o gpurun_out/r02_ncu_full_dual_pivots python bench.py --workload dense_tableau_dual_4096x12288 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_dual_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
