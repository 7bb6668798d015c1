#!/usr/bin/env bash
# Round 2: dual simplex on the blocked tableau -- parity tests, then bench lines (1 GPU).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dual_tableau or dual_pivot_sequence_on_the_tableau or generated_dual_lp_on_the_tableau" > gpurun_out/r2_dual_tests.log 2>&1; echo "dual tests rc=$?"; tail -25 gpurun_out/r2_dual_tests.log | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest2.log | cut -c1-300
for wl in dense_tableau_dual_4096x12288 dense_revised_dual_4096x12288; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r02_bench_${wl}_g1.json 2> gpurun_out/r02_bench_${wl}_g1.err; echo "bench $wl rc=$?"; cut -c1-900 gpurun_out/r02_bench_${wl}_g1.json
done
for bk in 16 32 64; do
  timeout 300 python bench.py --workload dense_tableau_dual_4096x12288 --block-k $bk --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_bench_dual4k_bk${bk}.json 2>/dev/null; echo "bk=$bk rc=$?"; cut -c1-200 gpurun_out/r02_bench_dual4k_bk${bk}.json
done
