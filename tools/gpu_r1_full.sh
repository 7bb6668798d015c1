#!/usr/bin/env bash
# Round-1 consolidated GPU run (1 GPU): full parity suite, smoke, bench lines, ncu launch list + full captures.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_full.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log | cut -c1-200
for w in dense_tableau_32768x65536 dense_tableau_16384x32768 dense_tableau_4096x12288; do
  timeout 900 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/bench_$w.err; cut -c1-1500 gpurun_out/bench_$w.json
done
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-600 gpurun_out/bench_reference.json
CMD="python bench.py --steps 1 --warmup 1 --pivots 112 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_blk.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_blk56_32k.csv $CMD > gpurun_out/ncu_list_blk.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_blk_flush4 -s 2 -c 1 -o gpurun_out/prof_blk_flush4_k56 $CMD > gpurun_out/ncu_flush.log 2>&1; echo "ncu flush rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_blk_pivots_fused -s 2 -c 1 -o gpurun_out/prof_blk_pivots_fused_k56 $CMD > gpurun_out/ncu_pivots.log 2>&1; echo "ncu pivots rc=$?"
ls -la gpurun_out/*.ncu-rep
