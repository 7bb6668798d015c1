#!/usr/bin/env bash
# tableau engine tests + first bench lines + ncu launch list / full capture of K3
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tableau or generated" > gpurun_out/pytest_tableau.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tableau.log
tail -15 gpurun_out/pytest_tableau.log
timeout 300 python bench.py --workload dense_tableau_tiny --steps 2 --warmup 1 > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; echo "tiny rc=$?"; tail -3 gpurun_out/bench_tiny.err
timeout 900 python bench.py --workload dense_tableau_16384x32768 > gpurun_out/bench_16k.json 2> gpurun_out/bench_16k.err; echo "16k rc=$?"; tail -3 gpurun_out/bench_16k.err
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"; tail -3 gpurun_out/bench_default.err
CMD="python bench.py --workload dense_tableau_16384x32768 --steps 1 --warmup 1 --pivots 4 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_16k.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_rank1 -s 4 -c 3 -o gpurun_out/prof_k3_16k $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
cat gpurun_out/bench_tiny.json gpurun_out/bench_16k.json gpurun_out/bench_default.json
