#!/usr/bin/env bash
mkdir -p gpurun_out
for w in dense_revised_dual_4096x12288 dense_revised_dual_dse_4096x12288; do
CMD="python bench.py --workload $w --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$w.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 1500 --csv --log-file gpurun_out/launches_$w.csv $CMD > gpurun_out/ncu_list_$w.log 2>&1
echo "$w ncu rc=$?"; cut -c1-300 gpurun_out/plain_$w.log | tail -1
done
