#!/usr/bin/env bash
mkdir -p gpurun_out
CMD="python tools/k4_one.py 8192"
$CMD > gpurun_out/plain_k4c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_k4c_8192.csv $CMD > gpurun_out/ncu_list_k4c.log 2>&1
echo "ncu list rc=$?"; cat gpurun_out/plain_k4c.log
ncu --set full --clock-control none --import-source on -k regex:k_blk_flush3 -s 140 -c 1 -o gpurun_out/prof_k4c_flush3 $CMD > gpurun_out/ncu_k4c_flush.log 2>&1; echo "ncu flush rc=$?"
