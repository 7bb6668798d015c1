#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu7.log | cut -c1-250
for w in dense_revised_dual_4096x12288 dense_revised_dual_dse_4096x12288; do
timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/bench_$w.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
    print("$w value",round(d["value"],1),"pivots/step",d["config"]["pivots_per_step"],"dev ms/step",round(d["device_ms_per_step"],2),"k3 ms",round(d["roofline"]["ms_per_launch"],4),"share",round(d["roofline"]["share_of_step_device_time"],3),"e2e",d["e2e"] and round(d["e2e"]["value"],1),"cpu",d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2))
except Exception as e: print("ERR", e)
PY
done
