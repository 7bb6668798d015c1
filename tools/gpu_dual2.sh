#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "dual or golden or netlib or afiro or refactorisation or steepest" > gpurun_out/pytest_dual2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_dual2.log | cut -c1-700
for w in dense_revised_dual_4096x12288 dense_revised_dual_dse_4096x12288; do
timeout 600 python bench.py --workload $w --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/bench_$w.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
print("$w value",round(d["value"],1),"pivots/step",d["config"]["pivots_per_step"],"dev ms/step",round(d["device_ms_per_step"],2),"k3 ms",round(d["roofline"]["ms_per_launch"],4),"frac",round(d["roofline"]["frac"],3),"e2e",d["e2e"] and round(d["e2e"]["value"],1))
PY
done
