#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -12 gpurun_out/pytest_gpu3.log
timeout 900 python bench.py --workload dense_tableau_16384x32768 > gpurun_out/bench_16k.json 2> gpurun_out/bench_16k.err; echo "16k rc=$?"; tail -3 gpurun_out/bench_16k.err
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_16k.json","gpurun_out/bench_default.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1),"ms/step",round(d["ms_per_step"],2),"dev",round(d["device_ms_per_step"],2),"roof",round(d["roofline"]["achieved"],1),round(d["roofline"]["frac"],3),"k3ms",round(d["roofline"]["ms_per_launch"],4),"share",round(d["roofline"]["share_of_step_device_time"],3),"e2e",d["e2e"] and round(d["e2e"]["value"],1),"cpu",d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "launches", d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
