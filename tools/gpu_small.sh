#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or netlib or afiro or dual or max_iter or refactorisation or steepest or invert" > gpurun_out/pytest_small.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_small.log | cut -c1-900
for w in netlib_afiro netlib_adlittle netlib_blend; do
  timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/bench_$w.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
for k,v in d["solvers"].items(): print("$w",k,{a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ("wall_ms_median","device_ms_median","pivots","launches_per_solve","cpu_ms_median","cpu_pivots")})
PY
done
