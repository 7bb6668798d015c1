#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "condensed_fast or tableau_engine or netlib or max_iter" > gpurun_out/pytest_e2e.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_e2e.log | cut -c1-800
for w in dense_tableau_32768x65536 dense_tableau_16384x32768; do
  timeout 900 python bench.py --workload $w --no-cpu > gpurun_out/bench_e2e_$w.json 2> gpurun_out/bench_e2e_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/bench_e2e_$w.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_e2e_$w.json").read().strip().splitlines()[-1])
print("$w value",round(d["value"]),"e2e",d["e2e"],"clk",d["clocks"])
PY
done
