#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload batch_small_lps_tiny --steps 2 --warmup 1 > gpurun_out/bench_batch_tiny.json 2> gpurun_out/bench_batch_tiny.err; echo "tiny rc=$?"; tail -3 gpurun_out/bench_batch_tiny.err
timeout 1200 python bench.py --workload batch_small_lps_65536x64x128 --steps 3 --warmup 1 > gpurun_out/bench_batch.json 2> gpurun_out/bench_batch.err; echo "batch rc=$?"; tail -3 gpurun_out/bench_batch.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_batch_tiny.json","gpurun_out/bench_batch.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"]/1e6,2),"M pivots/s; ms/step",round(d["ms_per_step"],2),"dev",round(d["device_ms_per_step"],2),"lps/s",round(d["config"]["lps_per_s"]),"e2e",d["e2e"] and round(d["e2e"]["value"]/1e6,2),"cpu",d["cpu_baseline"] and round(d["cpu_baseline"]["value"],1), d["roofline"]["note"])
    except Exception as e: print(f, "ERR", e)
PY
