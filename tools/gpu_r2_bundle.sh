#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "harris or starts_optimal or small_cases" > gpurun_out/r2_pytest12.log 2>&1; echo "pytest subset rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest12.log | head -20; grep -n "^E  " gpurun_out/r2_pytest12.log | head -12 | cut -c1-500
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_default_g1.json 2> gpurun_out/r02_bench_default_g1.err; echo "bench default rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_default_g1.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["roofline"]["frac"]); print(json.dumps(d["other_configs"]["configs[2]"])[:700])
PY
