#!/usr/bin/env bash
# final 1-GPU validation of round 2: full suite, smoke, default bench line, AFIRO
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_final.log | head -20; grep -n "^E  " gpurun_out/r2_pytest_final.log | head -10 | cut -c1-600
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r2_smoke.log | cut -c1-160
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_default_g1.json 2> gpurun_out/r02_bench_default_g1.err; echo "bench default rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_default_g1.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d.get("objective_after_timed_steps"), d["e2e"]["value"]); print(json.dumps(d.get("other_configs"))[:1200])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo "ref arm rc=$?"
timeout 300 python bench.py --workload netlib_afiro --steps 10 --warmup 3 > gpurun_out/r02_bench_netlib_afiro_g1.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_netlib_afiro_g1.json
