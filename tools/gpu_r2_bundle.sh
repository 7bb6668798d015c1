#!/usr/bin/env bash
# Round 2, 1 GPU bundle
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest9.log | head -20; grep -n "^E  " gpurun_out/r2_pytest9.log | head -10 | cut -c1-600
timeout 900 python tools/full_solve_stats.py > gpurun_out/r02_full_solve_stats.jsonl 2> gpurun_out/full_solve.err; echo "full solve rc=$?"; cut -c1-330 gpurun_out/r02_full_solve_stats.jsonl; tail -3 gpurun_out/full_solve.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_default_g1.json 2> gpurun_out/r02_bench_default_g1.err; echo "bench default rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_default_g1.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d.get("objective_after_timed_steps")); print(json.dumps(d.get("other_configs"))[:1500])
PY
for wl in dense_tableau_dual_4096x12288 dense_tableau_4096x12288 dense_tableau_dual_32768x65536 batch_small_lps_65536x64x128; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r02_bench_${wl}_g1.json 2> gpurun_out/r02_bench_${wl}_g1.err; echo "bench $wl rc=$? $(cut -c1-110 gpurun_out/r02_bench_${wl}_g1.json)"
done
python tools/phase_timing.py > gpurun_out/r02_phase_timing.jsonl 2>&1; grep '"coop_threads": 256' gpurun_out/r02_phase_timing.jsonl | cut -c1-400
