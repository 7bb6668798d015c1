#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "blocked or tableau or sharded or full_size or condensed_fast" > gpurun_out/pytest_fused3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_fused3.log | cut -c1-1000
timeout 300 python - <<'PY' > gpurun_out/phase_timing4.jsonl 2>&1
import sys, json
sys.path.insert(0, "tools"); sys.argv = ["x"]
import phase_timing as PT
from ellp_b200 import _native as N
ctx = N.Context(0)
for thr in (256, 512):
    ctx.set_tuning("coop_threads", thr)
    for (m, ns, bk) in ((1024, 2048, 32), (4096, 8192, 48), (32768, 32768, 64)):
        d = PT.run(ctx, m, ns, bk, 256); d["coop_threads"] = thr
        print(json.dumps(d), flush=True)
PY
cut -c1-440 gpurun_out/phase_timing4.jsonl
