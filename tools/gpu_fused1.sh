#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_fused.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_fused.log | cut -c1-400
python - <<'PY' > gpurun_out/fused_sweep.jsonl
import sys, json
sys.path.insert(0, "tools"); sys.argv = ["x"]
import blk_sweep as B
from ellp_b200 import _native as N
ctx = N.Context(0)
for cp in (1, 2):
    ctx.set_tuning("coop_pivots", cp)
    for (m, ns) in ((32768, 32768), (16384, 16384), (4096, 8192), (1024, 2048)):
        for bk in (32, 48):
            d = B.loop_point(ctx, m, ns, bk, 480); d["coop_pivots"] = cp
            print(json.dumps(d), flush=True)
PY
cut -c1-330 gpurun_out/fused_sweep.jsonl
