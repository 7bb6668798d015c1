#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "rankk or blocked" > gpurun_out/pytest_flush3b.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_flush3b.log | cut -c1-600
timeout 600 python - <<'PY' > gpurun_out/flush3b_sweep.jsonl
import sys, json
sys.path.insert(0, "tools"); sys.argv = ["x"]
import blk_sweep as B
from ellp_b200 import _native as N
ctx = N.Context(0)
for fk in (3, 2):
    ctx.set_tuning("flush_kernel", fk)
    for k in (24, 32, 48, 64):
        d = B.flush_point(ctx, 32768, 32768, k, 32); d["flush_kernel"] = fk
        print(json.dumps(d), flush=True)
ctx.set_tuning("flush_kernel", 3)
for (m, ns) in ((32768, 32768), (16384, 16384)):
    for bk in (48, 64):
        d = B.loop_point(ctx, m, ns, bk, 960 if m < 32768 else 640, 32); d["flush_kernel"] = 3
        print(json.dumps(d), flush=True)
PY
cut -c1-300 gpurun_out/flush3b_sweep.jsonl
