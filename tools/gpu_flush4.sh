#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "rankk or blocked or lu or invert" > gpurun_out/pytest_flush4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_flush4.log | cut -c1-600
timeout 600 python - <<'PY' > gpurun_out/flush4_sweep.jsonl
import sys, json
sys.path.insert(0, "tools"); sys.argv = ["x"]
import blk_sweep as B
from ellp_b200 import _native as N
ctx = N.Context(0)
for fk in (4, 3):
    ctx.set_tuning("flush_kernel", fk)
    for k in (32, 40, 48, 56, 64):
        for cs in ((16, 32) if fk == 4 else (32,)):
            d = B.flush_point(ctx, 32768, 32768, k, cs); d["flush_kernel"] = fk
            print(json.dumps(d), flush=True)
ctx.set_tuning("flush_kernel", 0)
for (m, ns) in ((32768, 32768), (16384, 16384), (4096, 8192)):
    for bk in (48, 64):
        d = B.loop_point(ctx, m, ns, bk, 960 if m < 32768 else 640, 32); d["flush_kernel"] = 0
        print(json.dumps(d), flush=True)
PY
cut -c1-300 gpurun_out/flush4_sweep.jsonl
