#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rankk or blocked or tableau or sharded_engines" > gpurun_out/pytest_flush3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_flush3.log | cut -c1-600
timeout 600 python - <<'PY' > gpurun_out/flush3_sweep.jsonl
import sys, json
sys.path.insert(0, "tools"); sys.argv = ["x"]
import blk_sweep as B
from ellp_b200 import _native as N
ctx = N.Context(0)
for fk, css in ((3, (16, 32, 64)),):
    ctx.set_tuning("flush_kernel", fk)
    for k in (16, 24, 32, 40, 48, 56, 64):
        for cs in css:
            d = B.flush_point(ctx, 32768, 32768, k, cs); d["flush_kernel"] = fk
            print(json.dumps(d), flush=True)
for cs in (32,):
    for (m, ns) in ((32768, 32768), (16384, 16384), (4096, 8192)):
        for bk in (32, 48, 64):
            d = B.loop_point(ctx, m, ns, bk, 960 if m < 32768 else 640, cs); d["flush_kernel"] = 3
            print(json.dumps(d), flush=True)
PY
cut -c1-300 gpurun_out/flush3_sweep.jsonl
