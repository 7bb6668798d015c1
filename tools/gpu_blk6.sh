#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rankk or blocked or tableau" > gpurun_out/pytest_blk6.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_blk6.log | cut -c1-400
python - <<'PY' > gpurun_out/blk_sweep5.jsonl
import sys, json
sys.path.insert(0, "tools"); sys.argv = ["x"]
import blk_sweep as B
from ellp_b200 import _native as N
ctx = N.Context(0)
for (m, ns) in ((32768, 32768), (16384, 16384), (4096, 8192)):
    for bk in (16, 24, 32, 40, 48):
        print(json.dumps(B.loop_point(ctx, m, ns, bk, 480)), flush=True)
PY
cat gpurun_out/blk_sweep5.jsonl | cut -c1-330
