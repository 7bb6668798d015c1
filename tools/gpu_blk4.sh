#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rankk or blocked" > gpurun_out/pytest_blk4.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_blk4.log | cut -c1-300
python - <<'PY' > gpurun_out/blk_sweep3.jsonl
import sys, json
sys.path.insert(0, "tools"); sys.argv = ["x"]
import blk_sweep as B
from ellp_b200 import _native as N
ctx = N.Context(0)
for k in (8, 16, 24, 32, 40, 48, 56, 64):
    for cs in (8, 32):
        print(json.dumps(B.flush_point(ctx, 32768, 65536, k, cs)), flush=True)
PY
cat gpurun_out/blk_sweep3.jsonl | cut -c1-250
