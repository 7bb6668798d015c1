#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_blk5.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_blk5.log | cut -c1-300
python - <<'PY' > gpurun_out/blk_sweep4.jsonl
import sys, json
sys.path.insert(0, "tools"); sys.argv = ["x"]
import blk_sweep as B
from ellp_b200 import _native as N
ctx = N.Context(0)
for (m, ns) in ((32768, 32768), (16384, 16384)):
    for bk in (0, 16, 24, 32, 40, 48):
        print(json.dumps(B.loop_point(ctx, m, ns, bk, 240 if bk else 40)), flush=True)
PY
cat gpurun_out/blk_sweep4.jsonl | cut -c1-330
