#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/flush_shard_sweep.py > gpurun_out/r02_flush_shard_sweep.jsonl 2> gpurun_out/sweep.err; echo "sweep rc=$?"; python - <<'PY'
import json
for l in open("gpurun_out/r02_flush_shard_sweep.jsonl"):
    d=json.loads(l); print(d["C"], d["k"], d["heuristic"], d["col_steps"], d["ms"], d["TFLOPs"])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "residual" 2>&1 | tail -3
