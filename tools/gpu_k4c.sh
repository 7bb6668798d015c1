#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lu or invert or refactor or dual or golden or rankk" > gpurun_out/pytest_k4c.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_k4c.log | cut -c1-800
timeout 600 python - <<'PY' > gpurun_out/k4c_bench.jsonl 2>&1
import ctypes as C, json, sys
sys.path.insert(0, ".")
from ellp_b200 import _native as N
ctx = N.Context(0)
for panel in (0, 1):
    ctx.set_tuning("refactor_panel", panel)
    for m in (1024, 4096, 8192, 16384):
        if panel == 1 and m > 8192: continue
        ms = C.c_float()
        ctx.check(N.lib.ellp_b200_refactor_bench(ctx.h, m, 1, 2, 2, C.byref(ms)))
        flops = (2.0 / 3 + 2.0) * m ** 3
        print(json.dumps({"m": m, "panel": "coop" if panel == 0 else "single_cta", "ms": round(ms.value, 3), "TFLOPs_equiv": round(flops / (ms.value * 1e-3) / 1e12, 3)}), flush=True)
PY
cat gpurun_out/k4c_bench.jsonl
