"""Times the basis refactorisation (K4): blocked LU + DMMA vs Gauss-Jordan, dense random basis resident in HBM."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ellp_b200 import _native as N  # noqa: E402

ctx = N.Context(0)
sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192]
for m in sizes:
    for mode, name in ((2, "blocked_lu_dmma"), (1, "gauss_jordan")):
        if mode == 1 and m > 4096:
            continue
        ms = C.c_float()
        ctx.check(N.lib.ellp_b200_refactor_bench(ctx.h, m, 1, mode, 2, C.byref(ms)))
        flops = (2.0 / 3 + 2.0) * m ** 3
        print(json.dumps({"m": m, "mode": name, "ms": round(ms.value, 3), "TFLOPs_equiv": round(flops / (ms.value * 1e-3) / 1e12, 3),
                          "flops_model": "2/3 m^3 LU + m^3 forward on I + m^3 backward"}), flush=True)
