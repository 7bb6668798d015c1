"""One blocked-LU refactorisation (K4) of a random dense m x m basis resident in HBM: the command profiled by ncu."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ellp_b200 import _native as N
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = N.Context(0)
ms = C.c_float()
ctx.check(N.lib.ellp_b200_refactor_bench(ctx.h, m, 1, 2, 1, C.byref(ms)))
print(json.dumps({"m": m, "ms": round(ms.value, 3), "TFLOPs_equiv": round((2.0 / 3 + 2.0) * m ** 3 / (ms.value * 1e-3) / 1e12, 3)}))
