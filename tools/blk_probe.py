"""One short run of the blocked tableau engine (for ncu): python tools/blk_probe.py m ns block_k pivots [runs]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ellp_b200 import _native as N  # noqa: E402

m, ns, bk, piv = [int(a) for a in sys.argv[1:5]]
runs = int(sys.argv[5]) if len(sys.argv) > 5 else 1
ctx = N.Context(0)
o = N.default_opts(piv, engine=N.ENGINE_TABLEAU, block_k=bk, check_every=min(piv, 32), profile=True)
ctx.check(N.lib.ellp_b200_generate_dense(ctx.h, m, ns, 0, C.byref(o)))
for _ in range(runs):
    res = N.Result()
    ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
    print(f"status {res.status} pivots {res.iters} dev_ms {res.ms_device:.3f} row_reduction_ms {res.ms_rank1:.3f} x{res.n_rank1} launches {res.launches} obj {res.obj!r}")
