"""Complete solves of the bench generator's dense LP through the C ABI on host buffers (slack start): pivots to optimality, device
time, wall time incl. upload / download, rebuild count, objective against the HiGHS fixture (tests/golden/highs_dense_lp.json).
    python tools/full_solve_stats.py [m ns] -> one JSON line per (variant, pricing)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench_lp
from ellp_b200 import _native as N
from ellp_b200 import solver as S

m, ns = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 8192)
fx = json.load(open(os.path.join(ROOT, "tests", "golden", "highs_dense_lp.json")))
ctx = N.Context(0)
for variant, pricing, rf in [(0, "dantzig", 0), (0, "devex", 0), (1, "reference", 0), (1, "devex", 0), (0, "dantzig", 2000), (1, "devex", 500)]:
    lp = bench_lp.dense_lp(m, ns, 0, variant)
    key = f"{m}x{ns}_seed0_variant{variant}"
    cls = S.GpuDualSimplexSolver if variant else S.GpuPrimalSimplexSolver
    sol = cls.new(None, ctx=ctx, engine=N.ENGINE_TABLEAU, block_k=48, check_every=96, refactor_every=rf,
                  pricing=N.PRICE_DEVEX if pricing == "devex" else N.PRICE_REFERENCE, tie_rule=N.TIES_CANONICAL)
    best = None
    ctx.set_tuning("fast_upload", 0 if rf else 1)  # a rebuild needs the whole A on the device
    for rep in range(2):
        st = [lp[k].copy() for k in ("x", "B", "N", "N_side")] + ([lp["y"].copy(), lp["d"].copy()] if variant else [])
        t0 = time.perf_counter()
        res, _ = sol.solve_with_initial(m, m + ns, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], *st)
        wall = time.perf_counter() - t0
        best = wall if best is None else min(best, wall)
    x = st[0]
    obj = float(lp["c"] @ x)
    out = {"lp": key, "solver": "dual" if variant else "primal", "pricing": pricing, "refactor_every": rf, "status": int(res.status), "pivots": int(res.iters),
           "rebuilds": int(res.refactors), "ms_device": float(res.ms_device), "ms_wall_host_buffers": 1e3 * best, "obj": obj,
           "highs_obj": fx.get(key, {}).get("obj"), "rel_err_vs_highs": abs(obj - fx[key]["obj"]) / max(1.0, abs(fx[key]["obj"])) if key in fx else None,
           "primal_residual_rel": float(np.abs(lp["A"] @ x - lp["b"]).max() / np.abs(lp["b"]).max()), "highs_seconds": fx.get(key, {}).get("seconds")}
    print(json.dumps(out), flush=True)
ctx.close()
