"""Generates tests/golden/highs_dense_lp.json: HiGHS optima (scipy.optimize.linprog, method="highs") of the synthetic dense LPs of
bench_lp.dense_lp (numpy twin of the device generator), used by the full-solve parity tests / the full-solve bench workload:
    variant 0: min -c.x  s.t. A x <= b, x >= 0      variant 1: min c.x  s.t. A x >= b, x >= 0
Run here (CPU, minutes):  python tests/golden/make_highs_dense_fixture.py  [m ns seed variant ...]"""
import json, os, sys, time
import numpy as np
from scipy.optimize import linprog
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench_lp

out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "highs_dense_lp.json")
cases = [(512, 1024, 0, 0), (512, 1024, 0, 1), (4096, 8192, 0, 0), (4096, 8192, 0, 1)]
if len(sys.argv) > 1:
    v = [int(a) for a in sys.argv[1:]]
    cases = [tuple(v[i:i + 4]) for i in range(0, len(v), 4)]
res = {}
if os.path.exists(out_path):
    res = json.load(open(out_path))
for (m, ns, seed, variant) in cases:
    lp = bench_lp.dense_lp(m, ns, seed, variant)
    A = lp["A"][:, :ns]; c = lp["c"][:ns]; b = lp["b"]
    t0 = time.time()
    r = linprog(c, A_ub=A if variant == 0 else -A, b_ub=b if variant == 0 else -b, bounds=[(0, None)] * ns, method="highs")
    key = f"{m}x{ns}_seed{seed}_variant{variant}"
    res[key] = {"status": int(r.status), "obj": float(r.fun), "seconds": time.time() - t0, "nit": int(r.nit)}
    print(key, res[key], flush=True)
    json.dump(res, open(out_path, "w"), indent=1)
