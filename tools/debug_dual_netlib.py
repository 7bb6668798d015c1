"""Debug aid: dual phase 1 of a netlib LP on the tableau engine vs the revised engine vs the oracle (condition numbers of the
visited bases, pivot counts) for several refactor periods / block sizes."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import problems as P
from ellp_b200 import _native as N
from ellp_b200 import solver as S
from ellp_b200.problem import EllPError
from oracle import binding as O

name = sys.argv[1] if len(sys.argv) > 1 else "adlittle"
prob, _ = P.netlib(name)
st = O.stage(prob, 2)
m, n = st.m, st.n
A = np.asfortranarray(st.A); c = np.array(st.c); b = np.array(st.b); kind = np.array(st.kind); lb = np.array(st.lb); ub = np.array(st.ub)
start = [np.array(st.x), np.array(st.B, dtype=np.int32), np.array(st.N, dtype=np.int32), np.array(st.N_side, dtype=np.uint8), np.array(st.y), np.array(st.d)]
ref = O.solve_with_initial(O.DUAL, m, n, A, c, b, kind, lb, ub, *[a.copy() for a in start], max_iter=1000, trace_cap=4096)
print(f"{name} dual phase 1: m={m} n={n} oracle status={ref.status} pivots={len(ref.trace)} obj={ref.obj!r}")
ctx = N.Context(0)
for engine, bk, rf in [(N.ENGINE_REVISED, 0, 0), (N.ENGINE_TABLEAU, 8, 0), (N.ENGINE_TABLEAU, 8, 100), (N.ENGINE_TABLEAU, 8, 25), (N.ENGINE_TABLEAU, 8, 100000),
                       (N.ENGINE_TABLEAU, 32, 100), (N.ENGINE_TABLEAU, 2, 100)]:
    sg = [a.copy() for a in start]
    sol = S.GpuDualSimplexSolver.new(1000, ctx=ctx, engine=engine, block_k=bk, refactor_every=rf, trace_cap=4096, check_every=4)
    try:
        res, tr = sol.solve_with_initial(m, n, A, c, b, kind, lb, ub, *sg)
        k = min(len(tr), len(ref.trace))
        same = int(np.argmax(np.concatenate([(tr["entering"][:k] != ref.trace["entering"][:k]) | (tr["leaving"][:k] != ref.trace["leaving"][:k]), [True]])))
        cond = np.linalg.cond(A[:, sg[1]])
        resid = np.abs(A @ sg[0] - b).max()
        print(f"  engine={engine} bk={bk} refactor_every={rf}: status={res.status} pivots={res.iters} obj={res.obj!r} first_divergence={same} cond(B_final)={cond:.3g} |Ax-b|={resid:.3g} max|step|={np.abs(tr['step']).max():.3g}")
    except Exception as e:
        print(f"  engine={engine} bk={bk} refactor_every={rf}: EXC {type(e).__name__}: {e}")
ctx.close()

# primal Devex through the public API (two phases) for several rebuild periods
from ellp_b200.solver import GpuPrimalSimplexSolver
ctx = N.Context(0)
refp = O.solve(prob, O.PRIMAL, 1000, O.MODE_EXACT)
print(f"{name} primal: oracle {refp.status_name} obj={refp.obj!r} iters={refp.iters}")
for pricing, rf, bk in [(N.PRICE_REFERENCE, 0, 16), (N.PRICE_DEVEX, 0, 16), (N.PRICE_DEVEX, 25, 16), (N.PRICE_DEVEX, 50, 16), (N.PRICE_DEVEX, 10, 8)]:
    try:
        r = GpuPrimalSimplexSolver.default(ctx=ctx, engine=N.ENGINE_TABLEAU, block_k=bk, pricing=pricing, tie_rule=N.TIES_CANONICAL, refactor_every=rf).solve(prob)
        print(f"  pricing={pricing} refactor_every={rf} bk={bk}: {r.kind} obj={r.solution.obj() if r.is_optimal else None!r} iters={r.iters}")
    except Exception as e:
        print(f"  pricing={pricing} refactor_every={rf} bk={bk}: EXC {type(e).__name__}: {e}")
ctx.close()
