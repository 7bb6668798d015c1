"""Independent cross-check of the ORACLE (test infrastructure): its restatement of ellp's two solvers against HiGHS
(scipy.optimize.linprog) on seeded random LPs.  SURVEY 8(c) names HiGHS as the offline cross-check; the reference's own golden
vectors pin the oracle in test_oracle_golden.py, this file adds problems the reference has no expectation for.
Only Lower / Upper / Free / Fixed bounds: quirk Q3 (TwoSided `<` vs `>`, primal :359-363) makes the reference itself -- and
hence the oracle in exact mode -- return wrong optima on TwoSided variables (demonstrated at the bottom)."""
import numpy as np
import pytest

from ellp_b200.problem import Bound, ConstraintOp, Problem
from oracle import binding as O

scipy_opt = pytest.importorskip("scipy.optimize")


def _random_lp(seed):
    rng = np.random.default_rng(seed)
    nv, nc = int(rng.integers(2, 9)), int(rng.integers(1, 7))
    p = Problem.new()
    lo, hi, ids = [], [], []
    for j in range(nv):
        kind = rng.choice(["lower", "lower", "lower", "upper", "free", "fixed"])
        c = float(np.round(rng.normal(), 2))
        if kind == "lower":
            b = Bound.Lower(float(rng.integers(-2, 3))); lo.append(b.lb); hi.append(None)
        elif kind == "upper":
            b = Bound.Upper(float(rng.integers(0, 6))); lo.append(None); hi.append(b.ub)
        elif kind == "free":
            b = Bound.Free(); lo.append(None); hi.append(None)
        else:
            v = float(rng.integers(-1, 3)); b = Bound.Fixed(v); lo.append(v); hi.append(v)
        ids.append(p.add_var(c, b, f"x{j}"))
    A_ub, b_ub, A_eq, b_eq = [], [], [], []
    for i in range(nc):
        row = np.round(rng.normal(size=nv), 2)
        row[rng.random(nv) < 0.3] = 0.0
        if not row.any():
            row[int(rng.integers(nv))] = 1.0
        rhs = float(np.round(rng.normal() * 3, 2))
        op = rng.choice(["lte", "gte", "eq"], p=[0.45, 0.35, 0.2])
        terms = [(ids[j], float(row[j])) for j in range(nv) if row[j] != 0.0]
        if op == "lte":
            p.add_constraint(terms, ConstraintOp.Lte, rhs); A_ub.append(row); b_ub.append(rhs)
        elif op == "gte":
            p.add_constraint(terms, ConstraintOp.Gte, rhs); A_ub.append(-row); b_ub.append(-rhs)
        else:
            p.add_constraint(terms, ConstraintOp.Eq, rhs); A_eq.append(row); b_eq.append(rhs)
    cost = [v.obj_coeff for v in p.variables]
    return p, cost, A_ub, b_ub, A_eq, b_eq, list(zip(lo, hi))


@pytest.mark.parametrize("seed", range(80))
def test_oracle_agrees_with_highs_on_random_lps(seed):
    p, cost, A_ub, b_ub, A_eq, b_eq, bounds = _random_lp(seed)
    hs = scipy_opt.linprog(cost, A_ub=np.array(A_ub) if A_ub else None, b_ub=b_ub or None, A_eq=np.array(A_eq) if A_eq else None,
                           b_eq=b_eq or None, bounds=bounds, method="highs")
    for which in (O.PRIMAL, O.DUAL):
        if which == O.DUAL and any(lo is not None and lo == hi for lo, hi in bounds):
            # dual :200-236 never selects a Fixed (or Free) basic variable as the leaving one, while a nonbasic Fixed variable
            # is priced like Lower and may enter (:263-279): once basic it is never driven back to its value.  The reference --
            # and therefore the oracle -- returns such points as Optimal; not comparable with HiGHS.
            continue
        if which == O.DUAL and len(A_eq) >= len(cost):
            # Q10 (dual :175-177): a square standard form has no nonbasic variable and the dual solver returns its starting point
            # as Optimal without checking the bounds -- the oracle reproduces that, HiGHS (rightly) does not
            continue
        try:
            r = O.solve(p, which, 1000, O.MODE_EXACT)
        except O.OracleError as e:  # the reference panics / errors on a few structures (quirks Q12, Q16, Q17): not a disagreement
            pytest.skip(f"reference quirk: {e}")
        if hs.status == 0:
            # Q12: the reference reports Infeasible for consistent but redundant equality rows (rank test on R's diagonal)
            if r.status_name == "Infeasible" and len(A_eq) > 0:
                continue
            assert r.status_name == "Optimal", (seed, which, r.status_name, hs.message)
            assert abs(r.obj - hs.fun) <= 1e-7 * max(1.0, abs(hs.fun)), (seed, which, r.obj, hs.fun)
        elif hs.status == 2:
            assert r.status_name == "Infeasible", (seed, which, r.status_name)
        elif hs.status == 3:
            assert r.status_name in ("Unbounded", "Infeasible"), (seed, which, r.status_name)  # HiGHS may call an infeasible LP unbounded (dual ray)


@pytest.mark.parametrize("variant", [0, 1])
def test_oracle_agrees_with_highs_on_the_dense_generator_lp(variant):
    """the bench generator's dense LP (numpy twin, 256 x 512) solved to optimality by the oracle's restatement of
    solve_with_initial (slack start; primal for `A x <= b`, dual for `A x >= b`) against HiGHS (tests/golden/highs_dense_lp.json)."""
    import json, os
    import bench_lp
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "highs_dense_lp.json")))
    m, ns = 256, 512
    key = f"{m}x{ns}_seed0_variant{variant}"
    lp = bench_lp.dense_lp(m, ns, 0, variant)
    st = [lp[k].copy() for k in ("x", "B", "N", "N_side")] + ([lp["y"].copy(), lp["d"].copy()] if variant else [])
    r = O.solve_with_initial(O.DUAL if variant else O.PRIMAL, m, m + ns, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], *st, max_iter=100000)
    assert r.status == O.OPTIMAL == fx[key]["status"]
    obj = float(lp["c"] @ st[0])
    assert abs(obj - fx[key]["obj"]) <= 1e-9 * max(1.0, abs(fx[key]["obj"]))
    assert np.abs(lp["A"] @ st[0] - lp["b"]).max() <= 1e-9 * np.abs(lp["b"]).max()
