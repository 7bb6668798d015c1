"""Pins the CPU oracle against every golden value the reference's own tests hold for the hot path
(reference: tests/integration_tests.rs:51-127, tests/problems/mod.rs:130-674)."""
import numpy as np
import pytest

import problems as P
from oracle import binding as O

SOLVERS = {"primal": O.PRIMAL, "dual": O.DUAL}


@pytest.mark.parametrize("solver", list(SOLVERS))
@pytest.mark.parametrize("make", P.GOLDEN, ids=[f.__name__ for f in P.GOLDEN])
def test_golden_integration(solver, make):
    # `$solver.solve(test_prob.prob).unwrap()` with the Default solver (max_iter = 1000)
    prob, exp = make()
    r = O.solve(prob, SOLVERS[solver], max_iter=1000, mode=O.MODE_EXACT)
    P.check_expectation(exp, r.status_name, r.obj, r.x)


@pytest.mark.parametrize("solver", list(SOLVERS))
@pytest.mark.parametrize("name", P.NETLIB)
def test_golden_netlib(solver, name):
    # feature "benchmarks": Default solver, objective within rel 1e-6 of the published netlib value
    prob, exp = P.netlib(name)
    r = O.solve(prob, SOLVERS[solver], max_iter=1000, mode=O.MODE_EXACT)
    P.check_expectation(exp, r.status_name, r.obj, r.x)
    assert prob.is_feasible(list(r.x)) or name != "afiro"  # afiro's optimum is exactly representable enough


def test_doctest_example_value():
    # src/lib.rs:85,95 / README.md:88,98 print 19.157894736842103 for both solvers
    prob, _ = P.small_prob_1()
    for s in SOLVERS.values():
        r = O.solve(prob, s)
        assert abs(r.obj - 19.157894736842103) < 1e-12


def test_canonical_mode_matches_exact_on_non_degenerate():
    # order-free tie rules coincide with the sequential folds when there are no ties
    rng = np.random.default_rng(7)
    from ellp_b200.problem import Bound, ConstraintOp, Problem
    m, n = 12, 20
    A = rng.random((m, n)); b = rng.uniform(1, 2, m) * n / 4; c = rng.uniform(0.5, 1.5, n)
    p = Problem.new()
    ids = [p.add_var(-c[j], Bound.Lower(0.0)) for j in range(n)]
    for i in range(m):
        p.add_constraint([(ids[j], A[i, j]) for j in range(n)], ConstraintOp.Lte, b[i])
    # skip phase 1 (all-tied artificial pricing): compare phase-2 traces through solve_with_initial
    sf = O.stage(p, 0)
    def run(mode):
        x = np.zeros(sf.n); x[n:] = sf.b  # slack basis (slack for row i is column n_tot-1-i', rows reordered by QR)
        B = np.zeros(sf.m, dtype=np.int32)
        for i in range(sf.m):
            B[i] = int(np.nonzero(sf.A[i, n:])[0][0]) + n
        x[:] = 0; x[B] = sf.b
        N = np.arange(n, dtype=np.int32); Ns = np.zeros(n, dtype=np.uint8)
        return O.solve_with_initial(O.PRIMAL, sf.m, sf.n, sf.A, sf.c, sf.b, sf.kind, sf.lb, sf.ub, x, B, N, Ns,
                                    max_iter=None, mode=mode, trace_cap=4096)
    r0, r1 = run(O.MODE_EXACT), run(O.MODE_CANONICAL)
    assert r0.status == r1.status == O.OPTIMAL
    assert len(r0.trace) == len(r1.trace) > 3
    assert (r0.trace["entering"] == r1.trace["entering"]).all() and (r0.trace["leaving"] == r1.trace["leaving"]).all()


def test_canonical_rule_cycles_on_beale_but_reference_fold_does_not():
    # documents WHY the GPU path reproduces the reference's sequential folds instead of an order-free rule
    prob, exp = P.beale_cycle()
    assert O.solve(prob, O.PRIMAL, 1000, O.MODE_EXACT).status == O.OPTIMAL
    assert O.solve(prob, O.PRIMAL, 1000, O.MODE_CANONICAL).status == O.MAXITER


def test_error_and_panic_paths_are_reported_not_aborted():
    # "invalid B, has {} elements but {} expected" (primal_simplex_solver.rs:125-129)
    A = np.asfortranarray(np.eye(2)); c = np.zeros(2); b = np.ones(2)
    kind = np.ones(2, dtype=np.uint8); lb = np.zeros(2); ub = np.zeros(2)
    x = np.zeros(2); B = np.zeros(1, dtype=np.int32); N = np.zeros(1, dtype=np.int32); Ns = np.zeros(1, dtype=np.uint8)
    with pytest.raises(O.OracleError) as e:
        O.solve_with_initial(O.PRIMAL, 2, 2, A, c, b, kind, lb, ub, x, B, N, Ns)
    assert e.value.code == -1 and "invalid B, has 1 elements but 2 expected" in e.value.msg
