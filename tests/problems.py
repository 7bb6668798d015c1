"""The reference's golden problems (restated from /root/reference/tests/problems/mod.rs:130-674).

Each entry is ``name -> (Problem, expectation)`` where expectation is one of
``("optimal", obj, x)``, ``("optimal_obj", obj)``, ``("unbounded",)``, ``("infeasible",)`` --
the four assertion macros of the reference (tests/problems/mod.rs:9-71; EPS 1e-8, REL_EPS 1e-6).
"""
from __future__ import annotations

import json
import os

from ellp_b200.problem import Bound, ConstraintOp, Problem

ABS_EPS = 1e-8   # tests/problems/mod.rs:6
REL_EPS = 1e-6   # tests/problems/mod.rs:7
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

Lte, Eq, Gte = ConstraintOp.Lte, ConstraintOp.Eq, ConstraintOp.Gte


def _lp(vars_, cons):
    """vars_: [(obj, Bound, name)], cons: [([(var_pos, coeff)], op, rhs)]"""
    p = Problem.new()
    ids = [p.add_var(c, b, n) for c, b, n in vars_]
    for coeffs, op, rhs in cons:
        p.add_constraint([(ids[i], a) for i, a in coeffs], op, rhs)
    return p


def empty_problem():  # :130
    return Problem.new(), ("optimal", 0.0, [])


def one_variable_no_constraints():  # :135
    return _lp([(2.0, Bound.TwoSided(-1.0, 1.0), "x1")], []), ("optimal", -2.0, [-1.0])


def one_variable_infeasible():  # :144
    return _lp([(2.0, Bound.Upper(0.0), "x1")], [([(0, 1.0)], Gte, 1.0)]), ("infeasible",)


def one_variable_unbounded_upper():  # :156
    return _lp([(2.0, Bound.Upper(0.0), "x1")], []), ("unbounded",)


def one_variable_unbounded_free():  # :165
    return _lp([(2.0, Bound.Free(), "x1")], []), ("unbounded",)


def two_variables_unbounded():  # :174
    return _lp([(2.0, Bound.Lower(0.0), "x1"), (2.0, Bound.Upper(1.0), "x2")], []), ("unbounded",)


def two_variables_infeasible_with_bounds():  # :186
    return _lp([(2.0, Bound.Lower(0.0), "x1"), (2.0, Bound.Lower(1.0), "x2")],
               [([(0, 1.0), (1, 1.0)], Lte, 0.0)]), ("infeasible",)


def two_variables_infeasible_free():  # :203
    return _lp([(2.0, Bound.Free(), "x1"), (2.0, Bound.Free(), "x2")],
               [([(0, 1.0), (1, 1.0)], Eq, -1.0), ([(0, 2.0), (1, 2.0)], Eq, 1.0)]), ("infeasible",)


def infeasible_constraint_without_coeffs():  # :223
    return _lp([(2.0, Bound.Free(), "x1")], [([], Eq, 1.0)]), ("infeasible",)


def feasible_constraint_without_coeffs():  # :234
    return _lp([(2.0, Bound.Lower(3.0), "x1")], [([], Eq, 0.0)]), ("optimal", 6.0, [3.0])


def feasible_constraint_without_coeffs_and_no_vars():  # :245
    return _lp([], [([], Eq, 0.0)]), ("optimal", 0.0, [])


def infeasible_constraint_without_coeffs_and_no_vars():  # :251
    return _lp([], [([], Eq, 1.0)]), ("infeasible",)


def linear_system_2d():  # :257
    return _lp([(0.0, Bound.Free(), "x"), (0.0, Bound.Free(), "y")],
               [([(0, 2.0), (1, 1.0)], Eq, 1.0), ([(0, 3.0), (1, 1.0)], Eq, 1.0)]), ("optimal", 0.0, [0.0, 1.0])


def _linear_system_3d(last):
    return _lp([(0.0, Bound.Free(), "x"), (0.0, Bound.Free(), "y"), (0.0, Bound.Free(), "z")],
               [([(0, 1.0), (1, 2.0), (2, 4.0)], Eq, 1.0), ([(0, 3.0), (1, 4.0), (2, 8.0)], Eq, 2.0),
                ([(0, 5.0), (1, 6.0), (2, last)], Eq, 5.0)])


def linear_system_3d():  # :277
    return _linear_system_3d(13.0), ("optimal", 0.0, [0.0, -3.5, 2.0])


def linear_system_3d_infeasible():  # :304
    return _linear_system_3d(12.0), ("infeasible",)


def small_prob_1():  # :331 (the README / doctest example, src/lib.rs:11-66)
    p = _lp([(2.0, Bound.TwoSided(-1.0, 1.0), "x1"), (10.0, Bound.Upper(6.0), "x2"), (0.0, Bound.Lower(0.0), "x3"),
             (1.0, Bound.Fixed(0.0), "x4"), (0.0, Bound.Free(), "x5")],
            [([(0, 2.5), (1, 3.5)], Gte, 5.0), ([(1, 2.5), (0, 4.5)], Lte, 1.0),
             ([(2, -1.0), (3, -3.0), (4, -4.0)], Eq, 2.0)])
    return p, ("optimal", 19.1578947368421, [-0.94736842105, 2.105263157894, 0.0, 0.0, -0.5])


def small_prob_2():  # :372
    return _lp([(-5.0, Bound.Lower(0.0), "x"), (-4.0, Bound.Lower(0.0), "y")],
               [([(0, 1.0)], Lte, 6.0), ([(0, 0.25), (1, 1.0)], Lte, 6.0), ([(0, 3.0), (1, 2.0)], Lte, 22.0)]), \
        ("optimal", -40.0, [4.0, 5.0])


def small_prob_3():  # :396
    return _lp([(3.0, Bound.Lower(0.0), "x"), (-6.0, Bound.Lower(0.0), "y")],
               [([(0, 1.0), (1, 2.0)], Gte, -1.0), ([(0, 2.0), (1, 1.0)], Gte, 0.0), ([(0, 1.0), (1, -1.0)], Gte, -1.0),
                ([(0, 1.0), (1, -4.0)], Gte, -13.0), ([(0, -4.0), (1, 1.0)], Gte, -23.0)]), \
        ("optimal", -15.0, [3.0, 4.0])


def small_prob_4():  # :426 (multiple optima: objective only)
    return _lp([(-1.0, Bound.Lower(0.0), "x"), (-1.0, Bound.Lower(0.0), "y"), (-1.0, Bound.Lower(0.0), "z")],
               [([(0, 1.0), (1, -1.0), (2, 1.0)], Gte, -2.0), ([(0, -1.0), (1, 1.0), (2, 1.0)], Gte, -3.0),
                ([(0, 1.0), (1, 1.0), (2, -1.0)], Gte, -1.0), ([(0, -1.0), (1, -1.0), (2, -1.0)], Gte, -4.0)]), \
        ("optimal_obj", -4.0)


def small_prob_5():  # :457
    return _lp([(4.0, Bound.Lower(0.0), "x"), (5.0, Bound.Lower(0.0), "y")],
               [([(0, 1.0), (1, 1.0)], Gte, -1.0), ([(0, 1.0), (1, 2.0)], Gte, 1.0), ([(0, 4.0), (1, 2.0)], Gte, 8.0),
                ([(0, -1.0), (1, -1.0)], Gte, -3.0), ([(0, -1.0), (1, 1.0)], Gte, 1.0)]), \
        ("optimal", 14.0, [1.0, 2.0])


def small_prob_6():  # :486
    return _lp([(-2.0, Bound.Lower(0.0), "x"), (-4.0, Bound.Lower(0.0), "y"), (-1.0, Bound.Lower(0.0), "z"),
                (-1.0, Bound.Lower(0.0), "w")],
               [([(0, -1.0), (1, -3.0), (3, -1.0)], Gte, -4.0), ([(0, -2.0), (1, -1.0)], Gte, -3.0),
                ([(1, -1.0), (2, -4.0), (3, -1.0)], Gte, -3.0), ([(0, 1.0), (1, 1.0), (2, 2.0)], Gte, 1.0),
                ([(0, -1.0), (1, 1.0), (2, 4.0)], Gte, 1.0)]), \
        ("optimal", -6.5, [1.0, 1.0, 0.5, 0.0])


def small_prob_7():  # :523
    return _lp([(2.0, Bound.Lower(0.0), "x"), (-1.0, Bound.Lower(0.0), "y"), (1.0, Bound.Free(), "z")],
               [([(0, 1.0), (1, -1.0), (2, 4.0)], Gte, -1.0), ([(0, 1.0), (1, -1.0), (2, -1.0)], Gte, 2.0),
                ([(0, 1.0), (1, 3.0), (2, 2.0)], Eq, 3.0)]), \
        ("optimal", 2.9, [2.1, 0.7, -0.6])


def small_prob_unbounded_1():  # :552
    return _lp([(-2.0, Bound.Lower(0.0), "x"), (-3.0, Bound.Lower(0.0), "y"), (1.0, Bound.Lower(0.0), "z")],
               [([(0, 1.0), (1, 1.0), (2, 1.0)], Gte, -3.0), ([(0, -1.0), (1, 1.0), (2, -1.0)], Gte, -4.0),
                ([(0, 1.0), (1, -1.0), (2, -2.0)], Gte, -1.0)]), ("unbounded",)


def small_prob_unbounded_2():  # :580
    return _lp([(-2.0, Bound.Lower(0.0), "x"), (-3.0, Bound.Lower(0.0), "y"), (1.0, Bound.Lower(0.0), "z"),
                (1.0, Bound.Lower(0.0), "w")],
               [([(1, 1.0), (2, -2.0), (3, -1.0)], Gte, -4.0), ([(0, 2.0), (1, -1.0), (2, -1.0), (3, 4.0)], Gte, -5.0),
                ([(0, -1.0), (1, 1.0), (3, -2.0)], Gte, -3.0)]), ("unbounded",)


def beale_cycle():  # :616
    return _lp([(-10.0, Bound.Lower(0.0), "x"), (57.0, Bound.Lower(0.0), "y"), (9.0, Bound.Lower(0.0), "z"),
                (24.0, Bound.Lower(0.0), "w")],
               [([(0, -0.5), (1, 5.5), (2, 2.5), (3, -9.0)], Gte, 0.0), ([(0, -0.5), (1, 1.5), (2, 0.5), (3, -1.0)], Gte, 0.0),
                ([(0, -1.0)], Gte, -1.0)]), ("optimal", -1.0, [1.0, 0.0, 1.0, 0.0])


# order of tests/integration_tests.rs:51-109
GOLDEN = [
    empty_problem, one_variable_no_constraints, one_variable_infeasible, one_variable_unbounded_upper,
    one_variable_unbounded_free, two_variables_unbounded, two_variables_infeasible_with_bounds,
    two_variables_infeasible_free, infeasible_constraint_without_coeffs, feasible_constraint_without_coeffs,
    feasible_constraint_without_coeffs_and_no_vars, infeasible_constraint_without_coeffs_and_no_vars,
    linear_system_2d, linear_system_3d, linear_system_3d_infeasible, small_prob_1, small_prob_2, small_prob_3,
    small_prob_4, small_prob_5, small_prob_6, small_prob_7, small_prob_unbounded_1, small_prob_unbounded_2,
    beale_cycle,
]
GOLDEN_BY_NAME = {f.__name__: f for f in GOLDEN}

NETLIB = ["afiro", "adlittle", "blend"]  # tests/integration_tests.rs:111-127, tests/problems/mod.rs:657-674


def load_netlib_fixture(name: str) -> dict:
    with open(os.path.join(GOLDEN_DIR, f"netlib_{name}.json")) as f:
        return json.load(f)


def netlib(name: str):
    """Problem in FILE order (rows as listed, columns in order of first appearance); every variable Lower(0)
    (no BOUNDS section: src/parse_mps.rs:31)."""
    fx = load_netlib_fixture(name)
    p = Problem.new()
    ids = [p.add_var(c["obj"], Bound.Lower(0.0), c["name"]) for c in fx["cols"]]
    per_row = [[] for _ in fx["rows"]]
    for j, c in enumerate(fx["cols"]):
        for r, v in zip(c["rows"], c["vals"]):
            per_row[r].append((ids[j], v))
    for r, row in enumerate(fx["rows"]):
        p.add_constraint(per_row[r], ConstraintOp(row["op"]), row["rhs"])
    return p, ("optimal_obj", fx["expected_obj"])


def netlib_mps_text(name: str) -> str:
    """Re-serialises the fixture in the layout the reference's reader accepts (src/parse_mps.rs:68-115)."""
    fx = load_netlib_fixture(name)
    out = [f"NAME          {name.upper()}", "ROWS", " N  COST"]
    for row in fx["rows"]:
        out.append(f" {'LEG'[row['op']]}  {row['name']}")
    out.append("COLUMNS")
    for c in fx["cols"]:
        if c["obj"] != 0.0:
            out.append(f"    {c['name']}  COST  {c['obj']!r}")
        for r, v in zip(c["rows"], c["vals"]):
            out.append(f"    {c['name']}  {fx['rows'][r]['name']}  {v!r}")
    out.append("RHS")
    for row in fx["rows"]:
        if row["rhs"] != 0.0:
            out.append(f"    B  {row['name']}  {row['rhs']!r}")
    out.append("ENDATA")
    return "\n".join(out) + "\n"


def check_expectation(exp, status_name: str, obj: float, x) -> None:
    """The reference's assert_* macros (tests/problems/mod.rs:9-71)."""
    kind = exp[0]
    if kind == "infeasible":
        assert status_name == "Infeasible", f"not infeasible: {status_name}"
    elif kind == "unbounded":
        assert status_name == "Unbounded", f"not unbounded: {status_name}"
    elif kind == "optimal":
        assert status_name == "Optimal", f"not optimal: {status_name}"
        assert abs(obj - exp[1]) < ABS_EPS, f"obj: {obj}, expected: {exp[1]}"
        assert len(x) == len(exp[2])
        for x1, x2 in zip(x, exp[2]):
            assert abs(x1 - x2) < ABS_EPS, f"x_i: {x1}, expected: {x2}"
    elif kind == "optimal_obj":
        assert status_name == "Optimal", f"not optimal: {status_name}"
        assert abs(obj - exp[1]) < ABS_EPS or abs(obj / exp[1] - 1.0) < REL_EPS, f"obj: {obj}, expected: {exp[1]}"
    else:
        raise AssertionError(kind)


def random_lp_all_bounds(seed):
    """Seeded random LP with every Bound kind incl. TwoSided (quirks Q3 / Q17 live there) and <=, =, >= rows.  Used for
    GPU-vs-oracle fuzzing only: the reference's own verdicts on TwoSided variables are not comparable with an exact solver
    (tests/test_oracle_highs.py explains why), but whatever the reference does, the GPU path must do the same."""
    import numpy as np
    from ellp_b200.problem import Bound, ConstraintOp, Problem
    rng = np.random.default_rng(1000 + seed)
    nv, nc = int(rng.integers(2, 10)), int(rng.integers(1, 7))
    p = Problem.new()
    ids = []
    for j in range(nv):
        kind = rng.choice(["lower", "upper", "free", "fixed", "twosided", "twosided", "twosided"])
        c = float(np.round(rng.normal(), 2))
        if kind == "lower":
            b = Bound.Lower(float(rng.integers(-2, 3)))
        elif kind == "upper":
            b = Bound.Upper(float(rng.integers(0, 6)))
        elif kind == "free":
            b = Bound.Free()
        elif kind == "fixed":
            b = Bound.Fixed(float(rng.integers(-1, 3)))
        else:
            lo = float(rng.integers(-3, 2))
            b = Bound.TwoSided(lo, lo + float(rng.integers(1, 6)))
        ids.append(p.add_var(c, b, f"x{j}"))
    for i in range(nc):
        row = np.round(rng.normal(size=nv), 2)
        row[rng.random(nv) < 0.3] = 0.0
        if not row.any():
            row[int(rng.integers(nv))] = 1.0
        rhs = float(np.round(rng.normal() * 3, 2))
        op = rng.choice([ConstraintOp.Lte, ConstraintOp.Gte, ConstraintOp.Eq], p=[0.45, 0.35, 0.2])
        p.add_constraint([(ids[j], float(row[j])) for j in range(nv) if row[j] != 0.0], op, rhs)
    return p
