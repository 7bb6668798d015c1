"""Model builder with the API shape of ellp's ``Problem`` (reference: src/problem.rs:11-303).

``Problem.add_var`` / ``add_var_with_id`` / ``add_constraint`` / ``is_feasible``, ``Bound``,
``ConstraintOp`` and ``VariableId`` keep the reference's names, argument meaning and error
behaviour (``EllPError`` with the reference's messages).  This is host-side bookkeeping only:
the arithmetic lives behind the C ABI in ``libellp_b200.so``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from enum import IntEnum
from typing import List, Optional, Sequence, Tuple

import numpy as np

EPS = 1e-10  # src/util.rs:1


def rust_f64(v: float) -> str:
    """`format!("{}", v)` of Rust for an f64: shortest round-trip digits, never an exponent, integers without ".0"."""
    v = float(v)
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    r = repr(v)
    if "e" in r or "E" in r:
        from decimal import Decimal
        r = format(Decimal(r), "f")
    if r.endswith(".0"):
        r = r[:-2]
    return r


class EllPError(Exception):
    """src/error.rs:3-11"""


class ConstraintOp(IntEnum):  # src/problem.rs:298-303; values cross the C ABI
    Lte = 0
    Eq = 1
    Gte = 2


class BoundKind(IntEnum):  # src/problem.rs:190-197; values cross the C ABI
    Free = 0
    Lower = 1
    Upper = 2
    TwoSided = 3
    Fixed = 4


@dataclass(frozen=True)
class Bound:
    """``Bound::{Free, Lower(lb), Upper(ub), TwoSided(lb, ub), Fixed(v)}`` (src/problem.rs:190-197)."""

    kind: BoundKind
    lb: float = 0.0
    ub: float = 0.0

    @staticmethod
    def Free() -> "Bound":
        return Bound(BoundKind.Free)

    @staticmethod
    def Lower(lb: float) -> "Bound":
        return Bound(BoundKind.Lower, float(lb), 0.0)

    @staticmethod
    def Upper(ub: float) -> "Bound":
        return Bound(BoundKind.Upper, 0.0, float(ub))

    @staticmethod
    def TwoSided(lb: float, ub: float) -> "Bound":
        return Bound(BoundKind.TwoSided, float(lb), float(ub))

    @staticmethod
    def Fixed(val: float) -> "Bound":
        return Bound(BoundKind.Fixed, float(val), float(val))

    def __str__(self) -> str:  # impl Display for Bound, src/problem.rs:213-223
        inf = "\u221e"
        if self.kind == BoundKind.Free:
            return f"(-{inf}, {inf})"
        if self.kind == BoundKind.Lower:
            return f"[{rust_f64(self.lb)}, {inf})"
        if self.kind == BoundKind.Upper:
            return f"(-{inf}, {rust_f64(self.ub)}]"
        if self.kind == BoundKind.TwoSided:
            return f"[{rust_f64(self.lb)}, {rust_f64(self.ub)}]"
        return f"[{rust_f64(self.lb)}, {rust_f64(self.lb)}]"

    def display(self, var: "Variable") -> str:  # Bound::display, src/problem.rs:199-210
        lte, gte = "\u2264", "\u2265"
        if self.kind == BoundKind.Free:
            return f"{var} free"
        if self.kind == BoundKind.Lower:
            return f"{var} {gte} {rust_f64(self.lb)}"
        if self.kind == BoundKind.Upper:
            return f"{var} {lte} {rust_f64(self.ub)}"
        if self.kind == BoundKind.TwoSided:
            return f"{rust_f64(self.lb)} {lte} {var} {lte} {rust_f64(self.ub)}"
        return f"{var} = {rust_f64(self.lb)}"


@dataclass(frozen=True)
class VariableId:  # src/problem.rs:277-296
    id: int

    def __int__(self) -> int:
        return self.id

    def __index__(self) -> int:
        return self.id


@dataclass
class Variable:  # src/problem.rs:156-188
    id: VariableId
    obj_coeff: float
    bound: Bound
    name: Optional[str] = None

    def __str__(self) -> str:  # impl Display for Variable, src/problem.rs:353-360
        return self.name if self.name is not None else f"id[{int(self.id)}]"


@dataclass
class Constraint:  # src/problem.rs:225-275
    coeffs: List[Tuple[VariableId, float]]
    op: ConstraintOp
    rhs: float

    def add_coeff(self, var: VariableId, coeff: float) -> None:
        self.coeffs.append((var, coeff))

    def is_feasible(self, x: Sequence[float]) -> bool:  # :236-249
        lhs = 0.0
        for var, coeff in self.coeffs:
            lhs += coeff * x[int(var)]
        if self.op == ConstraintOp.Lte:
            return lhs <= self.rhs + EPS
        if self.op == ConstraintOp.Eq:
            return abs(lhs - self.rhs) < EPS
        return lhs >= self.rhs - EPS


@dataclass
class Problem:
    """``ellp::Problem`` (src/problem.rs:11-154)."""

    variables: List[Variable] = field(default_factory=list)
    constraints: List[Constraint] = field(default_factory=list)
    _var_names: set = field(default_factory=set, repr=False)
    _var_ids: set = field(default_factory=set, repr=False)

    @staticmethod
    def new() -> "Problem":
        return Problem()

    def add_var(self, obj_coeff: float, bound: Bound, name: Optional[str] = None) -> VariableId:  # :24-33
        vid = VariableId(len(self.variables))
        self.add_var_with_id(obj_coeff, bound, vid, name)
        return vid

    def add_var_with_id(self, obj_coeff: float, bound: Bound, id: VariableId, name: Optional[str] = None) -> VariableId:  # :35-84
        if not isinstance(id, VariableId):
            id = VariableId(int(id))
        if bound.kind == BoundKind.TwoSided and bound.lb > bound.ub:
            raise EllPError(f"invalid variable bounds: ({bound.lb}, {bound.ub})")
        valid = {
            BoundKind.Free: True,
            BoundKind.Lower: math.isfinite(bound.lb),
            BoundKind.Upper: math.isfinite(bound.ub),
            BoundKind.TwoSided: math.isfinite(bound.lb) and math.isfinite(bound.ub),
            BoundKind.Fixed: math.isfinite(bound.lb),
        }[bound.kind]
        if not valid:
            raise EllPError(f"invalid bound: {bound!r}")
        if name is not None:
            if name in self._var_names:
                raise EllPError(f"variable names must be unique, {name} was added twice")
            self._var_names.add(name)
        # the reference pushes the variable BEFORE the id check (:74-82); keep that order
        self.variables.append(Variable(id, float(obj_coeff), bound, name))
        if id in self._var_ids:
            raise EllPError(f"cannot add variable with {id!r}, that id is already used")
        self._var_ids.add(id)
        return id

    def add_constraint(self, coeffs: Sequence[Tuple[VariableId, float]], op: ConstraintOp, rhs: float) -> None:  # :86-106
        coeffs = [(v if isinstance(v, VariableId) else VariableId(int(v)), float(c)) for v, c in coeffs]
        for vid, _ in coeffs:
            if vid not in self._var_ids:
                raise EllPError(f"{vid!r} is invalid")
        self.constraints.append(Constraint(list(coeffs), ConstraintOp(op), float(rhs)))

    def is_feasible(self, x: Sequence[float]) -> bool:  # :108-153
        if len(x) != len(self.variables):
            return False
        for var, val in zip(self.variables, x):
            b = var.bound
            if b.kind == BoundKind.Lower and val < b.lb - EPS:
                return False
            if b.kind == BoundKind.Upper and val > b.ub + EPS:
                return False
            if b.kind == BoundKind.TwoSided and (val < b.lb - EPS or val > b.ub + EPS):
                return False
            if b.kind == BoundKind.Fixed and abs(val - b.lb) > EPS:
                return False
        return all(c.is_feasible(x) for c in self.constraints)

    # ------------------------------------------------------------------ flat views for the C ABI
    def to_arrays(self) -> dict:
        """CSR-over-variable-ids view consumed by the C ABI (include/ellp_b200.h: ellp_problem_desc)."""
        nv, nc = len(self.variables), len(self.constraints)
        obj = np.array([v.obj_coeff for v in self.variables], dtype=np.float64)
        kind = np.array([int(v.bound.kind) for v in self.variables], dtype=np.uint8)
        lb = np.array([v.bound.lb for v in self.variables], dtype=np.float64)
        ub = np.array([v.bound.ub for v in self.variables], dtype=np.float64)
        var_id = np.array([int(v.id) for v in self.variables], dtype=np.int64)
        row_ptr = np.zeros(nc + 1, dtype=np.int32)
        col_id: List[int] = []
        coef: List[float] = []
        for i, c in enumerate(self.constraints):
            for vid, cf in c.coeffs:
                col_id.append(int(vid))
                coef.append(cf)
            row_ptr[i + 1] = len(col_id)
        return dict(
            nvars=nv, ncons=nc, obj=obj, kind=kind, lb=lb, ub=ub, var_id=var_id, row_ptr=row_ptr,
            col_id=np.array(col_id, dtype=np.int64), coef=np.array(coef, dtype=np.float64),
            op=np.array([int(c.op) for c in self.constraints], dtype=np.uint8),
            rhs=np.array([c.rhs for c in self.constraints], dtype=np.float64),
        )

    def __str__(self) -> str:  # impl Display for Problem, src/problem.rs:305-351 (trailing blanks included)
        out = [f"{len(self.variables)} variables and {len(self.constraints)} constraints\n\n", "minimize\n"]
        by_id = {}
        for v in self.variables:
            assert v.id not in by_id  # "should have have repeated ids" (sic, :323)
            by_id[v.id] = v
            if v.obj_coeff == 0.0:
                continue
            out.append(f"{'+' if v.obj_coeff > 0.0 else '-'} {rust_f64(abs(v.obj_coeff))} {v} ")
        out.append("\n\nsubject to\n")
        sym = {ConstraintOp.Lte: "\u2264", ConstraintOp.Eq: "=", ConstraintOp.Gte: "\u2265"}
        for c in self.constraints:  # Constraint::display, :252-274
            for vid, cf in c.coeffs:
                if cf == 0.0:
                    continue
                out.append(f"{'+' if cf >= 0.0 else '-'} {rust_f64(abs(cf))} {by_id[vid]} ")
            out.append(f"{sym[c.op]} {rust_f64(c.rhs)}\n")
        out.append("\nwith the bounds\n")
        for v in self.variables:
            out.append(v.bound.display(v) + "\n")
        return "".join(out)
