"""``StandardForm`` of the reference as seen from Python (src/standard_form.rs:27-76, Display :223-237) and the nalgebra
``Display`` of vectors / matrices it relies on.  The arrays come from the product's native host layer
(``ellp_b200_stage_new(which = 0)`` = ``Option<StandardForm>::from(Problem)``, standard_form.rs:78-191); nothing here is on
the hot path -- it exists so that a user of the reference finds ``println!("{}", std_form)`` / ``println!("{}", sol.x())``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _native as N
from .problem import Bound, BoundKind, EllPError, Problem, rust_f64


def nalgebra_display(a) -> str:
    """``format!("{}", m)`` of a nalgebra DVector / DMatrix<f64> (nalgebra base/matrix.rs, impl_fmt!): a leading newline, a
    box of right-aligned entries, a trailing blank line; ``[ ]`` for an empty matrix."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    nrows, ncols = a.shape
    if nrows == 0 or ncols == 0:
        return "[ ]"
    cells = [[rust_f64(a[i, j]) for j in range(ncols)] for i in range(nrows)]
    width = max(len(s) for row in cells for s in row) + 1
    inner = " " * (width * ncols - 1)
    out = ["\n", f"  ┌ {inner} ┐\n"]
    for row in cells:
        out.append("  │")
        for s in row:
            out.append(" " + " " * (width - (len(s) + 1)) + s)
        out.append(" │\n")
    out.append(f"  └ {inner} ┘\n")
    out.append("\n")
    return "".join(out)


@dataclass
class StandardForm:
    c: np.ndarray
    A: np.ndarray           # m x n, column-major
    b: np.ndarray
    bounds: List[Bound]

    def rows(self) -> int:
        return self.A.shape[0]

    def cols(self) -> int:
        return self.A.shape[1]

    def obj(self, x) -> float:  # standard_form.rs:47-50
        return float(np.dot(self.c, np.asarray(x, dtype=np.float64)[: len(self.c)]))

    @staticmethod
    def from_problem(prob: Problem) -> Optional["StandardForm"]:
        """``Option<StandardForm>::from(Problem)``: None when a constraint without coefficients is infeasible."""
        desc, _keep = N.problem_desc(prob.to_arrays())
        h = C.c_void_p()
        infeasible = C.c_int(0)
        err = C.create_string_buffer(256)
        rc = N.lib.ellp_b200_stage_new(C.byref(desc), 0, C.byref(h), C.byref(infeasible), err)
        if rc != N.OK:
            raise EllPError(err.value.decode(errors="replace"))
        if infeasible.value or not h:
            return None
        try:
            dims = [C.c_int32() for _ in range(7)]
            N.lib.ellp_b200_stage_dims(h, *[C.byref(v) for v in dims])
            m, n, nx, nB, nN, lc, nb = [v.value for v in dims]
            A = np.zeros((m, n), order="F"); c = np.zeros(max(lc, 1)); b = np.zeros(max(m, 1))
            kind = np.zeros(max(nb, 1), dtype=np.uint8); lb = np.zeros(max(nb, 1)); ub = np.zeros(max(nb, 1))
            x = np.zeros(max(nx, 1)); B = np.zeros(max(nB, 1), dtype=np.int32); Nn = np.zeros(max(nN, 1), dtype=np.int32)
            Ns = np.zeros(max(nN, 1), dtype=np.uint8)
            N.lib.ellp_b200_stage_copy(h, N.ptr(A), N.ptr(c), N.ptr(b), N.ptr(kind), N.ptr(lb), N.ptr(ub), N.ptr(x), N.ptr(B), N.ptr(Nn),
                                       N.ptr(Ns), None, None)
            bounds = [Bound(BoundKind(int(kind[j])), float(lb[j]), float(ub[j])) for j in range(nb)]
            return StandardForm(c[:lc].copy(), A, b[:m].copy(), bounds)
        finally:
            N.lib.ellp_b200_stage_free(h)

    def __str__(self) -> str:  # impl Display for StandardForm, standard_form.rs:223-237
        out = [f"c:{nalgebra_display(self.c)}\n", f"A:{nalgebra_display(self.A)}\n", f"b:{nalgebra_display(self.b)}\n", "bounds:\n\n"]
        for i, bound in enumerate(self.bounds):
            out.append(f"x{i}: {bound}\n")
        return "".join(out)
