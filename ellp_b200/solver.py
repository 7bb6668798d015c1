"""GPU solvers with the method set of ellp's ``PrimalSimplexSolver`` / ``DualSimplexSolver``.

Reference API being mirrored: ``default()`` (max_iter 1000), ``new(Option<u64>)``, ``solve(Problem) ->
EllPResult`` (src/solvers/primal/primal_simplex_solver.rs:15-93, src/solvers/dual/dual_simplex_solver.rs:16-108),
``SolverResult::{Optimal(Solution), Infeasible, Unbounded, MaxIter{obj}}`` and ``Solution::{obj, x}``
(src/solver.rs:6-54).  Everything numerical happens behind the C ABI in libellp_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _native as N
from .problem import EllPError, Problem, rust_f64


class EllPPanic(RuntimeError):
    """A ``panic!``/``assert!`` site of the reference was reached; reported instead of aborting."""


@dataclass
class Solution:  # src/solver.rs:35-54
    _obj: float
    _x: np.ndarray

    def obj(self) -> float:
        return self._obj

    def x(self) -> np.ndarray:
        return self._x

    def x_display(self) -> str:
        """``format!("{}", sol.x())``: the nalgebra Display of the solution vector (see the example in src/lib.rs:45-59)."""
        from .standard_form import nalgebra_display
        return nalgebra_display(self._x)


@dataclass
class SolverResult:  # src/solver.rs:6-25
    kind: str  # "Optimal" | "Infeasible" | "Unbounded" | "MaxIter"
    solution: Optional[Solution] = None
    obj: Optional[float] = None  # MaxIter { obj }
    iters: List[int] = field(default_factory=lambda: [0, 0, 0, 0])
    used_primal_fallback: bool = False
    trace: Optional[np.ndarray] = field(default=None, repr=False)
    launches: int = 0
    ms_device: float = 0.0

    @property
    def is_optimal(self) -> bool:
        return self.kind == "Optimal"

    def __str__(self) -> str:  # src/solver.rs:14-25
        if self.kind == "Optimal":
            return f"found optimal point with objective {rust_f64(self.solution.obj())}"
        if self.kind == "Infeasible":
            return "problem is infeasible"
        if self.kind == "Unbounded":
            return "problem is unbounded"
        return f"reached max iterations, current objective = {rust_f64(self.obj)}"


_STATUS = {N.OPTIMAL: "Optimal", N.INFEASIBLE: "Infeasible", N.UNBOUNDED: "Unbounded", N.MAXITER: "MaxIter"}
_default_ctx: Optional[N.Context] = None


def default_context() -> N.Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = N.Context(0)
    return _default_ctx


def _raise(ctx: N.Context, rc: int):
    msg = N.lib.ellp_b200_last_error(ctx.h).decode(errors="replace")
    if rc == N.E_ELLP:
        raise EllPError(msg)
    if rc == N.E_PANIC:
        raise EllPPanic(msg)
    raise N.NativeError(rc, msg)


class _GpuSimplexSolver:
    _solver = N.PRIMAL

    def __init__(self, max_iter: Optional[int] = 1000, *, ctx: Optional[N.Context] = None,
                 tie_rule: int = N.TIES_REFERENCE, refactor_every: int = 0, check_every: int = 0, trace_cap: int = 0,
                 engine: int = N.ENGINE_AUTO, pricing: int = 0, ratio: int = 0, block_k: int = 0):
        self.max_iter = max_iter
        self.block_k = block_k
        self.pricing = pricing
        self.ratio = ratio
        self._ctx = ctx
        self.tie_rule = tie_rule
        self.refactor_every = refactor_every
        self.check_every = check_every
        self.trace_cap = trace_cap
        self.engine = engine

    @classmethod
    def default(cls, **kw):  # Default::default(): max_iter = 1000
        return cls(1000, **kw)

    @classmethod
    def new(cls, max_iter: Optional[int], **kw):  # new(None) => u64::MAX
        return cls(max_iter, **kw)

    @property
    def ctx(self) -> N.Context:
        return self._ctx or default_context()

    def _opts(self):
        o = N.default_opts(self.max_iter, self.tie_rule, self.refactor_every, self.check_every, engine=self.engine,
                           pricing=self.pricing, ratio=self.ratio, block_k=self.block_k)
        tr = None
        if self.trace_cap:
            tr = np.zeros(self.trace_cap, dtype=N.TRACE_DTYPE)
            o.trace = N.ptr(tr)
            o.trace_cap = self.trace_cap
        return o, tr

    def solve(self, prob: Problem) -> SolverResult:
        """``solver.solve(prob)`` -> SolverResult (raises EllPError where the reference returns Err)."""
        arr = prob.to_arrays()
        desc, _keep = N.problem_desc(arr)
        o, tr = self._opts()
        x = np.zeros(max(arr["nvars"], 1), dtype=np.float64)
        sol = N.Solution()
        sol.x = N.ptr(x)
        ctx = self.ctx
        rc = N.lib.ellp_b200_solve(ctx.h, C.byref(desc), self._solver, C.byref(o), C.byref(sol))
        if rc != N.OK:
            _raise(ctx, rc)
        kind = _STATUS[sol.status]
        res = SolverResult(kind, iters=list(sol.iters), used_primal_fallback=bool(sol.used_primal_fallback),
                           launches=int(sol.launches), ms_device=float(sol.ms_device))
        if tr is not None:
            res.trace = tr[: min(int(sol.trace_len), self.trace_cap)].copy()
        if kind == "Optimal":
            res.solution = Solution(float(sol.obj), x[: arr["nvars"]].copy())
        elif kind == "MaxIter":
            res.obj = float(sol.obj)
        return res

    def solve_with_initial(self, m, n, A, c, b, kind, lb, ub, x, B, Nv, N_side, y=None, d=None, profile=False):
        """The hot-path boundary on explicit arrays (updated in place).  Returns the native Result struct."""
        A = np.asfortranarray(A, dtype=np.float64)
        sf = N.StdForm(m, n, N.ptr(A), N.ptr(c), N.ptr(b), N.ptr(kind), N.ptr(lb), N.ptr(ub))
        pt = N.Point(N.ptr(x), N.ptr(B), N.ptr(Nv), N.ptr(N_side), N.ptr(y), N.ptr(d), len(B), len(Nv))
        o, tr = self._opts()
        o.profile = 1 if profile else 0
        o.phase_tag = 1 if self._solver == N.PRIMAL else 3
        res = N.Result()
        ctx = self.ctx
        fn = N.lib.ellp_b200_primal_solve_with_initial if self._solver == N.PRIMAL else N.lib.ellp_b200_dual_solve_with_initial
        rc = fn(ctx.h, C.byref(sf), C.byref(pt), C.byref(o), C.byref(res))
        if rc != N.OK:
            _raise(ctx, rc)
        trace = None if tr is None else tr[: min(int(res.trace_len), self.trace_cap)].copy()
        return res, trace


class GpuPrimalSimplexSolver(_GpuSimplexSolver):
    """Drop-in for ``ellp::PrimalSimplexSolver`` (primal_simplex_solver.rs:15-93), pivoting on the GPU."""
    _solver = N.PRIMAL


class GpuDualSimplexSolver(_GpuSimplexSolver):
    """Drop-in for ``ellp::DualSimplexSolver`` (dual_simplex_solver.rs:16-108), pivoting on the GPU."""
    _solver = N.DUAL


def parse_mps(text: str) -> Problem:
    """``ellp::parse_mps`` (src/parse_mps.rs:23-66) with deterministic file order; parsing is native (host C++)."""
    from .problem import Bound, BoundKind, ConstraintOp
    h = C.c_void_p()
    err = C.create_string_buffer(256)
    rc = N.lib.ellp_b200_parse_mps(text.encode(), C.byref(h), err)
    if rc != N.OK:
        raise EllPError(err.value.decode(errors="replace"))
    try:
        d = N.ProblemDesc()
        N.lib.ellp_b200_model_desc(h, C.byref(d))

        def arr(p, n, ty):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ty)), shape=(n,)).copy() if n else np.zeros(0)

        nv, nc = d.nvars, d.ncons
        obj, kind, lb, ub = arr(d.obj, nv, C.c_double), arr(d.kind, nv, C.c_uint8), arr(d.lb, nv, C.c_double), arr(d.ub, nv, C.c_double)
        rp = arr(d.row_ptr, nc + 1, C.c_int32)
        nnz = int(rp[-1]) if nc else 0
        col, coef = arr(d.col_id, nnz, C.c_int64), arr(d.coef, nnz, C.c_double)
        op, rhs = arr(d.op, nc, C.c_uint8), arr(d.rhs, nc, C.c_double)
        p = Problem.new()
        ids = []
        for j in range(nv):
            ids.append(p.add_var(float(obj[j]), Bound(BoundKind(int(kind[j])), float(lb[j]), float(ub[j]))))
        for r in range(nc):
            p.add_constraint([(ids[int(col[k])], float(coef[k])) for k in range(rp[r], rp[r + 1])], ConstraintOp(int(op[r])), float(rhs[r]))
        return p
    finally:
        N.lib.ellp_b200_model_free(h)


@dataclass
class BatchSolveResult:
    status: np.ndarray      # SolverResult codes per LP (0 Optimal, 1 Infeasible, 2 Unbounded, 3 MaxIter, -1 not solved)
    obj: np.ndarray
    x: np.ndarray           # (nlp, n + m) standard-form points (incl. artificial columns)
    iters: np.ndarray       # (nlp, 2) pivots of phase 1 / phase 2
    err: np.ndarray
    trace: Optional[np.ndarray]
    trace_len: Optional[np.ndarray]
    ms_device: float
    pivots: int


def primal_solve_batch(A, c, b, kind, lb, ub, max_iter: Optional[int] = 1000, tie_rule: int = N.TIES_REFERENCE,
                       trace_cap: int = 0, ctx: Optional[N.Context] = None) -> BatchSolveResult:
    """K6: PrimalSimplexSolver::solve (minus the standard-form step) for a batch of equally-shaped standard forms,
    one CTA per LP, one launch (include/ellp_b200.h: ellp_b200_primal_solve_batch).
    A: (nlp, m, n) with each LP column-major (i.e. pass np.asfortranarray per LP stacked) -- here given as (nlp, n, m)
    C-contiguous = per-LP column-major; c, kind, lb, ub: (nlp, n); b: (nlp, m)."""
    ctx = ctx or default_context()
    A = np.ascontiguousarray(A, dtype=np.float64)
    nlp, n, m = A.shape  # (nlp, n, m) C-order == column-major m x n per LP
    c = np.ascontiguousarray(c, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    kind = np.ascontiguousarray(kind, dtype=np.uint8); lb = np.ascontiguousarray(lb, dtype=np.float64); ub = np.ascontiguousarray(ub, dtype=np.float64)
    bt = N.Batch(nlp, m, n, N.ptr(A), N.ptr(c), N.ptr(b), N.ptr(kind), N.ptr(lb), N.ptr(ub))
    status = np.zeros(nlp, dtype=np.int32); obj = np.zeros(nlp); x = np.zeros((nlp, n + m)); iters = np.zeros((nlp, 2), dtype=np.int32)
    err = np.zeros(nlp, dtype=np.int32)
    tr = np.zeros((nlp, max(trace_cap, 1)), dtype=N.TRACE_DTYPE) if trace_cap else None
    tl = np.zeros(nlp, dtype=np.int32)
    res = N.BatchResult(N.ptr(status), N.ptr(obj), N.ptr(x), N.ptr(iters), N.ptr(err), N.ptr(tr) if trace_cap else None, trace_cap,
                        N.ptr(tl), 0.0, 0, 0)
    o = N.default_opts(max_iter, tie_rule)
    rc = N.lib.ellp_b200_primal_solve_batch(ctx.h, C.byref(bt), C.byref(o), C.byref(res))
    if rc != N.OK:
        _raise(ctx, rc)
    return BatchSolveResult(status, obj, x, iters, err, tr, tl, float(res.ms_device), int(res.pivots))
