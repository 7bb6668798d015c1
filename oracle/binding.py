"""ctypes binding of the CPU oracle (oracle/ellp_oracle.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under ellp_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libellp_oracle.so")

OPTIMAL, INFEASIBLE, UNBOUNDED, MAXITER = 0, 1, 2, 3
STATUS_NAMES = {0: "Optimal", 1: "Infeasible", 2: "Unbounded", 3: "MaxIter"}
MODE_EXACT, MODE_CANONICAL = 0, 1
PRIMAL, DUAL = 0, 1
U64_MAX = 2**64 - 1


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ only)."""
    src = os.path.join(_HERE, "ellp_oracle.cpp")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "ellp_oracle.h")))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class _Problem(C.Structure):
    _fields_ = [("nvars", C.c_int32), ("ncons", C.c_int32), ("obj", C.c_void_p), ("kind", C.c_void_p),
                ("lb", C.c_void_p), ("ub", C.c_void_p), ("var_id", C.c_void_p), ("row_ptr", C.c_void_p),
                ("col_id", C.c_void_p), ("coef", C.c_void_p), ("op", C.c_void_p), ("rhs", C.c_void_p)]


class _TraceRec(C.Structure):
    _fields_ = [("phase", C.c_int32), ("iter", C.c_int32), ("entering", C.c_int32), ("leaving", C.c_int32),
                ("step", C.c_double), ("obj", C.c_double)]


TRACE_DTYPE = np.dtype([("phase", "<i4"), ("iter", "<i4"), ("entering", "<i4"), ("leaving", "<i4"),
                        ("step", "<f8"), ("obj", "<f8")])


class _Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("obj", C.c_double), ("x", C.c_void_p), ("iters", C.c_uint64 * 4),
                ("used_primal_fallback", C.c_int32), ("trace", C.c_void_p), ("trace_cap", C.c_int64),
                ("trace_len", C.c_int64), ("err", C.c_char * 256)]


class _StdForm(C.Structure):
    _fields_ = [("m", C.c_int32), ("n", C.c_int32), ("A", C.c_void_p), ("c", C.c_void_p), ("b", C.c_void_p),
                ("kind", C.c_void_p), ("lb", C.c_void_p), ("ub", C.c_void_p)]


class _Point(C.Structure):
    _fields_ = [("x", C.c_void_p), ("B", C.c_void_p), ("N", C.c_void_p), ("N_side", C.c_void_p),
                ("y", C.c_void_p), ("d", C.c_void_p), ("nB", C.c_int32), ("nN", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ellp_oracle_solve.restype = C.c_int
        _lib.ellp_oracle_solve.argtypes = [C.POINTER(_Problem), C.c_int, C.c_uint64, C.c_int, C.POINTER(_Result)]
        for fn in (_lib.ellp_oracle_primal_solve_with_initial, _lib.ellp_oracle_dual_solve_with_initial):
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(_StdForm), C.POINTER(_Point), C.c_uint64, C.c_int, C.POINTER(_Result)]
        _lib.ellp_oracle_stage_new.restype = C.c_void_p
        _lib.ellp_oracle_stage_new.argtypes = [C.POINTER(_Problem), C.c_int, C.POINTER(C.c_int), C.c_char_p]
        _lib.ellp_oracle_stage_free.argtypes = [C.c_void_p]
        _lib.ellp_oracle_stage_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 5
        _lib.ellp_oracle_stage_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 14
        _lib.ellp_oracle_rank1_update.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                                  C.c_void_p, C.c_int64]
        _lib.ellp_oracle_gemv_t.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        _lib.ellp_oracle_version.restype = C.c_char_p
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleError(Exception):
    def __init__(self, code: int, msg: str):
        super().__init__(f"oracle rc={code}: {msg}")
        self.code, self.msg = code, msg


@dataclass
class OracleResult:
    status: int
    obj: float
    x: np.ndarray
    iters: List[int]
    used_primal_fallback: bool
    trace: np.ndarray = field(repr=False, default=None)

    @property
    def status_name(self) -> str:
        return STATUS_NAMES[self.status]


def _c_problem(arr: dict):
    keep = dict(arr)  # keep numpy arrays alive
    p = _Problem(arr["nvars"], arr["ncons"], _ptr(arr["obj"]), _ptr(arr["kind"]), _ptr(arr["lb"]), _ptr(arr["ub"]),
                 _ptr(arr["var_id"]), _ptr(arr["row_ptr"]), _ptr(arr["col_id"]), _ptr(arr["coef"]), _ptr(arr["op"]),
                 _ptr(arr["rhs"]))
    return p, keep


def solve(problem, solver: int, max_iter: Optional[int] = 1000, mode: int = MODE_EXACT,
          trace_cap: int = 0) -> OracleResult:
    """{Primal,Dual}SimplexSolver::solve on ``problem`` (an ellp_b200.problem.Problem or its to_arrays())."""
    arr = problem if isinstance(problem, dict) else problem.to_arrays()
    p, _keep = _c_problem(arr)
    x = np.zeros(max(arr["nvars"], 1), dtype=np.float64)
    tr = np.zeros(max(trace_cap, 1), dtype=TRACE_DTYPE)
    res = _Result()
    res.x = _ptr(x)
    res.trace = _ptr(tr) if trace_cap else None
    res.trace_cap = trace_cap
    rc = lib().ellp_oracle_solve(C.byref(p), solver, U64_MAX if max_iter is None else max_iter, mode, C.byref(res))
    if rc != 0:
        raise OracleError(rc, res.err.decode())
    return OracleResult(res.status, res.obj, x[: arr["nvars"]].copy(), list(res.iters), bool(res.used_primal_fallback),
                        tr[: min(res.trace_len, trace_cap)].copy())


def solve_with_initial(kind: int, m: int, n: int, A, c, b, bkind, lb, ub, x, B, N, N_side, y=None, d=None,
                       max_iter: Optional[int] = 1000, mode: int = MODE_EXACT, trace_cap: int = 0):
    """solve_with_initial on an explicit standard form + basic point.  Arrays x/B/N/N_side/y/d are updated in place."""
    A = np.asfortranarray(A, dtype=np.float64)
    sf = _StdForm(m, n, _ptr(A), _ptr(c), _ptr(b), _ptr(bkind), _ptr(lb), _ptr(ub))
    pt = _Point(_ptr(x), _ptr(B), _ptr(N), _ptr(N_side), _ptr(y), _ptr(d), len(B), len(N))
    tr = np.zeros(max(trace_cap, 1), dtype=TRACE_DTYPE)
    res = _Result()
    res.trace = _ptr(tr) if trace_cap else None
    res.trace_cap = trace_cap
    fn = lib().ellp_oracle_primal_solve_with_initial if kind == PRIMAL else lib().ellp_oracle_dual_solve_with_initial
    rc = fn(C.byref(sf), C.byref(pt), U64_MAX if max_iter is None else max_iter, mode, C.byref(res))
    if rc != 0:
        raise OracleError(rc, res.err.decode())
    return OracleResult(res.status, res.obj, x, list(res.iters), False, tr[: min(res.trace_len, trace_cap)].copy())


@dataclass
class Stage:
    m: int
    n: int
    A: np.ndarray
    c: np.ndarray
    b: np.ndarray
    kind: np.ndarray
    lb: np.ndarray
    ub: np.ndarray
    x: np.ndarray
    B: np.ndarray
    N: np.ndarray
    N_side: np.ndarray
    y: np.ndarray
    d: np.ndarray


def stage(problem, which: int) -> Optional[Stage]:
    """which: 0 standard form, 1 primal phase 1, 2 dual phase 1.  None when the reference returns None (Infeasible)."""
    arr = problem if isinstance(problem, dict) else problem.to_arrays()
    p, _keep = _c_problem(arr)
    infeasible = C.c_int(0)
    err = C.create_string_buffer(256)
    h = lib().ellp_oracle_stage_new(C.byref(p), which, C.byref(infeasible), err)
    if not h:
        if infeasible.value:
            return None
        raise OracleError(-2, err.value.decode())
    try:
        dims = [C.c_int32() for _ in range(5)]
        lib().ellp_oracle_stage_dims(h, *[C.byref(v) for v in dims])
        m, n, nx, nB, nN = [v.value for v in dims]
        cap = n + m + 8
        A = np.zeros((m, n), dtype=np.float64, order="F")
        c = np.zeros(cap); b = np.zeros(max(m, 1)); kind = np.zeros(cap, dtype=np.uint8)
        lb = np.zeros(cap); ub = np.zeros(cap); x = np.zeros(max(nx, 1))
        B = np.zeros(max(nB, 1), dtype=np.int32); N = np.zeros(max(nN, 1), dtype=np.int32)
        Ns = np.zeros(max(nN, 1), dtype=np.uint8); y = np.zeros(max(m, 1)); d = np.zeros(cap)
        len_c, len_b = C.c_int32(), C.c_int32()
        lib().ellp_oracle_stage_copy(h, _ptr(A), _ptr(c), C.cast(C.byref(len_c), C.c_void_p), _ptr(b), _ptr(kind),
                                     _ptr(lb), _ptr(ub), C.cast(C.byref(len_b), C.c_void_p), _ptr(x), _ptr(B), _ptr(N),
                                     _ptr(Ns), _ptr(y), _ptr(d))
        return Stage(m, n, A, c[: len_c.value], b[:m], kind[: len_b.value], lb[: len_b.value], ub[: len_b.value],
                     x[:nx], B[:nB], N[:nN], Ns[:nN], y[:m] if which == 2 else y[:0], d[:n] if which == 2 else d[:0])
    finally:
        lib().ellp_oracle_stage_free(h)


def rank1_update(E: np.ndarray, alpha: np.ndarray, rho: np.ndarray, r: int) -> None:
    assert E.flags.f_contiguous and E.dtype == np.float64
    R, Cc = E.shape
    lib().ellp_oracle_rank1_update(_ptr(E), R, Cc, R, _ptr(alpha), _ptr(rho), r)


def gemv_t(M: np.ndarray, v: np.ndarray) -> np.ndarray:
    assert M.flags.f_contiguous and M.dtype == np.float64
    R, Cc = M.shape
    y = np.zeros(Cc)
    lib().ellp_oracle_gemv_t(_ptr(M), R, Cc, R, _ptr(v), _ptr(y))
    return y
