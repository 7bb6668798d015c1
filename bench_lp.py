"""Numpy twin of the device LP generator (ellp_b200/csrc/kernels.cuh: k_gen_dense_cols / k_gen_dense_vectors).

Used by tests (bit-for-bit check against the device) and by bench.py's CPU legs, which must not touch the GPU library.
    min -c.x  s.t.  A x + s = b, x, s >= 0;   A ~ U(0,1) m x ns,  b_i ~ U(1,2) * ns/4,  c_j ~ U(0.5,1.5)
(SURVEY.md 8(d), configs 4/5).  Starting point: slack basis.
"""
import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _uniform01(seed, idx):
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(seed) * np.uint64(0x2545F4914F6CDD1D) + idx.astype(np.uint64))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def dense_lp(m: int, ns: int, seed: int, variant: int = 0) -> dict:
    """variant 0: min -c.x, Ax + s = b (primal-feasible slack basis); variant 1: min c.x, Ax - s = b (dual-feasible)."""
    if variant == 1:
        lp = dense_lp(m, ns, seed, 0)
        lp["A"][:, ns:] = 0.0  # off-diagonal zeros are +0.0 on the device, so do not negate an identity
        np.fill_diagonal(lp["A"][:, ns:], -1.0)
        lp["c"] = -lp["c"]
        lp["c"][ns:] = 0.0
        lp["x"][ns:] = -lp["b"]
        lp["y"] = np.zeros(m)
        lp["d"] = lp["c"].copy()
        return lp
    n = ns + m
    A = np.zeros((m, n), order="F")
    idx = np.arange(m * ns, dtype=np.uint64)
    A[:, :ns] = _uniform01(seed, idx).reshape((m, ns), order="F")
    A[:, ns:] = np.eye(m)
    c = np.zeros(n)
    c[:ns] = -(0.5 + _uniform01(seed + 1, np.arange(ns, dtype=np.uint64)))
    b = (1.0 + _uniform01(seed + 2, np.arange(m, dtype=np.uint64))) * (ns * 0.25)
    x = np.zeros(n)
    x[ns:] = b
    return dict(m=m, n=n, A=A, c=c, b=b, kind=np.ones(n, dtype=np.uint8), lb=np.zeros(n), ub=np.zeros(n), x=x,
                B=np.arange(ns, n, dtype=np.int32), N=np.arange(ns, dtype=np.int32), N_side=np.zeros(ns, dtype=np.uint8))


def batch_lp(m: int, ns: int, seed: int, lpid: int) -> dict:
    """Numpy twin of k_gen_batch (ellp_b200/csrc/batch.cuh): LP `lpid` of the configs[3] batch,
    min -c.x, A x <= b, x >= 0.  Returns the structural data (A m x ns, c ns, b m)."""
    idx = np.arange(m * ns, dtype=np.uint64)
    A = _uniform01(seed + 3 * lpid, idx).reshape((m, ns), order="F")
    c = -(0.5 + _uniform01(seed + 3 * lpid + 1, np.arange(ns, dtype=np.uint64)))
    b = (1.0 + _uniform01(seed + 3 * lpid + 2, np.arange(m, dtype=np.uint64))) * (ns * 0.25)
    return dict(A=A, c=c, b=b)


def batch_lp_problem_arrays(m: int, ns: int, seed: int, lpid: int) -> dict:
    """The same LP in the flat layout of Problem.to_arrays() (dense CSR rows, every variable Lower(0), rows Lte)."""
    lp = batch_lp(m, ns, seed, lpid)
    return dict(nvars=ns, ncons=m, obj=lp["c"].copy(), kind=np.ones(ns, dtype=np.uint8), lb=np.zeros(ns), ub=np.zeros(ns),
                var_id=np.arange(ns, dtype=np.int64), row_ptr=(np.arange(m + 1, dtype=np.int32) * ns),
                col_id=np.tile(np.arange(ns, dtype=np.int64), m), coef=np.ascontiguousarray(lp["A"]).reshape(-1),
                op=np.zeros(m, dtype=np.uint8), rhs=lp["b"].copy())
