"""Small invocations of every kernel family for compute-sanitizer.

    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py [case ...]
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py [case ...]

Cases (default: all): batch (k_batch_primal, 600 LPs 16x24), fused (k_blk_pivots_fused + k_blk_flush*, 1 rank, 1024x3072,
k = 8), fused_dual (k_blk_dual_pivots_fused, 1024x3072 dual, k = 8), flush (k_blk_flush / 3 / 4 on a 512 x 1024 matrix, k = 24 / 48),
lu (k_lu_panel_coop + triangular solves, m = 512), small (k_dual_small / k_gj_small on AFIRO), rank1 (k_rank1 513 x 70).
Every case checks its result so that a silently skipped kernel fails the run.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ellp_b200 import _native as N  # noqa: E402


def case_batch(ctx):
    nlp, m, ns = 600, 16, 24
    ctx.check(N.lib.ellp_b200_batch_generate(ctx.h, nlp, m, ns, 3, 0, 0))
    o = N.default_opts(None)
    res = N.BatchResult()
    ctx.check(N.lib.ellp_b200_batch_run(ctx.h, C.byref(o), C.byref(res)))
    status = np.zeros(nlp, dtype=np.int32)
    iters = np.zeros((nlp, 2), dtype=np.int32)
    err = np.zeros(nlp, dtype=np.int32)
    out = N.BatchResult(N.ptr(status), None, None, N.ptr(iters), N.ptr(err), None, 0, None, 0.0, 0, 0)
    ctx.check(N.lib.ellp_b200_batch_download(ctx.h, C.byref(out)))
    assert (err == 0).all() and (status == N.OPTIMAL).all(), (np.unique(status), np.unique(err))
    return f"{nlp} LPs, {out.pivots} pivots"


def _blocked(ctx, variant):
    m, ns, k, piv = 1024, 2048, 8, 64
    o = N.default_opts(piv, engine=N.ENGINE_TABLEAU, block_k=k, check_every=32)
    ctx.check(N.lib.ellp_b200_generate_dense_ex(ctx.h, m, ns, 3, variant, C.byref(o)))
    res = N.Result()
    ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
    assert res.iters == piv and res.status == N.MAXITER, (res.iters, res.status)
    return f"{res.iters} pivots, {res.launches} launches, obj {res.obj!r}"


def case_fused(ctx):
    return _blocked(ctx, 0)


def case_fused_dual(ctx):
    return _blocked(ctx, 1)


def case_flush(ctx):
    out = []
    rng = np.random.default_rng(1)
    R, Cc = 512, 1000
    for kern, k in ((1, 24), (3, 24), (4, 48), (3, 64), (5, 56), (8, 56), (9, 64), (9, 40)):
        ctx.set_tuning("flush_kernel", kern)
        E = np.asfortranarray(rng.standard_normal((R, Cc)))
        U = np.asfortranarray(rng.standard_normal((R, k)))
        V = np.ascontiguousarray(rng.standard_normal((k, Cc)))
        got = E.copy(order="F")
        ctx.check(N.lib.ellp_b200_rankk_update(ctx.h, N.ptr(got), R, Cc, R, N.ptr(U), N.ptr(V), k))
        np.testing.assert_allclose(got, E - U @ V, rtol=0, atol=1e-11)
        out.append(f"kernel {kern} k={k} ok")
    ctx.set_tuning("flush_kernel", 0)
    return ", ".join(out)


def case_lu(ctx):
    m = 512
    rng = np.random.default_rng(2)
    Bm = np.asfortranarray(rng.standard_normal((m, m)) + 3 * np.eye(m))
    inv = np.zeros((m, m), order="F")
    ctx.set_tuning("refactor_mode", 2)
    try:
        ctx.check(N.lib.ellp_b200_invert(ctx.h, N.ptr(Bm), m, N.ptr(inv)))
    finally:
        ctx.set_tuning("refactor_mode", 0)
    np.testing.assert_allclose(inv @ Bm, np.eye(m), atol=1e-8)
    return f"m={m} blocked LU inverse ok"


def case_small(ctx):
    import problems as P
    from ellp_b200.solver import GpuDualSimplexSolver, GpuPrimalSimplexSolver
    prob, exp = P.netlib("afiro")
    out = []
    for cls in (GpuPrimalSimplexSolver, GpuDualSimplexSolver):
        res = cls.default(ctx=ctx).solve(prob)
        P.check_expectation(exp, res.kind, res.solution.obj(), res.solution.x())
        out.append(f"{cls.__name__}: {res.iters} pivots, {res.launches} launches")
    return "; ".join(out)


def case_rank1(ctx):
    rng = np.random.default_rng(3)
    R, Cc, r = 513, 70, 512
    E = np.asfortranarray(rng.standard_normal((R, Cc)))
    alpha = rng.standard_normal(R)
    alpha[r] = 1.5
    got = E.copy(order="F")
    ctx.check(N.lib.ellp_b200_rank1_update(ctx.h, N.ptr(got), R, Cc, R, N.ptr(alpha), r))
    p = E[r, :] / alpha[r]
    want = E - np.outer(alpha, p)
    want[r, :] = p
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    return "513 x 70 ok"


CASES = {"batch": case_batch, "fused": case_fused, "fused_dual": case_fused_dual, "flush": case_flush, "lu": case_lu,
         "small": case_small, "rank1": case_rank1}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    ctx = N.Context(0)
    for name in names:
        print(f"[sanitize] {name}: {CASES[name](ctx)}", flush=True)
    ctx.close()
    print("[sanitize] all cases ok")
