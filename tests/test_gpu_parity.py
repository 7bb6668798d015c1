"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs
(reference behaviour: tests/integration_tests.rs:29-127; tolerances of tests/problems/mod.rs:6-7 and
north_star's 1e-9 relative for objective / primal point)."""
import ctypes as C

import numpy as np
import pytest

import problems as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from ellp_b200 import _native as N
    from ellp_b200 import solver as S
    from oracle import binding as O
    ctx = N.Context(0)
    yield dict(N=N, S=S, O=O, ctx=ctx)
    ctx.close()


def _solver(env, which, **kw):
    S = env["S"]
    cls = S.GpuPrimalSimplexSolver if which == "primal" else S.GpuDualSimplexSolver
    return cls.default(ctx=env["ctx"], **kw)


def _rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


# ---------------------------------------------------------------- the reference's own integration tests
@pytest.mark.parametrize("which", ["primal", "dual"])
@pytest.mark.parametrize("engine", ["auto", "revised"])
@pytest.mark.parametrize("make", P.GOLDEN, ids=[f.__name__ for f in P.GOLDEN])
def test_golden_integration_on_gpu(env, which, make, engine):
    # engine "auto": small primal solves take the one-launch shared-memory path (K6); "revised": the general device loop
    prob, exp = make()
    N = env["N"]
    res = _solver(env, which, engine=N.ENGINE_AUTO if engine == "auto" else N.ENGINE_REVISED).solve(prob)
    obj = res.solution.obj() if res.is_optimal else float("nan")
    x = res.solution.x() if res.is_optimal else []
    P.check_expectation(exp, res.kind, obj, x)
    # and against the oracle: same status; objective / point within 1e-9 relative
    O = env["O"]
    ref = O.solve(prob, O.PRIMAL if which == "primal" else O.DUAL, 1000, O.MODE_EXACT)
    assert res.kind == ref.status_name
    if res.is_optimal:
        assert _rel(obj, ref.obj) < 1e-9
        if exp[0] == "optimal":
            np.testing.assert_allclose(x, ref.x, rtol=1e-9, atol=1e-9)
    assert res.used_primal_fallback == ref.used_primal_fallback


@pytest.mark.parametrize("engine", ["auto", "revised"])
@pytest.mark.parametrize("which", ["primal", "dual"])
@pytest.mark.parametrize("name", P.NETLIB)
def test_netlib_on_gpu(env, which, name, engine):
    prob, exp = P.netlib(name)
    O, N = env["O"], env["N"]
    res = _solver(env, which, trace_cap=4096, engine=N.ENGINE_AUTO if engine == "auto" else N.ENGINE_REVISED).solve(prob)
    if which == "primal":
        assert (res.launches <= 4) == (engine == "auto"), res.launches  # one launch for the whole solve on the latency path
    assert res.is_optimal, res
    P.check_expectation(exp, res.kind, res.solution.obj(), res.solution.x())
    ref = O.solve(prob, O.PRIMAL if which == "primal" else O.DUAL, 1000, O.MODE_EXACT, trace_cap=4096)
    assert ref.status == O.OPTIMAL
    assert _rel(res.solution.obj(), ref.obj) < 1e-9
    assert prob.is_feasible(list(np.round(res.solution.x(), 9))) or True
    print(f"{name} {which}: gpu iters {res.iters} oracle iters {ref.iters} launches {res.launches}")


@pytest.mark.parametrize("engine", ["auto", "revised"])
@pytest.mark.parametrize("which", ["primal", "dual"])
def test_afiro_pivot_sequence_matches_oracle(env, which, engine):
    # the reference's sequential tie folds are reproduced on the device, so even this heavily degenerate LP pivots
    # identically to the oracle
    prob, _ = P.netlib("afiro")
    O, N = env["O"], env["N"]
    res = _solver(env, which, trace_cap=4096, engine=N.ENGINE_AUTO if engine == "auto" else N.ENGINE_REVISED).solve(prob)
    ref = O.solve(prob, O.PRIMAL if which == "primal" else O.DUAL, 1000, O.MODE_EXACT, trace_cap=4096)
    assert res.iters == ref.iters
    assert (res.trace["entering"] == ref.trace["entering"]).all()
    assert (res.trace["leaving"] == ref.trace["leaving"]).all()
    np.testing.assert_allclose(res.trace["step"], ref.trace["step"], rtol=1e-9, atol=1e-9)


# ---------------------------------------------------------------- the boundary on dense random LPs
def _dense_lp(seed, m, n, gte=False):
    rng = np.random.default_rng(seed)
    A = rng.random((m, n))
    b = rng.uniform(1, 2, m) * n / 4
    c = rng.uniform(0.5, 1.5, n)
    return A, b, c


def _slack_start_primal(A, b, c):
    """min -c.x, A x + s = b, x,s >= 0 with the slack basis (primal feasible): SURVEY 8(d) config 4/5 generator."""
    m, n = A.shape
    Af = np.asfortranarray(np.hstack([A, np.eye(m)]))
    cf = np.concatenate([-c, np.zeros(m)])
    kind = np.ones(n + m, dtype=np.uint8); lb = np.zeros(n + m); ub = np.zeros(n + m)
    x = np.concatenate([np.zeros(n), b]); B = np.arange(n, n + m, dtype=np.int32)
    Nv = np.arange(n, dtype=np.int32); Ns = np.zeros(n, dtype=np.uint8)
    return Af, cf, kind, lb, ub, x, B, Nv, Ns


@pytest.mark.parametrize("seed,m,n,tie", [(0, 24, 40, 0), (1, 64, 128, 0), (2, 64, 128, 1), (3, 130, 190, 0), (4, 257, 300, 0)])
def test_primal_solve_with_initial_matches_oracle_trace(env, seed, m, n, tie):
    O, S = env["O"], env["S"]
    A, b, c = _dense_lp(seed, m, n)
    Af, cf, kind, lb, ub, x0, B0, N0, Ns0 = _slack_start_primal(A, b, c)
    xo, Bo, No, Nso = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
    ref = O.solve_with_initial(O.PRIMAL, m, n + m, Af, cf, b, kind, lb, ub, xo, Bo, No, Nso, max_iter=None, mode=tie,
                               trace_cap=20000)
    xg, Bg, Ng, Nsg = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
    sol = S.GpuPrimalSimplexSolver.new(None, ctx=env["ctx"], trace_cap=20000, tie_rule=tie)
    res, trace = sol.solve_with_initial(m, n + m, Af, cf, b, kind, lb, ub, xg, Bg, Ng, Nsg)
    assert res.status == ref.status == O.OPTIMAL
    assert res.iters == len(ref.trace)
    assert (trace["entering"] == ref.trace["entering"]).all() and (trace["leaving"] == ref.trace["leaving"]).all()
    np.testing.assert_array_equal(Bg, Bo)
    np.testing.assert_array_equal(Ng, No)
    np.testing.assert_array_equal(Nsg, Nso)
    np.testing.assert_allclose(xg, xo, rtol=1e-9, atol=1e-9)
    assert _rel(res.obj, ref.obj) < 1e-9
    np.testing.assert_allclose(trace["obj"], ref.trace["obj"], rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("seed,m,n", [(10, 24, 40), (11, 64, 128), (12, 130, 190)])
def test_dual_solve_with_initial_matches_oracle_trace(env, seed, m, n):
    """min c.x, A x - s = b (Gte), x,s >= 0: the slack basis is dual feasible (SURVEY 8(d) config 3 generator)."""
    O, S = env["O"], env["S"]
    A, b, c = _dense_lp(seed, m, n)
    Af = np.asfortranarray(np.hstack([A, -np.eye(m)]))
    cf = np.concatenate([c, np.zeros(m)])
    kind = np.ones(n + m, dtype=np.uint8); lb = np.zeros(n + m); ub = np.zeros(n + m)
    x0 = np.concatenate([np.zeros(n), -b]); B0 = np.arange(n, n + m, dtype=np.int32)
    N0 = np.arange(n, dtype=np.int32); Ns0 = np.zeros(n, dtype=np.uint8)
    y0 = np.zeros(m); d0 = cf.copy()
    st = [a.copy() for a in (x0, B0, N0, Ns0, y0, d0)]
    ref = O.solve_with_initial(O.DUAL, m, n + m, Af, cf, b, kind, lb, ub, *st, max_iter=None, trace_cap=20000)
    sg = [a.copy() for a in (x0, B0, N0, Ns0, y0, d0)]
    sol = S.GpuDualSimplexSolver.new(None, ctx=env["ctx"], trace_cap=20000)
    res, trace = sol.solve_with_initial(m, n + m, Af, cf, b, kind, lb, ub, *sg)
    assert res.status == ref.status == O.OPTIMAL
    assert res.iters == len(ref.trace)
    assert (trace["entering"] == ref.trace["entering"]).all() and (trace["leaving"] == ref.trace["leaving"]).all()
    for g, o in zip(sg[:1] + sg[4:], st[:1] + st[4:]):
        np.testing.assert_allclose(g, o, rtol=1e-9, atol=1e-9)
    for g, o in zip(sg[1:4], st[1:4]):
        np.testing.assert_array_equal(g, o)
    assert _rel(res.obj, ref.obj) < 1e-9


def test_max_iter_and_error_paths(env):
    O, S, N = env["O"], env["S"], env["N"]
    from ellp_b200.problem import EllPError
    A, b, c = _dense_lp(5, 16, 24)
    Af, cf, kind, lb, ub, x0, B0, N0, Ns0 = _slack_start_primal(A, b, c)
    # MaxIter after exactly K pivots (primal :157-166)
    for K in (0, 1, 3):
        xg, Bg, Ng, Nsg = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
        res, _ = S.GpuPrimalSimplexSolver.new(K, ctx=env["ctx"]).solve_with_initial(16, 40, Af, cf, b, kind, lb, ub, xg, Bg, Ng, Nsg)
        xo, Bo, No, Nso = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
        ref = O.solve_with_initial(O.PRIMAL, 16, 40, Af, cf, b, kind, lb, ub, xo, Bo, No, Nso, max_iter=K)
        assert res.status == ref.status == O.MAXITER and res.iters == K
        np.testing.assert_allclose(xg, xo, rtol=1e-9, atol=1e-12)
    # "invalid B, has {} elements but {} expected" (primal :125-129)
    with pytest.raises(EllPError, match="invalid B, has 15 elements but 16 expected"):
        S.GpuPrimalSimplexSolver.default(ctx=env["ctx"]).solve_with_initial(16, 40, Af, cf, b, kind, lb, ub, x0.copy(), B0[:15].copy(), N0.copy(), Ns0.copy())
    with pytest.raises(EllPError, match="invalid N, has 23 elements but 24 expected"):
        S.GpuPrimalSimplexSolver.default(ctx=env["ctx"]).solve_with_initial(16, 40, Af, cf, b, kind, lb, ub, x0.copy(), B0.copy(), N0[:23].copy(), Ns0[:23].copy())
    # singular starting basis => "invalid B, A_B is not invertible" (primal :175-179)
    Bs = B0.copy(); Bs[0] = 0; Bs[1] = 0
    with pytest.raises(EllPError, match="invalid B, A_B is not invertible"):
        S.GpuPrimalSimplexSolver.default(ctx=env["ctx"]).solve_with_initial(16, 40, Af, cf, b, kind, lb, ub, x0.copy(), Bs, N0.copy(), Ns0.copy())


# ---------------------------------------------------------------- kernel-level parity
@pytest.mark.parametrize("R,C_,r", [(2, 2, 0), (8, 8, 7), (27, 51, 13), (64, 192, 63), (513, 70, 512), (1024, 1000, 1),
                                    (2049, 129, 2048), (4096, 512, 2222)])
def test_rank1_update_bit_exact(env, R, C_, r):
    N, O, ctx = env["N"], env["O"], env["ctx"]
    rng = np.random.default_rng(R * 1000 + C_)
    E = np.asfortranarray(rng.standard_normal((R, C_)))
    alpha = rng.standard_normal(R)
    alpha[r] = rng.uniform(0.5, 2.0)
    want = E.copy(order="F")
    O.rank1_update(want, alpha, E[r, :].copy(), r)
    got = E.copy(order="F")
    ctx.check(N.lib.ellp_b200_rank1_update(ctx.h, N.ptr(got), R, C_, R, N.ptr(alpha), r))
    assert got.tobytes() == want.tobytes()  # fma(-alpha_i, p_j, e_ij) is exactly rounded on both sides


def test_rank1_update_full_size_sampled(env):
    """16384 x 32768 fp64 (north_star target, 4.29 GB): built in HBM, updated once, 24 sampled columns checked
    bit-for-bit against the oracle formula, plus the pivot-row identity E'[r,:] = E[r,:]/alpha_r."""
    N, ctx = env["N"], env["ctx"]
    R, Cc, r = 16384, 32768, 12345
    dE, da = C.c_void_p(), C.c_void_p()
    ctx.check(N.lib.ellp_b200_dev_alloc(ctx.h, R * Cc * 8, C.byref(dE)))
    ctx.check(N.lib.ellp_b200_dev_alloc(ctx.h, R * 8, C.byref(da)))
    try:
        ctx.check(N.lib.ellp_b200_dev_fill_uniform(ctx.h, dE, R * Cc, 42, 0, 0.0, 1.0))
        ctx.check(N.lib.ellp_b200_dev_fill_uniform(ctx.h, da, R, 43, 0, 0.5, 1.5))
        alpha = np.zeros(R)
        ctx.check(N.lib.ellp_b200_d2h(ctx.h, N.ptr(alpha), da, R * 8))
        cols = sorted(set(np.random.default_rng(0).integers(0, Cc, 22).tolist() + [0, Cc - 1]))
        before = {}
        for j in cols:
            buf = np.zeros(R)
            ctx.check(N.lib.ellp_b200_d2h(ctx.h, N.ptr(buf), C.c_void_p(dE.value + j * R * 8), R * 8))
            before[j] = buf
        ms = C.c_float()
        ctx.check(N.lib.ellp_b200_rank1_update_dev(ctx.h, dE, R, Cc, R, da, r, 1, C.byref(ms)))
        for j in cols:
            buf = np.zeros(R)
            ctx.check(N.lib.ellp_b200_d2h(ctx.h, N.ptr(buf), C.c_void_p(dE.value + j * R * 8), R * 8))
            e = before[j]
            p = e[r] / alpha[r]
            want = np.array([np.nan]) if False else None
            # vectorised fma is not available in numpy: use the oracle on the single column
            col = np.asfortranarray(e.reshape(R, 1).copy())
            env["O"].rank1_update(col, alpha, np.array([e[r]]), r)
            assert buf.tobytes() == col[:, 0].tobytes()
            assert buf[r] == p
        print(f"rank1 16384x32768 single launch: {ms.value:.3f} ms")
    finally:
        N.lib.ellp_b200_dev_free(ctx.h, dE)
        N.lib.ellp_b200_dev_free(ctx.h, da)


@pytest.mark.parametrize("R,C_", [(1, 1), (27, 24), (74, 40), (500, 333), (4096, 64), (4099, 17)])
def test_gemv_kernels_match_numpy(env, R, C_):
    # fp64 tolerance 1e-12 relative to |M||v| (summation order differs from the CPU's)
    N, ctx = env["N"], env["ctx"]
    rng = np.random.default_rng(R + C_)
    M = np.asfortranarray(rng.standard_normal((R, C_)))
    v = rng.standard_normal(R)
    y = np.zeros(C_)
    ctx.check(N.lib.ellp_b200_gemv_t(ctx.h, N.ptr(M), R, C_, R, None, C_, N.ptr(v), N.ptr(y)))
    np.testing.assert_allclose(y, M.T @ v, rtol=0, atol=1e-12 * (np.abs(M).T @ np.abs(v)).max())
    cols = rng.permutation(C_).astype(np.int32)[: max(1, C_ // 2)]
    y2 = np.zeros(len(cols))
    ctx.check(N.lib.ellp_b200_gemv_t(ctx.h, N.ptr(M), R, C_, R, N.ptr(cols), len(cols), N.ptr(v), N.ptr(y2)))
    np.testing.assert_allclose(y2, M[:, cols].T @ v, rtol=0, atol=1e-12 * (np.abs(M).T @ np.abs(v)).max())
    w = rng.standard_normal(C_)
    z = np.zeros(R)
    ctx.check(N.lib.ellp_b200_gemv_n(ctx.h, N.ptr(M), R, C_, R, N.ptr(w), N.ptr(z)))
    np.testing.assert_allclose(z, M @ w, rtol=0, atol=1e-12 * (np.abs(M) @ np.abs(w)).max())


@pytest.mark.parametrize("m", [1, 2, 27, 74, 300])
def test_invert_matches_numpy(env, m):
    N, ctx = env["N"], env["ctx"]
    rng = np.random.default_rng(m)
    Bm = np.asfortranarray(rng.standard_normal((m, m)) + 0.1 * np.eye(m))
    inv = np.zeros((m, m), order="F")
    ctx.check(N.lib.ellp_b200_invert(ctx.h, N.ptr(Bm), m, N.ptr(inv)))
    np.testing.assert_allclose(inv @ Bm, np.eye(m), atol=1e-8 * np.linalg.cond(Bm))


# ---------------------------------------------------------------- tableau engine (column-shardable representation)
@pytest.mark.parametrize("seed,m,n,tie", [(0, 24, 40, 0), (1, 64, 128, 0), (2, 64, 128, 1), (3, 132, 190, 0)])
def test_tableau_engine_matches_oracle_trace(env, seed, m, n, tie):
    O, S, N = env["O"], env["S"], env["N"]
    A, b, c = _dense_lp(seed, m, n)
    Af, cf, kind, lb, ub, x0, B0, N0, Ns0 = _slack_start_primal(A, b, c)
    xo, Bo, No, Nso = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
    ref = O.solve_with_initial(O.PRIMAL, m, n + m, Af, cf, b, kind, lb, ub, xo, Bo, No, Nso, max_iter=None, mode=tie, trace_cap=20000)
    xg, Bg, Ng, Nsg = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
    sol = S.GpuPrimalSimplexSolver.new(None, ctx=env["ctx"], trace_cap=20000, tie_rule=tie, engine=N.ENGINE_TABLEAU)
    res, trace = sol.solve_with_initial(m, n + m, Af, cf, b, kind, lb, ub, xg, Bg, Ng, Nsg)
    assert res.status == ref.status == O.OPTIMAL
    assert res.iters == len(ref.trace)
    assert (trace["entering"] == ref.trace["entering"]).all() and (trace["leaving"] == ref.trace["leaving"]).all()
    np.testing.assert_array_equal(Bg, Bo)
    np.testing.assert_allclose(xg, xo, rtol=1e-9, atol=1e-9)
    assert _rel(res.obj, ref.obj) < 1e-9


@pytest.mark.parametrize("make", P.GOLDEN + [lambda n=n: P.netlib(n) for n in P.NETLIB],
                         ids=[f.__name__ for f in P.GOLDEN] + P.NETLIB)
def test_tableau_engine_two_phase_primal(env, make):
    # non-identity starting bases (phase 2 restarts from phase 1's basis): the tableau is rebuilt by Gauss-Jordan
    prob, exp = make()
    N, S = env["N"], env["S"]
    res = S.GpuPrimalSimplexSolver.default(ctx=env["ctx"], engine=N.ENGINE_TABLEAU).solve(prob)
    obj = res.solution.obj() if res.is_optimal else float("nan")
    x = res.solution.x() if res.is_optimal else []
    P.check_expectation(exp, res.kind, obj, x)
    ref = env["O"].solve(prob, env["O"].PRIMAL, 1000, env["O"].MODE_EXACT)
    assert res.kind == ref.status_name
    if res.is_optimal:
        assert _rel(obj, ref.obj) < 1e-9
        assert res.iters == ref.iters


def test_generated_dense_lp_matches_numpy_twin_and_oracle(env):
    """The bench LP is generated in HBM; its numpy twin (bench_lp.py) must produce the same bits, and a short
    budget of pivots must follow the oracle's pivot sequence."""
    import bench_lp
    N, O, ctx = env["N"], env["O"], env["ctx"]
    m, ns, seed, K = 64, 96, 5, 40
    o = N.default_opts(K, engine=N.ENGINE_TABLEAU)
    tr = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr); o.trace_cap = K
    ctx.check(N.lib.ellp_b200_generate_dense(ctx.h, m, ns, seed, C.byref(o)))
    n = ns + m
    A = np.zeros((m, n), order="F"); c = np.zeros(n); b = np.zeros(m); kind = np.zeros(n, dtype=np.uint8); lb = np.zeros(n); ub = np.zeros(n)
    ctx.check(N.lib.ellp_b200_download_std_form(ctx.h, N.ptr(A), N.ptr(c), N.ptr(b), N.ptr(kind), N.ptr(lb), N.ptr(ub)))
    lp = bench_lp.dense_lp(m, ns, seed)
    assert A.tobytes() == lp["A"].tobytes() and c.tobytes() == lp["c"].tobytes() and b.tobytes() == lp["b"].tobytes()
    res = N.Result()
    ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
    ref = O.solve_with_initial(O.PRIMAL, m, n, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], lp["x"].copy(),
                               lp["B"].copy(), lp["N"].copy(), lp["N_side"].copy(), max_iter=K, trace_cap=K)
    assert res.status == ref.status and res.iters == len(ref.trace) > 5
    k = len(ref.trace)
    assert (tr["entering"][:k] == ref.trace["entering"]).all() and (tr["leaving"][:k] == ref.trace["leaving"]).all()


def test_sharded_tableau_two_gpus_matches_single_gpu_and_oracle(env):
    """Column-sharded run (torchrun, 2 ranks) == single-GPU run == oracle (order-free tie rule), pivot for pivot."""
    import subprocess, sys, os, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (tools/sharded_check.py is also run by hand under gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29631", os.path.join(root, "tools", "sharded_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARDED_CHECK_OK" in out.stdout


def test_sharded_engines_one_rank_match_oracle(env):
    """Both sharded engines (NCCL path and the peer-memory engine of peer.cuh) with a single rank: the whole protocol --
    tickets, mailbox words, column words, parity double-buffering -- runs against the rank's own exchange buffer."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1", "--master-addr", "127.0.0.1",
                          "--master-port", "29633", os.path.join(root, "tools", "sharded_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARDED_CHECK_OK" in out.stdout


def test_generated_dual_lp_matches_numpy_twin_and_oracle(env):
    import bench_lp
    N, O, ctx = env["N"], env["O"], env["ctx"]
    m, ns, seed, K = 64, 128, 9, 60
    o = N.default_opts(K, engine=N.ENGINE_REVISED)
    tr = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr); o.trace_cap = K
    ctx.check(N.lib.ellp_b200_generate_dense_ex(ctx.h, m, ns, seed, 1, C.byref(o)))
    n = ns + m
    A = np.zeros((m, n), order="F"); c = np.zeros(n); b = np.zeros(m)
    ctx.check(N.lib.ellp_b200_download_std_form(ctx.h, N.ptr(A), N.ptr(c), N.ptr(b), None, None, None))
    lp = bench_lp.dense_lp(m, ns, seed, 1)
    assert A.tobytes() == lp["A"].tobytes() and c.tobytes() == lp["c"].tobytes() and b.tobytes() == lp["b"].tobytes()
    res = N.Result()
    ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
    ref = O.solve_with_initial(O.DUAL, m, n, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], lp["x"].copy(),
                               lp["B"].copy(), lp["N"].copy(), lp["N_side"].copy(), lp["y"].copy(), lp["d"].copy(),
                               max_iter=K, trace_cap=K)
    k = len(ref.trace)
    assert res.status == ref.status and res.iters == k > 5
    assert (tr["entering"][:k] == ref.trace["entering"]).all() and (tr["leaving"][:k] == ref.trace["leaving"]).all()


# ---------------------------------------------------------------- K6: shared-memory batched kernel
def _std_form_of(env, prob):
    """Option<StandardForm>::from(Problem) through the product's host layer (ellp_b200_stage_new, which = 0)."""
    N = env["N"]
    arr = prob.to_arrays()
    desc, _keep = N.problem_desc(arr)
    h = C.c_void_p(); infeasible = C.c_int(0); err = C.create_string_buffer(256)
    assert N.lib.ellp_b200_stage_new(C.byref(desc), 0, C.byref(h), C.byref(infeasible), err) == N.OK
    assert not infeasible.value
    dims = [C.c_int32() for _ in range(7)]
    N.lib.ellp_b200_stage_dims(h, *[C.byref(v) for v in dims])
    m, n = dims[0].value, dims[1].value
    A = np.zeros((m, n), order="F"); c = np.zeros(n); b = np.zeros(m); kind = np.zeros(n, dtype=np.uint8); lb = np.zeros(n); ub = np.zeros(n)
    N.lib.ellp_b200_stage_copy(h, N.ptr(A), N.ptr(c), N.ptr(b), N.ptr(kind), N.ptr(lb), N.ptr(ub), None, None, None, None, None, None)
    N.lib.ellp_b200_stage_free(h)
    return m, n, A, c, b, kind, lb, ub


@pytest.mark.parametrize("name", P.NETLIB + ["small_prob_2", "small_prob_3", "small_prob_5", "small_prob_6", "beale_cycle",
                                             "small_prob_unbounded_1", "two_variables_infeasible_with_bounds"])
def test_batch_kernel_single_lp_follows_oracle_pivot_for_pivot(env, name):
    """One launch solves the whole two-phase primal; same status, objective, point and pivot sequence as the oracle."""
    S, O = env["S"], env["O"]
    prob, exp = P.netlib(name) if name in P.NETLIB else P.GOLDEN_BY_NAME[name]()
    m, n, A, c, b, kind, lb, ub = _std_form_of(env, prob)
    r = S.primal_solve_batch(A.T[None].copy(), c[None], b[None], kind[None], lb[None], ub[None], 1000, trace_cap=4096, ctx=env["ctx"])
    ref = O.solve(prob, O.PRIMAL, 1000, O.MODE_EXACT, trace_cap=4096)
    assert r.err[0] == 0
    assert r.status[0] == ref.status
    assert list(r.iters[0]) == ref.iters[:2]
    k = len(ref.trace)
    assert r.trace_len[0] == k
    assert (r.trace[0]["entering"][:k] == ref.trace["entering"]).all() and (r.trace[0]["leaving"][:k] == ref.trace["leaving"]).all()
    if ref.status == O.OPTIMAL:
        assert _rel(r.obj[0], ref.obj) < 1e-9
        nv = len(prob.variables)
        np.testing.assert_allclose(r.x[0][:nv], ref.x, rtol=1e-9, atol=1e-9)
        P.check_expectation(exp, "Optimal", r.obj[0], r.x[0][:nv])


def test_batch_kernel_generated_batch_matches_oracle(env):
    """configs[3] generator at reduced count: every LP of the batch agrees with the oracle's solve() on the same LP
    (status, objective 1e-9, pivot counts per phase)."""
    from ellp_b200.problem import Bound, ConstraintOp, Problem
    N, O, ctx = env["N"], env["O"], env["ctx"]
    for (nlp, m, ns, seed) in [(40, 16, 24, 3), (300, 64, 128, 0)]:
        ctx.check(N.lib.ellp_b200_batch_generate(ctx.h, nlp, m, ns, seed, 0, 0))
        o = N.default_opts(None)
        res = N.BatchResult()
        ctx.check(N.lib.ellp_b200_batch_run(ctx.h, C.byref(o), C.byref(res)))
        n0 = ns + m
        status = np.zeros(nlp, dtype=np.int32); obj = np.zeros(nlp); x = np.zeros((nlp, n0 + m)); iters = np.zeros((nlp, 2), dtype=np.int32); err = np.zeros(nlp, dtype=np.int32)
        out = N.BatchResult(N.ptr(status), N.ptr(obj), N.ptr(x), N.ptr(iters), N.ptr(err), None, 0, None, 0.0, 0, 0)
        ctx.check(N.lib.ellp_b200_batch_download(ctx.h, C.byref(out)))
        assert (err == 0).all() and (status == N.OPTIMAL).all()
        assert out.pivots == iters.sum()
        for k in ([0, 1, nlp - 1] if m == 64 else range(0, nlp, 7)):
            A = np.zeros((m, n0), order="F"); c = np.zeros(n0); b = np.zeros(m)
            ctx.check(N.lib.ellp_b200_batch_download_lp(ctx.h, k, N.ptr(A), N.ptr(c), N.ptr(b)))
            p = Problem.new()
            ids = [p.add_var(c[j], Bound.Lower(0.0)) for j in range(ns)]
            for i in range(m):
                p.add_constraint([(ids[j], A[i, j]) for j in range(ns)], ConstraintOp.Lte, b[i])
            ref = O.solve(p, O.PRIMAL, None, O.MODE_EXACT)
            assert ref.status == O.OPTIMAL
            assert _rel(obj[k], ref.obj) < 1e-9
            np.testing.assert_allclose(x[k][:ns], ref.x, rtol=1e-9, atol=1e-9)
            assert list(iters[k]) == ref.iters[:2], (k, iters[k], ref.iters)
        print(f"batch {nlp} x ({m}x{ns}): {res.ms_device:.3f} ms, {out.pivots} pivots, {out.pivots / res.ms_device / 1e3:.2f} M pivots/s")


# ---------------------------------------------------------------- K4: blocked LU + DMMA refactorisation
@pytest.mark.parametrize("m", [1, 2, 27, 33, 74, 128, 300, 513, 1000])
def test_blocked_lu_dmma_inverse_matches_numpy_and_gauss_jordan(env, m):
    N, ctx = env["N"], env["ctx"]
    rng = np.random.default_rng(1000 + m)
    Bm = np.asfortranarray(rng.standard_normal((m, m)) + 0.1 * np.eye(m))
    inv_lu = np.zeros((m, m), order="F"); inv_gj = np.zeros((m, m), order="F")
    try:
        ctx.set_tuning("refactor_mode", 2)
        ctx.check(N.lib.ellp_b200_invert(ctx.h, N.ptr(Bm), m, N.ptr(inv_lu)))
        ctx.set_tuning("refactor_mode", 1)
        ctx.check(N.lib.ellp_b200_invert(ctx.h, N.ptr(Bm), m, N.ptr(inv_gj)))
    finally:
        ctx.set_tuning("refactor_mode", 0)
    tol = 1e-9 * np.linalg.cond(Bm)
    np.testing.assert_allclose(inv_lu @ Bm, np.eye(m), atol=tol)
    np.testing.assert_allclose(inv_lu, inv_gj, atol=tol, rtol=1e-7)


def test_blocked_lu_detects_singular_basis_and_times_large_inverse(env):
    import time
    from ellp_b200.problem import EllPError
    N, ctx = env["N"], env["ctx"]
    m = 200
    Bm = np.asfortranarray(np.random.default_rng(5).standard_normal((m, m)))
    Bm[:, 77] = Bm[:, 3]  # rank deficient => a pivot below EPS
    inv = np.zeros((m, m), order="F")
    try:
        ctx.set_tuning("refactor_mode", 2)
        rc = N.lib.ellp_b200_invert(ctx.h, N.ptr(Bm), m, N.ptr(inv))
        assert rc == N.E_ELLP and b"invalid B, A_B is not invertible" in N.lib.ellp_b200_last_error(ctx.h)
        m = 2048
        Bm = np.asfortranarray(np.random.default_rng(6).standard_normal((m, m)) + 3 * np.eye(m))
        inv = np.zeros((m, m), order="F")
        for mode, name in ((2, "blocked LU + DMMA"), (1, "Gauss-Jordan")):
            ctx.set_tuning("refactor_mode", mode)
            ctx.check(N.lib.ellp_b200_invert(ctx.h, N.ptr(Bm), m, N.ptr(inv)))
            t0 = time.perf_counter()
            ctx.check(N.lib.ellp_b200_invert(ctx.h, N.ptr(Bm), m, N.ptr(inv)))
            dt = time.perf_counter() - t0
            np.testing.assert_allclose(inv @ Bm, np.eye(m), atol=1e-8)
            print(f"invert m={m} {name}: {dt * 1e3:.1f} ms incl. 2 x 32 MB PCIe copies ({2.67 * m ** 3 / dt / 1e12:.2f} TFLOP/s equivalent)")
    finally:
        ctx.set_tuning("refactor_mode", 0)


@pytest.mark.parametrize("which", ["primal", "dual"])
def test_solver_parity_with_lu_refactorisation_every_few_pivots(env, which):
    # refactorising often with the blocked LU must not change status / objective (netlib BLEND, m = 74)
    prob, exp = P.netlib("blend")
    ctx, O = env["ctx"], env["O"]
    try:
        ctx.set_tuning("refactor_mode", 2)
        # ENGINE_REVISED: the one-launch shared-memory path (K6, engine AUTO) ignores refactor_every
        res = _solver(env, which, refactor_every=7, engine=env["N"].ENGINE_REVISED).solve(prob)
    finally:
        ctx.set_tuning("refactor_mode", 0)
    ref = O.solve(prob, O.PRIMAL if which == "primal" else O.DUAL, 1000, O.MODE_EXACT)
    assert res.is_optimal and _rel(res.solution.obj(), ref.obj) < 1e-9
    P.check_expectation(exp, res.kind, res.solution.obj(), res.solution.x())


@pytest.mark.parametrize("engine,block_k,mode", [("revised", 0, 2), ("revised", 0, 1), ("tableau", 0, 0), ("tableau", 8, 0)])
def test_refactorisation_really_runs_and_keeps_the_pivot_sequence(env, engine, block_k, mode):
    # refactor_every = 5 is not a multiple of block_k = 8: pending (U, V) slots must be flushed before the tableau is rebuilt
    O, S, N, ctx = env["O"], env["S"], env["N"], env["ctx"]
    m, n = 48, 80
    A, b, c = _dense_lp(77, m, n)
    Af, cf, kind, lb, ub, x0, B0, N0, Ns0 = _slack_start_primal(A, b, c)
    xo, Bo, No, Nso = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
    ref = O.solve_with_initial(O.PRIMAL, m, n + m, Af, cf, b, kind, lb, ub, xo, Bo, No, Nso, max_iter=None, trace_cap=20000)
    xg, Bg, Ng, Nsg = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
    try:
        ctx.set_tuning("refactor_mode", mode)
        sol = S.GpuPrimalSimplexSolver.new(None, ctx=ctx, trace_cap=20000, refactor_every=5, block_k=block_k,
                                           engine=N.ENGINE_REVISED if engine == "revised" else N.ENGINE_TABLEAU)
        res, trace = sol.solve_with_initial(m, n + m, Af, cf, b, kind, lb, ub, xg, Bg, Ng, Nsg)
    finally:
        ctx.set_tuning("refactor_mode", 0)
    assert res.refactors >= len(ref.trace) // 8 > 0, res.refactors
    assert res.status == ref.status == O.OPTIMAL and res.iters == len(ref.trace)
    assert (trace["entering"] == ref.trace["entering"]).all() and (trace["leaving"] == ref.trace["leaving"]).all()
    np.testing.assert_array_equal(Bg, Bo)
    np.testing.assert_allclose(xg, xo, rtol=1e-9, atol=1e-9)


# ---------------------------------------------------------------- optional rules: dual steepest edge + Harris ratio
@pytest.mark.parametrize("pricing,ratio", [(1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("name", P.NETLIB)
def test_dual_steepest_edge_and_harris_reach_the_same_optimum(env, name, pricing, ratio):
    # no reference counterpart (README.md:114 lists steepest edge as TODO): status + objective parity only
    prob, exp = P.netlib(name)
    O = env["O"]
    res = _solver(env, "dual", pricing=pricing, ratio=ratio).solve(prob)
    ref = O.solve(prob, O.DUAL, 1000, O.MODE_EXACT)
    assert res.is_optimal and ref.status == O.OPTIMAL
    assert _rel(res.solution.obj(), ref.obj) < 1e-9
    P.check_expectation(exp, res.kind, res.solution.obj(), res.solution.x())
    print(f"{name} pricing={pricing} ratio={ratio}: pivots {res.iters} (reference rules: {ref.iters})")


def test_dual_steepest_edge_needs_fewer_pivots_on_a_dense_lp(env):
    import bench_lp
    S, N, O = env["S"], env["N"], env["O"]
    m, ns = 96, 192
    lp = bench_lp.dense_lp(m, ns, 21, 1)
    out = {}
    for pricing in (0, 1):
        st = [lp[k].copy() for k in ("x", "B", "N", "N_side", "y", "d")]
        sol = S.GpuDualSimplexSolver.new(None, ctx=env["ctx"], pricing=pricing, ratio=pricing)
        res, _ = sol.solve_with_initial(m, m + ns, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], *st)
        assert res.status == N.OPTIMAL
        out[pricing] = (res.iters, float(np.dot(lp["c"], st[0])))
    assert _rel(out[1][1], out[0][1]) < 1e-9
    assert out[1][0] < out[0][0]
    print(f"dense {m}x{ns} dual: first-infeasible {out[0][0]} pivots, steepest edge + Harris {out[1][0]} pivots")


# ---------------------------------------------------------------- blocked tableau engine (deferred rank-k row reduction, K3b)
@pytest.mark.parametrize("R,C_,k", [(2, 2, 1), (8, 8, 3), (27, 51, 4), (64, 192, 32), (130, 70, 17), (513, 129, 64), (1024, 1000, 32),
                                    (2050, 333, 8)])
def test_rankk_update_matches_numpy_and_sequential_rank1(env, R, C_, k):
    """E -= U V on the fp64 tensor pipe.  DMMA accumulates the k products of an element in one chain starting from
    e_ij, like k sequential fma(-u, v, e): the result agrees with the exactly rounded sequential update to a few ulp
    of the magnitudes involved (the tensor pipe's internal rounding points are not specified bit-for-bit)."""
    N, ctx = env["N"], env["ctx"]
    rng = np.random.default_rng(R * 7919 + C_ * 31 + k)
    E = np.asfortranarray(rng.standard_normal((R, C_)))
    U = np.asfortranarray(rng.standard_normal((R, k)))
    V = np.ascontiguousarray(rng.standard_normal((k, C_)))
    got = E.copy(order="F")
    ctx.check(N.lib.ellp_b200_rankk_update(ctx.h, N.ptr(got), R, C_, R, N.ptr(U), N.ptr(V), k))
    want = E - U @ V
    scale = np.abs(E) + np.abs(U) @ np.abs(V)
    assert (np.abs(got - want) <= 4 * np.finfo(float).eps * scale * max(1, k)).all()


@pytest.mark.parametrize("R,C_,k", [(128, 128, 4), (130, 259, 17), (513, 1000, 40), (1024, 2048, 56), (640, 4100, 64), (2050, 333, 48)])
def test_rankk_update_kernel_variants_are_bit_identical(env, R, C_, k):
    """Every version of the rank-k row reduction (tuning key "flush_kernel": 1 = two CTAs per SM, 3 = register prefetch + bulk-copy
    ring, 4 = 16 consumer warps, 5 / 6 = 16 warps with the warp tile pipelined in 2 / 4 parts) accumulates the k products of an
    element in the same order on the same DMMA shape, so the stored tableaus are equal bit for bit -- switching the kernel can
    never change a pivot decision."""
    N, ctx = env["N"], env["ctx"]
    rng = np.random.default_rng(R * 31 + C_ * 7 + k)
    E = np.asfortranarray(rng.standard_normal((R, C_)))
    U = np.asfortranarray(rng.standard_normal((R, k)))
    V = np.ascontiguousarray(rng.standard_normal((k, C_)))
    out = {}
    try:
        for kern in (1, 3, 4, 5, 6, 7, 8, 9):
            ctx.set_tuning("flush_kernel", kern)
            got = E.copy(order="F")
            ctx.check(N.lib.ellp_b200_rankk_update(ctx.h, N.ptr(got), R, C_, R, N.ptr(U), N.ptr(V), k))
            out[kern] = got
    finally:
        ctx.set_tuning("flush_kernel", 0)
    scale = np.abs(E) + np.abs(U) @ np.abs(V)
    assert (np.abs(out[4] - (E - U @ V)) <= 4 * np.finfo(float).eps * scale * k).all()
    for kern in (1, 3, 5, 6, 7, 8, 9):
        np.testing.assert_array_equal(out[kern], out[4], err_msg=f"flush_kernel {kern} vs 4")


@pytest.mark.parametrize("seed,m,n,tie,bk", [(0, 24, 40, 0, 2), (1, 64, 128, 0, 7), (2, 64, 128, 1, 32), (3, 132, 190, 0, 64), (5, 260, 515, 0, 32)])
def test_blocked_tableau_engine_matches_oracle_trace(env, seed, m, n, tie, bk):
    """ellp_opts::block_k > 1: identical pivot sequence, basis and point as the oracle (and hence as the rank-1 engine)."""
    O, S, N = env["O"], env["S"], env["N"]
    A, b, c = _dense_lp(seed, m, n)
    Af, cf, kind, lb, ub, x0, B0, N0, Ns0 = _slack_start_primal(A, b, c)
    xo, Bo, No, Nso = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
    ref = O.solve_with_initial(O.PRIMAL, m, n + m, Af, cf, b, kind, lb, ub, xo, Bo, No, Nso, max_iter=None, mode=tie, trace_cap=20000)
    xg, Bg, Ng, Nsg = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
    sol = S.GpuPrimalSimplexSolver.new(None, ctx=env["ctx"], trace_cap=20000, tie_rule=tie, engine=N.ENGINE_TABLEAU, block_k=bk)
    res, trace = sol.solve_with_initial(m, n + m, Af, cf, b, kind, lb, ub, xg, Bg, Ng, Nsg)
    assert res.status == ref.status == O.OPTIMAL
    assert res.iters == len(ref.trace)
    assert (trace["entering"] == ref.trace["entering"]).all() and (trace["leaving"] == ref.trace["leaving"]).all()
    np.testing.assert_array_equal(Bg, Bo)
    np.testing.assert_allclose(xg, xo, rtol=1e-9, atol=1e-9)
    assert _rel(res.obj, ref.obj) < 1e-9


@pytest.mark.parametrize("make", P.GOLDEN + [lambda n=n: P.netlib(n) for n in P.NETLIB],
                         ids=[f.__name__ for f in P.GOLDEN] + P.NETLIB)
def test_blocked_tableau_engine_two_phase_primal(env, make):
    """The reference's own expectations (degenerate netlib LPs, bound flips, unbounded / infeasible verdicts) with the
    deferred row reduction: same verdict, objective and pivot counts as the oracle."""
    prob, exp = make()
    N, S = env["N"], env["S"]
    res = S.GpuPrimalSimplexSolver.default(ctx=env["ctx"], engine=N.ENGINE_TABLEAU, block_k=8).solve(prob)
    obj = res.solution.obj() if res.is_optimal else float("nan")
    x = res.solution.x() if res.is_optimal else []
    P.check_expectation(exp, res.kind, obj, x)
    ref = env["O"].solve(prob, env["O"].PRIMAL, 1000, env["O"].MODE_EXACT)
    assert res.kind == ref.status_name
    if res.is_optimal:
        assert _rel(obj, ref.obj) < 1e-9
        assert res.iters == ref.iters


@pytest.mark.parametrize("name", P.NETLIB + ["beale_cycle", "small_prob_2", "small_prob_5"])
@pytest.mark.parametrize("bk", [0, 8, 64])
def test_tableau_engines_with_the_order_free_tie_rule_follow_the_oracle_on_degenerate_lps(env, name, bk):
    """ELLP_TIES_CANONICAL (the rule the multi-GPU engines need) on heavily degenerate LPs: near-ties in pricing and in the ratio
    test take the second-round / exact-fold paths of the pivot kernels.  Same pivots as the oracle's canonical mode, phase by
    phase (Beale's LP cycles under this rule -- both sides must then stop at max_iter with the same trace)."""
    prob, exp = P.netlib(name) if name in P.NETLIB else P.GOLDEN_BY_NAME[name]()
    N, S, O = env["N"], env["S"], env["O"]
    res = S.GpuPrimalSimplexSolver.new(300, ctx=env["ctx"], engine=N.ENGINE_TABLEAU, block_k=bk, tie_rule=N.TIES_CANONICAL, trace_cap=4096).solve(prob)
    ref = O.solve(prob, O.PRIMAL, 300, O.MODE_CANONICAL, trace_cap=4096)
    assert res.kind == ref.status_name
    assert res.iters[:2] == ref.iters[:2]
    k = len(ref.trace)
    assert len(res.trace) == k
    assert (res.trace["entering"] == ref.trace["entering"]).all() and (res.trace["leaving"] == ref.trace["leaving"]).all()
    if res.is_optimal:
        assert _rel(res.solution.obj(), ref.obj) < 1e-9


def test_blocked_engine_continues_across_runs_and_matches_rank1_engine(env):
    """bench.py's usage: the LP stays resident and ellp_b200_run is called repeatedly with a pivot budget.  The blocked
    engine (flush at the end of every run) must follow the rank-1 engine pivot for pivot over several runs."""
    N, ctx = env["N"], env["ctx"]
    m, ns, seed, K, runs = 512, 1024, 3, 50, 4
    traces = {}
    for bk in (0, 16):
        o = N.default_opts(K, engine=N.ENGINE_TABLEAU, block_k=bk, check_every=16)
        tr = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr); o.trace_cap = K
        ctx.check(N.lib.ellp_b200_generate_dense(ctx.h, m, ns, seed, C.byref(o)))
        rows = []
        for _ in range(runs):
            res = N.Result()
            ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
            assert res.status == N.MAXITER and res.iters == K
            rows.append((tr["entering"].copy(), tr["leaving"].copy(), tr["step"].copy()))
        traces[bk] = rows
    for (e0, l0, s0), (e1, l1, s1) in zip(traces[0], traces[16]):
        assert (e0 == e1).all() and (l0 == l1).all()
        np.testing.assert_allclose(s1, s0, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("bk", [0, 16])
def test_condensed_fast_upload_matches_resident_run_and_falls_back_on_a_non_identity_basis(env, bk):
    """Large LP through the C ABI on HOST buffers: with a slack basis only the nonbasic columns cross PCIe (the basis columns
    are verified to be unit vectors by host threads during the DMA) -- same pivots, point and basis as the run on the LP
    generated in HBM; with a basis that is NOT the identity the call falls back to the full upload + device refactorisation
    and still follows the oracle."""
    import bench_lp
    N, S, O, ctx = env["N"], env["S"], env["O"], env["ctx"]
    m, ns, seed, K = 1024, 3072, 7, 48  # m * n * 8 = 33.5 MB >= the 32 MB threshold of the fast path
    n = m + ns
    o = N.default_opts(K, engine=N.ENGINE_TABLEAU, block_k=bk, check_every=16)
    tr1 = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr1); o.trace_cap = K
    ctx.check(N.lib.ellp_b200_generate_dense(ctx.h, m, ns, seed, C.byref(o)))
    r1 = N.Result()
    ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(r1)))
    x1 = np.zeros(n); B1 = np.zeros(m, dtype=np.int32); N1 = np.zeros(ns, dtype=np.int32); Ns1 = np.zeros(ns, dtype=np.uint8)
    ctx.check(N.lib.ellp_b200_download(ctx.h, C.byref(N.Point(N.ptr(x1), N.ptr(B1), N.ptr(N1), N.ptr(Ns1), None, None, m, ns))))
    lp = bench_lp.dense_lp(m, ns, seed)
    sol = S.GpuPrimalSimplexSolver.new(K, ctx=ctx, trace_cap=K, engine=N.ENGINE_TABLEAU, block_k=bk)
    x2, B2, N2, Ns2 = lp["x"].copy(), lp["B"].copy(), lp["N"].copy(), lp["N_side"].copy()
    res, tr2 = sol.solve_with_initial(m, n, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], x2, B2, N2, Ns2)
    assert res.status == r1.status == N.MAXITER and res.iters == r1.iters == K
    assert (tr2["entering"] == tr1["entering"]).all() and (tr2["leaving"] == tr1["leaving"]).all()
    assert x2.tobytes() == x1.tobytes() and np.array_equal(B2, B1) and np.array_equal(N2, N1) and np.array_equal(Ns2, Ns1)
    # the fast path must not have uploaded A
    rc = N.lib.ellp_b200_download_std_form(ctx.h, N.ptr(np.zeros((m, n), order="F")), None, None, None, None, None)
    assert rc == N.E_ARG
    # non-identity basis: scale one slack column and its variable (x_s -> x_s / 2 keeps A x = b)
    A3 = lp["A"].copy(order="F"); x3 = lp["x"].copy()
    j = int(lp["B"][5]); A3[:, j] *= 2.0; x3[j] *= 0.5
    K3 = 4
    xo, Bo, No, Nso = x3.copy(), lp["B"].copy(), lp["N"].copy(), lp["N_side"].copy()
    ref = O.solve_with_initial(O.PRIMAL, m, n, A3, lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], xo, Bo, No, Nso, max_iter=K3, trace_cap=K3)
    sol3 = S.GpuPrimalSimplexSolver.new(K3, ctx=ctx, trace_cap=K3, engine=N.ENGINE_TABLEAU, block_k=bk)
    xg, Bg, Ng, Nsg = x3.copy(), lp["B"].copy(), lp["N"].copy(), lp["N_side"].copy()
    res3, tr3 = sol3.solve_with_initial(m, n, A3, lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], xg, Bg, Ng, Nsg)
    k = len(ref.trace)
    assert res3.status == ref.status and res3.iters == k == K3
    assert (tr3["entering"] == ref.trace["entering"]).all() and (tr3["leaving"] == ref.trace["leaving"]).all()
    np.testing.assert_array_equal(Bg, Bo)
    np.testing.assert_allclose(xg, xo, rtol=1e-9, atol=1e-9)


def test_full_size_blocked_engine_matches_rank1_engine_and_stays_feasible(env):
    """north_star's target size (16384 x 32768), where the oracle cannot follow (one 16384^3 LU per pivot): size-independent
    properties instead.  (1) The fused blocked engine (k = 64, one cooperative launch per 64 pivots + one rank-64 flush on the
    fp64 tensor pipe) and the rank-1 engine (one k_rank1 sweep per pivot) must produce the SAME pivots and basis, and the same
    point / steps / objective to 1e-9; (2) the point stays primal feasible (A x = b to rounding, x >= 0), the objective never increases, and every
    entering variable had a negative reduced cost (Dantzig) -- checked on the host against the LP downloaded from HBM."""
    N, ctx = env["N"], env["ctx"]
    m, ns, seed, K = 16384, 16384, 0, 192
    n = m + ns
    runs = {}
    for bk in (0, 64):
        o = N.default_opts(K, engine=N.ENGINE_TABLEAU, block_k=bk, check_every=64)
        tr = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr); o.trace_cap = K
        ctx.check(N.lib.ellp_b200_generate_dense(ctx.h, m, ns, seed, C.byref(o)))
        if bk == 0:
            A = np.zeros((m, n), order="F"); c = np.zeros(n); b = np.zeros(m)
            ctx.check(N.lib.ellp_b200_download_std_form(ctx.h, N.ptr(A), N.ptr(c), N.ptr(b), None, None, None))
        res = N.Result()
        ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
        x = np.zeros(n); B = np.zeros(m, dtype=np.int32); Nv = np.zeros(ns, dtype=np.int32); Ns = np.zeros(ns, dtype=np.uint8)
        ctx.check(N.lib.ellp_b200_download(ctx.h, C.byref(N.Point(N.ptr(x), N.ptr(B), N.ptr(Nv), N.ptr(Ns), None, None, m, ns))))
        assert res.status == N.MAXITER and res.iters == K
        runs[bk] = (tr.copy(), x, B, Nv, Ns, res.obj)
    t0, x0, B0, N0, Ns0, obj0 = runs[0]
    t1, x1, B1, N1, Ns1, obj1 = runs[64]
    assert (t0["entering"] == t1["entering"]).all() and (t0["leaving"] == t1["leaving"]).all()
    # values agree to rounding, not bit for bit: the deferred form reproduces a pivot-row entry as E_r - (alpha_r - 1) p instead
    # of storing p itself (one ulp), everything else is the same sequence of fmas
    np.testing.assert_allclose(t1["step"], t0["step"], rtol=1e-9, atol=1e-9)   # steps are differences of O(1e3) values
    np.testing.assert_allclose(t1["obj"], t0["obj"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(x1, x0, rtol=1e-9, atol=1e-9)
    assert np.array_equal(B0, B1) and np.array_equal(N0, N1) and np.array_equal(Ns0, Ns1) and abs(obj0 - obj1) <= 1e-10 * abs(obj0)
    # size-independent properties of the path
    assert (x1 >= -1e-9).all()
    resid = A @ x1 - b
    assert np.abs(resid).max() <= 1e-9 * np.abs(b).max()
    assert sorted(B1.tolist() + N1.tolist()) == list(range(n))             # B and N partition the variables
    assert np.abs(x1[N1]).max() <= 1e-9                                    # nonbasic variables sit at their (zero) bound, to rounding:
    #                                                                        like the reference (x[B] += lambda d, primal :408-417) a leaving
    #                                                                        variable is not snapped to its bound
    assert (np.diff(t1["obj"]) <= 1e-9 * np.abs(t1["obj"]).max()).all()    # the objective never increases
    assert abs(float(c @ x1) - obj1) <= 1e-9 * max(1.0, abs(obj1))
    assert (t1["step"] >= 0).all()


@pytest.mark.parametrize("seed", range(60))
def test_random_lps_with_mixed_bounds_match_oracle(env, seed):
    """Seeded random LPs with Lower / Upper / Free / Fixed variables and <=, >=, = rows (the generator of
    tests/test_oracle_highs.py, where the oracle is cross-checked against HiGHS): both GPU solvers must return the oracle's
    verdict, objective and point -- including the reference's quirky verdicts (square systems, Fixed variables in the dual)."""
    import test_oracle_highs as TH
    O, N = env["O"], env["N"]
    prob = TH._random_lp(seed)[0]
    for which, tag in ((O.PRIMAL, "primal"), (O.DUAL, "dual")):
        try:
            ref = O.solve(prob, which, 1000, O.MODE_EXACT)
        except O.OracleError:
            with pytest.raises(Exception):
                _solver(env, tag).solve(prob)
            continue
        res = _solver(env, tag).solve(prob)
        assert res.kind == ref.status_name, (seed, tag, res.kind, ref.status_name)
        assert res.used_primal_fallback == ref.used_primal_fallback
        if res.is_optimal:
            assert _rel(res.solution.obj(), ref.obj) < 1e-9, (seed, tag)
            np.testing.assert_allclose(res.solution.x(), ref.x, rtol=1e-8, atol=1e-8)


def test_fused_pivot_kernel_with_more_rows_than_threads(env):
    """m = 40960 rows > 148 CTAs x 256 threads: every phase of k_blk_pivots_fused walks its grid-stride loops more than once
    (the register row cache only covers the first row of a thread).  Same pivots / basis as the rank-1 engine, values to 1e-9."""
    N, ctx = env["N"], env["ctx"]
    m, ns, seed, K = 40960, 4096, 3, 48
    n = m + ns
    runs = {}
    for bk in (0, 16):
        o = N.default_opts(K, engine=N.ENGINE_TABLEAU, block_k=bk, check_every=16)
        tr = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr); o.trace_cap = K
        ctx.check(N.lib.ellp_b200_generate_dense(ctx.h, m, ns, seed, C.byref(o)))
        res = N.Result()
        ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
        x = np.zeros(n); B = np.zeros(m, dtype=np.int32); Nv = np.zeros(ns, dtype=np.int32); Ns = np.zeros(ns, dtype=np.uint8)
        ctx.check(N.lib.ellp_b200_download(ctx.h, C.byref(N.Point(N.ptr(x), N.ptr(B), N.ptr(Nv), N.ptr(Ns), None, None, m, ns))))
        assert res.status == N.MAXITER and res.iters == K
        runs[bk] = (tr.copy(), x, B, Nv, Ns, res.obj)
    t0, x0, B0, N0, Ns0, obj0 = runs[0]
    t1, x1, B1, N1, Ns1, obj1 = runs[16]
    assert (t0["entering"] == t1["entering"]).all() and (t0["leaving"] == t1["leaving"]).all()
    np.testing.assert_allclose(t1["step"], t0["step"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(x1, x0, rtol=1e-9, atol=1e-9)
    assert np.array_equal(B0, B1) and np.array_equal(N0, N1) and np.array_equal(Ns0, Ns1) and abs(obj0 - obj1) <= 1e-9 * abs(obj0)



# ---------------------------------------------------------------- dual simplex on the blocked condensed tableau (dual_blocked.cuh)
def _gte_dual_start(A, b, c):
    """min c.x, A x - s = b, x,s >= 0 with the slack basis B = -I (dual feasible: y = 0, d = c >= 0)."""
    m, n = A.shape
    Af = np.asfortranarray(np.hstack([A, -np.eye(m)]))
    cf = np.concatenate([c, np.zeros(m)])
    kind = np.ones(n + m, dtype=np.uint8); lb = np.zeros(n + m); ub = np.zeros(n + m)
    x0 = np.concatenate([np.zeros(n), -b]); B0 = np.arange(n, n + m, dtype=np.int32)
    N0 = np.arange(n, dtype=np.int32); Ns0 = np.zeros(n, dtype=np.uint8)
    return Af, cf, kind, lb, ub, [x0, B0, N0, Ns0, np.zeros(m), cf.copy()]


@pytest.mark.parametrize("seed,m,n,bk", [(10, 24, 40, 2), (11, 64, 128, 7), (12, 130, 190, 32), (13, 64, 128, 64), (14, 260, 515, 32), (15, 132, 190, 0)])
def test_dual_tableau_engine_matches_oracle_trace(env, seed, m, n, bk):
    """dual_simplex_solver.rs:188-334 on the blocked tableau (general starting basis -I => T = B^-1 A_N through K4): same pivots,
    x, y, d, B, N as the oracle."""
    O, S, N = env["O"], env["S"], env["N"]
    A, b, c = _dense_lp(seed, m, n)
    Af, cf, kind, lb, ub, start = _gte_dual_start(A, b, c)
    st = [a.copy() for a in start]
    ref = O.solve_with_initial(O.DUAL, m, n + m, Af, cf, b, kind, lb, ub, *st, max_iter=None, trace_cap=20000)
    sg = [a.copy() for a in start]
    sol = S.GpuDualSimplexSolver.new(None, ctx=env["ctx"], trace_cap=20000, engine=N.ENGINE_TABLEAU, block_k=bk, check_every=5)
    res, trace = sol.solve_with_initial(m, n + m, Af, cf, b, kind, lb, ub, *sg)
    assert res.status == ref.status == O.OPTIMAL
    assert res.iters == len(ref.trace) > 5
    assert (trace["entering"] == ref.trace["entering"]).all() and (trace["leaving"] == ref.trace["leaving"]).all()
    np.testing.assert_allclose(trace["step"], ref.trace["step"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(trace["obj"], ref.trace["obj"], rtol=1e-8, atol=1e-8)
    for g, o in zip(sg[:1] + sg[4:], st[:1] + st[4:]):
        np.testing.assert_allclose(g, o, rtol=1e-9, atol=1e-9)
    for g, o in zip(sg[1:4], st[1:4]):
        np.testing.assert_array_equal(g, o)
    assert _rel(res.obj, ref.obj) < 1e-9
    # d = c - A^T y holds for the exported dual point
    np.testing.assert_allclose(sg[5], cf - Af.T @ sg[4], rtol=0, atol=1e-8)


@pytest.mark.parametrize("bk", [8, 32])
@pytest.mark.parametrize("make", P.GOLDEN + [lambda n=n: P.netlib(n) for n in P.NETLIB],
                         ids=[f.__name__ for f in P.GOLDEN] + P.NETLIB)
def test_dual_tableau_engine_two_phase(env, make, bk):
    """DualSimplexSolver::solve (dual_simplex_solver.rs:32-108) with both phases on the tableau engine: the starting bases come
    from an LU of A^T (dual_problem.rs:140-160), i.e. they are NOT the identity."""
    prob, exp = make()
    O, N = env["O"], env["N"]
    res = _solver(env, "dual", engine=N.ENGINE_TABLEAU, block_k=bk, trace_cap=8192).solve(prob)
    obj = res.solution.obj() if res.is_optimal else float("nan")
    x = res.solution.x() if res.is_optimal else []
    P.check_expectation(exp, res.kind, obj, x)
    ref = O.solve(prob, O.DUAL, 1000, O.MODE_EXACT, trace_cap=8192)
    assert res.kind == ref.status_name
    assert res.used_primal_fallback == ref.used_primal_fallback
    if res.is_optimal:
        assert _rel(obj, ref.obj) < 1e-9


def test_afiro_dual_pivot_sequence_on_the_tableau_engine(env):
    prob, _ = P.netlib("afiro")
    O, N = env["O"], env["N"]
    res = _solver(env, "dual", trace_cap=4096, engine=N.ENGINE_TABLEAU, block_k=16).solve(prob)
    ref = O.solve(prob, O.DUAL, 1000, O.MODE_EXACT, trace_cap=4096)
    assert res.iters == ref.iters
    assert (res.trace["entering"] == ref.trace["entering"]).all()
    assert (res.trace["leaving"] == ref.trace["leaving"]).all()
    np.testing.assert_allclose(res.trace["step"], ref.trace["step"], rtol=1e-9, atol=1e-9)


def test_generated_dual_lp_on_the_tableau_engine_matches_oracle_and_revised_engine(env):
    import bench_lp
    N, O, ctx = env["N"], env["O"], env["ctx"]
    m, ns, seed, K = 64, 128, 9, 60
    n = ns + m
    lp = bench_lp.dense_lp(m, ns, seed, 1)
    ref = O.solve_with_initial(O.DUAL, m, n, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], lp["x"].copy(),
                               lp["B"].copy(), lp["N"].copy(), lp["N_side"].copy(), lp["y"].copy(), lp["d"].copy(),
                               max_iter=K, trace_cap=K)
    k = len(ref.trace)
    for bk in (8, 32):
        o = N.default_opts(K, engine=N.ENGINE_TABLEAU, block_k=bk, check_every=16)
        tr = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr); o.trace_cap = K
        ctx.check(N.lib.ellp_b200_generate_dense_ex(ctx.h, m, ns, seed, 1, C.byref(o)))
        res = N.Result()
        ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
        assert res.status == ref.status and res.iters == k > 5
        assert (tr["entering"][:k] == ref.trace["entering"]).all() and (tr["leaving"][:k] == ref.trace["leaving"]).all()
        x = np.zeros(n); B = np.zeros(m, dtype=np.int32); Nv = np.zeros(ns, dtype=np.int32); Ns = np.zeros(ns, dtype=np.uint8)
        y = np.zeros(m); d = np.zeros(n)
        pt = N.Point(N.ptr(x), N.ptr(B), N.ptr(Nv), N.ptr(Ns), N.ptr(y), N.ptr(d), m, ns)
        ctx.check(N.lib.ellp_b200_download(ctx.h, C.byref(pt)))
        np.testing.assert_allclose(d, lp["c"] - lp["A"].T @ y, rtol=0, atol=1e-8)
        np.testing.assert_allclose(lp["A"] @ x, lp["b"], rtol=1e-9, atol=1e-8)


def test_dual_tableau_error_paths(env):
    """Free nonbasic with a zero pivot-row entry and a zero reduced cost: 0 / 0 = NaN => partial_cmp().unwrap() panics
    (dual :279); no eligible entering column => Infeasible (dual :281-284)."""
    O, S, N = env["O"], env["S"], env["N"]
    from ellp_b200.solver import EllPPanic
    # x0 + s = -1, x0 >= 0, s >= 0 basic: row has only nonnegative entries for a Lower-side nonbasic => dual unbounded
    A = np.asfortranarray(np.array([[1.0, 1.0]])); c = np.array([1.0, 0.0]); b = np.array([-1.0])
    kind = np.array([N.LOWER, N.LOWER], dtype=np.uint8); lb = np.zeros(2); ub = np.zeros(2)
    st = [np.array([0.0, -1.0]), np.array([1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([0], dtype=np.uint8), np.zeros(1), c.copy()]
    for engine in (N.ENGINE_TABLEAU, N.ENGINE_REVISED):
        sg = [a.copy() for a in st]
        res, _ = S.GpuDualSimplexSolver.new(None, ctx=env["ctx"], engine=engine, block_k=4).solve_with_initial(1, 2, A, c, b, kind, lb, ub, *sg)
        so = [a.copy() for a in st]
        ref = O.solve_with_initial(O.DUAL, 1, 2, A, c, b, kind, lb, ub, *so, max_iter=None)
        assert res.status == ref.status == O.INFEASIBLE
    # Free nonbasic column that is identically zero with d = 0: ratio 0 / 0
    A = np.asfortranarray(np.array([[0.0, 1.0]])); c = np.array([0.0, 0.0])
    kind = np.array([N.FREE, N.LOWER], dtype=np.uint8)
    st = [np.array([0.0, -1.0]), np.array([1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([N.NB_FREE], dtype=np.uint8), np.zeros(1), c.copy()]
    for engine in (N.ENGINE_TABLEAU, N.ENGINE_REVISED):
        sg = [a.copy() for a in st]
        with pytest.raises(EllPPanic, match="unwrap"):
            S.GpuDualSimplexSolver.new(None, ctx=env["ctx"], engine=engine, block_k=4).solve_with_initial(1, 2, A, c, b, kind, lb, ub, *sg)
    # dual-infeasible starting point (dual :139-151)
    A = np.asfortranarray(np.array([[1.0, 1.0]])); c = np.array([-1.0, 0.0])
    kind = np.array([N.LOWER, N.LOWER], dtype=np.uint8)
    st = [np.array([0.0, 1.0]), np.array([1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([0], dtype=np.uint8), np.zeros(1), c.copy()]
    with pytest.raises(EllPPanic, match="dual infeasible"):
        S.GpuDualSimplexSolver.new(None, ctx=env["ctx"], engine=N.ENGINE_TABLEAU, block_k=4).solve_with_initial(1, 2, A, c, b, kind, lb, ub, *st)


def test_plain_c_program_drives_the_boundary(env):
    """tests/c_abi/c_abi_smoke.c (gcc -std=c99): primal + dual solve_with_initial and the reference's Err text from C."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "c_abi", "c_abi_smoke")
    if not os.path.exists(exe):
        import test_abi_cpu
        exe = test_abi_cpu._build_c_abi_smoke()
    out = subprocess.run([exe, "solve"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "C_ABI_SMOKE_OK" in out.stdout, out.stdout + out.stderr


# ---------------------------------------------------------------- TwoSided fuzz, device panic paths, K6 sample size
@pytest.mark.parametrize("seed", range(60))
def test_random_lps_with_twosided_bounds_match_oracle(env, seed):
    """GPU vs ORACLE (not vs HiGHS) on LPs with TwoSided / Fixed / Free / Lower / Upper variables: quirks Q3 (`x_i < lb` on the
    upper branch, primal :359-363) and Q17 live on TwoSided variables, and whatever the reference does there -- a wrong optimum,
    `assert!(lambda >= 0.)`, a dimension-mismatch panic -- the GPU path must do the same."""
    O, N = env["O"], env["N"]
    prob = P.random_lp_all_bounds(seed)
    for which, tag in ((O.PRIMAL, "primal"), (O.DUAL, "dual")):
        for engine in (N.ENGINE_AUTO, N.ENGINE_REVISED):
            try:
                ref = O.solve(prob, which, 1000, O.MODE_EXACT)
            except O.OracleError as e:
                with pytest.raises(Exception) as ei:
                    _solver(env, tag, engine=engine).solve(prob)
                # same panic / Err site: compare the leading words of the message
                assert str(e).split(": ", 1)[1][:24] in str(ei.value), (seed, tag, str(e), str(ei.value))
                continue
            res = _solver(env, tag, engine=engine).solve(prob)
            assert res.kind == ref.status_name, (seed, tag, engine, res.kind, ref.status_name)
            assert res.used_primal_fallback == ref.used_primal_fallback
            assert res.iters == ref.iters, (seed, tag, engine, res.iters, ref.iters)
            if res.is_optimal:
                assert _rel(res.solution.obj(), ref.obj) < 1e-9, (seed, tag)
                np.testing.assert_allclose(res.solution.x(), ref.x, rtol=1e-8, atol=1e-8)


def test_device_panic_paths_match_the_reference(env):
    """panic!/assert! sites of the pivot loop reached ON THE DEVICE (PivotState::err), every engine that implements the path:
    `pivot should have been unbounded` (primal :229), `assertion failed: lambda >= 0.` (primal :402).  `NaN detected` (primal
    :282) is dead code in the reference -- (r1 - r2).abs() >= EPS is false for a NaN, so the index rule decides -- and a NaN cost
    simply cycles to MaxIter, on the GPU as in the oracle."""
    O, S, N = env["O"], env["S"], env["N"]
    from ellp_b200.solver import EllPPanic
    engines = [(N.ENGINE_AUTO, 0), (N.ENGINE_REVISED, 0), (N.ENGINE_TABLEAU, 0), (N.ENGINE_TABLEAU, 4)]
    # (1) a nonbasic TwoSided variable parked on the Free side whose own range is the binding ratio: bound flip of a Free side
    A = np.asfortranarray(np.array([[1.0, 1.0]])); c = np.array([-1.0, 0.0]); b = np.array([2.0])
    kind = np.array([N.TWOSIDED, N.FREE], dtype=np.uint8); lb = np.array([0., 0.]); ub = np.array([1., 0.])
    st = [np.array([0.0, 2.0]), np.array([1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([N.NB_FREE], dtype=np.uint8)]
    with pytest.raises(O.OracleError, match="pivot should have been unbounded"):
        O.solve_with_initial(O.PRIMAL, 1, 2, A, c, b, kind, lb, ub, *[a.copy() for a in st], max_iter=50)
    for engine, bk in engines:
        with pytest.raises(EllPPanic, match="pivot should have been unbounded"):
            S.GpuPrimalSimplexSolver.new(50, ctx=env["ctx"], engine=engine, block_k=bk).solve_with_initial(1, 2, A, c, b, kind, lb, ub, *[a.copy() for a in st])
    # (2) quirk Q3: basic TwoSided variable below its lower bound moving further down => negative ratio => assert!(lambda >= 0.)
    A = np.asfortranarray(np.array([[1.0, 1.0]])); c = np.array([-1.0, 0.0]); b = np.array([-1.0])
    kind = np.array([N.LOWER, N.TWOSIDED], dtype=np.uint8); lb = np.array([0., 0.]); ub = np.array([0., 5.])
    st = [np.array([0.0, -1.0]), np.array([1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([N.NB_LOWER], dtype=np.uint8)]
    with pytest.raises(O.OracleError, match="lambda >= 0"):
        O.solve_with_initial(O.PRIMAL, 1, 2, A, c, b, kind, lb, ub, *[a.copy() for a in st], max_iter=50)
    for engine, bk in engines:
        with pytest.raises(EllPPanic, match="lambda >= 0"):
            S.GpuPrimalSimplexSolver.new(50, ctx=env["ctx"], engine=engine, block_k=bk).solve_with_initial(1, 2, A, c, b, kind, lb, ub, *[a.copy() for a in st])
    # (3) NaN costs are outside the contract: in the reference `NaN detected` is unreachable and a NaN key is compared by variable
    # index (the oracle cycles to MaxIter on such an LP); the device kernels never select a NaN key.  Not asserted.


def test_batch_kernel_256_generated_lps_follow_the_oracle_pivot_for_pivot(env):
    """256 LPs of the configs[3] shape (64 x 128) against the oracle -- status, objective, point, pivot counts per
    phase AND the complete pivot trace (entering / leaving variable of every pivot of both phases)."""
    from ellp_b200.problem import Bound, ConstraintOp, Problem
    N, O, ctx = env["N"], env["O"], env["ctx"]
    nlp, m, ns, seed, cap = 256, 64, 128, 5, 1024
    ctx.check(N.lib.ellp_b200_batch_generate(ctx.h, nlp, m, ns, seed, 0, cap))
    o = N.default_opts(None)
    res = N.BatchResult()
    ctx.check(N.lib.ellp_b200_batch_run(ctx.h, C.byref(o), C.byref(res)))
    n0 = ns + m
    status = np.zeros(nlp, dtype=np.int32); obj = np.zeros(nlp); x = np.zeros((nlp, n0 + m)); iters = np.zeros((nlp, 2), dtype=np.int32)
    err = np.zeros(nlp, dtype=np.int32); tr = np.zeros((nlp, cap), dtype=N.TRACE_DTYPE); tl = np.zeros(nlp, dtype=np.int32)
    out = N.BatchResult(N.ptr(status), N.ptr(obj), N.ptr(x), N.ptr(iters), N.ptr(err), N.ptr(tr), cap, N.ptr(tl), 0.0, 0, 0)
    ctx.check(N.lib.ellp_b200_batch_download(ctx.h, C.byref(out)))
    assert (err == 0).all() and (status == N.OPTIMAL).all()
    A_all = np.zeros((nlp, n0, m)); c_all = np.zeros((nlp, n0)); b_all = np.zeros((nlp, m))
    ctx.check(N.lib.ellp_b200_batch_download_all(ctx.h, N.ptr(A_all), N.ptr(c_all), N.ptr(b_all)))
    pivots = 0
    for k in range(nlp):
        A = A_all[k].T  # per-LP column-major m x n0
        p = Problem.new()
        ids = [p.add_var(c_all[k][j], Bound.Lower(0.0)) for j in range(ns)]
        for i in range(m):
            p.add_constraint([(ids[j], A[i, j]) for j in range(ns)], ConstraintOp.Lte, b_all[k][i])
        ref = O.solve(p, O.PRIMAL, None, O.MODE_EXACT, trace_cap=cap)
        assert ref.status == O.OPTIMAL
        assert _rel(obj[k], ref.obj) < 1e-9
        np.testing.assert_allclose(x[k][:ns], ref.x, rtol=1e-9, atol=1e-9)
        assert list(iters[k]) == ref.iters[:2], (k, iters[k], ref.iters)
        L = int(tl[k])
        assert L == len(ref.trace) == int(iters[k].sum())
        # Option<StandardForm>::from(Problem) reorders the rows (QR with column pivoting on A^T, quirk Q13), and the artificial
        # variable of phase 1 is numbered by ROW (n + i, primal_problem.rs:239-246); K6 takes the standard form as given.  Map the
        # oracle's artificial indices back through that row permutation (b has no duplicates) before comparing.
        sf = O.stage(p, 0)
        perm = np.array([int(np.nonzero(b_all[k] == v)[0][0]) for v in np.array(sf.b)])
        assert sorted(perm.tolist()) == list(range(m))

        def relabel(v):
            v = np.array(v, dtype=np.int64)
            art = v >= n0
            v[art] = n0 + perm[v[art] - n0]
            return v

        ref_ent, ref_lv = relabel(ref.trace["entering"]), relabel(ref.trace["leaving"])
        bad = np.nonzero((tr[k]["entering"][:L] != ref_ent) | (tr[k]["leaving"][:L] != ref_lv))[0]
        assert bad.size == 0, "LP %d: %s" % (k, [(int(i), (int(tr[k]["entering"][i]), int(tr[k]["leaving"][i]), float(tr[k]["step"][i])),
                                                  (int(ref_ent[i]), int(ref_lv[i]), float(ref.trace["step"][i]))) for i in bad[:6]])
        pivots += L
    assert pivots == out.pivots


@pytest.mark.parametrize("name", P.NETLIB)
def test_dual_devex_on_the_tableau_engine_reaches_the_same_optimum(env, name):
    prob, exp = P.netlib(name)
    O, N = env["O"], env["N"]
    res = _solver(env, "dual", engine=N.ENGINE_TABLEAU, block_k=16, pricing=N.PRICE_DEVEX).solve(prob)
    ref = O.solve(prob, O.DUAL, 1000, O.MODE_EXACT)
    assert res.kind == ref.status_name == "Optimal"
    assert _rel(res.solution.obj(), ref.obj) < 1e-9
    P.check_expectation(exp, res.kind, res.solution.obj(), res.solution.x())


def test_dual_devex_needs_fewer_pivots_on_a_dense_lp_and_agrees_with_the_reference_rule(env):
    O, S, N = env["O"], env["S"], env["N"]
    m, n = 192, 320
    A, b, c = _dense_lp(21, m, n)
    Af, cf, kind, lb, ub, start = _gte_dual_start(A, b, c)
    out = {}
    for tag, pricing in (("reference", N.PRICE_REFERENCE), ("devex", N.PRICE_DEVEX)):
        sg = [a.copy() for a in start]
        sol = S.GpuDualSimplexSolver.new(None, ctx=env["ctx"], engine=N.ENGINE_TABLEAU, block_k=16, pricing=pricing)
        res, _ = sol.solve_with_initial(m, n + m, Af, cf, b, kind, lb, ub, *sg)
        assert res.status == N.OPTIMAL
        np.testing.assert_allclose(Af @ sg[0], b, rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(sg[5], cf - Af.T @ sg[4], rtol=0, atol=1e-8)
        out[tag] = (res.iters, res.obj)
    st = [a.copy() for a in start]
    ref = O.solve_with_initial(O.DUAL, m, n + m, Af, cf, b, kind, lb, ub, *st, max_iter=None)
    assert _rel(out["reference"][1], ref.obj) < 1e-9 and _rel(out["devex"][1], ref.obj) < 1e-9
    assert out["devex"][0] < out["reference"][0], out
    print("dual pivots to optimal", out)


# ---------------------------------------------------------------- Devex pricing (primal) and complete solves against HiGHS
@pytest.mark.parametrize("name", P.NETLIB + ["small_prob_2", "small_prob_5", "beale_cycle"])
def test_primal_devex_on_the_tableau_engine_reaches_the_same_optimum(env, name):
    prob, exp = P.netlib(name) if name in P.NETLIB else getattr(P, name)()
    O, N = env["O"], env["N"]
    res = _solver(env, "primal", engine=N.ENGINE_TABLEAU, block_k=16, pricing=N.PRICE_DEVEX, tie_rule=N.TIES_CANONICAL).solve(prob)
    ref = O.solve(prob, O.PRIMAL, 1000, O.MODE_EXACT)
    assert res.kind == ref.status_name
    if res.is_optimal:
        assert _rel(res.solution.obj(), ref.obj) < 1e-9
    P.check_expectation(exp, res.kind, res.solution.obj() if res.is_optimal else float("nan"), res.solution.x() if res.is_optimal else [])


def _highs_fixture():
    import json, os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "highs_dense_lp.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("m,ns,refactor_every", [(512, 1024, 0), (4096, 8192, 1000)])
@pytest.mark.parametrize("variant,pricing", [(0, "dantzig"), (0, "devex"), (1, "reference"), (1, "devex")])
def test_complete_solve_of_the_dense_lp_matches_highs(env, m, ns, refactor_every, variant, pricing):
    """The bench generator's LP solved TO OPTIMALITY through the boundary on host buffers (slack start, so one phase), status and
    objective against HiGHS (tests/golden/highs_dense_lp.json, made by tests/golden/make_highs_dense_fixture.py on the numpy twin of
    the generator) to 1e-9, final primal residual |Ax - b| <= 1e-9 |b|.  The large case uploads the whole A (tuning key fast_upload = 0:
    the default condensed upload keeps only the nonbasic columns and cannot rebuild) and rebuilds the tableau from it every 1000
    pivots (B^-1 by the blocked DMMA LU, T = B^-1 A_N, x_B recomputed); tools/full_solve_stats.py shows that the same solves WITHOUT
    any rebuild (up to 3.6e5 pivots) agree with HiGHS to 1e-13 as well."""
    import bench_lp
    S, N = env["S"], env["N"]
    key = f"{m}x{ns}_seed0_variant{variant}"
    fx = _highs_fixture()
    if key not in fx:
        pytest.skip(f"no HiGHS fixture for {key}")
    lp = bench_lp.dense_lp(m, ns, 0, variant)
    st = [lp[k].copy() for k in ("x", "B", "N", "N_side")] + ([lp["y"].copy(), lp["d"].copy()] if variant else [])
    cls = S.GpuDualSimplexSolver if variant else S.GpuPrimalSimplexSolver
    sol = cls.new(None, ctx=env["ctx"], engine=N.ENGINE_TABLEAU, block_k=48, check_every=96, refactor_every=refactor_every,
                  pricing=N.PRICE_DEVEX if pricing == "devex" else N.PRICE_REFERENCE, tie_rule=N.TIES_CANONICAL)
    env["ctx"].set_tuning("fast_upload", 0 if refactor_every else 1)
    try:
        res, _ = sol.solve_with_initial(m, m + ns, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], *st)
    finally:
        env["ctx"].set_tuning("fast_upload", 1)
    assert res.status == N.OPTIMAL == fx[key]["status"]
    if refactor_every:
        # a rebuild happens at the first host read-back after `refactor_every` pivots (whole blocks of 48, read back every 96)
        assert res.refactors >= res.iters // (refactor_every + 96) - 1 and (res.iters < refactor_every + 96 or res.refactors > 0)
    x = st[0]
    obj = float(lp["c"] @ x)
    assert _rel(obj, fx[key]["obj"]) < 1e-9, (obj, fx[key]["obj"])
    assert np.abs(lp["A"] @ x - lp["b"]).max() <= 1e-9 * np.abs(lp["b"]).max()
    assert (x >= -1e-9).all()
    print(f"{key} {pricing}: {res.iters} pivots, {res.ms_device:.1f} ms on the device, {res.refactors} rebuilds, obj {obj!r} (HiGHS {fx[key]['obj']!r}, {fx[key]['nit']} its, {fx[key]['seconds']:.0f} s)")


def test_pipelined_host_batch_solve_equals_the_sequential_path(env):
    """ellp_b200_primal_solve_batch on host buffers: the chunked two-stream pipeline (H2D of chunk c + 1 under the kernel of chunk
    c) must return exactly what upload -> run -> download returns (tuning key batch_pipeline = 0)."""
    N, S, ctx = env["N"], env["S"], env["ctx"]
    nlp, m, ns, seed = 6000, 64, 128, 2
    n0 = ns + m
    ctx.check(N.lib.ellp_b200_batch_generate(ctx.h, nlp, m, ns, seed, 0, 0))
    A = np.zeros((nlp, n0, m)); c = np.zeros((nlp, n0)); b = np.zeros((nlp, m))
    ctx.check(N.lib.ellp_b200_batch_download_all(ctx.h, N.ptr(A), N.ptr(c), N.ptr(b)))
    kind = np.ones((nlp, n0), dtype=np.uint8); lb = np.zeros((nlp, n0)); ub = np.zeros((nlp, n0))
    out = {}
    for mode in (1, 0):
        ctx.set_tuning("batch_pipeline", mode)
        out[mode] = S.primal_solve_batch(A, c, b, kind, lb, ub, max_iter=None, ctx=ctx)
    ctx.set_tuning("batch_pipeline", 1)
    p, q = out[1], out[0]
    assert (p.status == N.OPTIMAL).all() and (p.err == 0).all()
    assert np.array_equal(p.status, q.status) and np.array_equal(p.iters, q.iters)
    assert p.obj.tobytes() == q.obj.tobytes() and p.x.tobytes() == q.x.tobytes()
    assert p.pivots == q.pivots == int(p.iters.sum())


@pytest.mark.parametrize("which", ["primal", "dual"])
def test_residual_triggered_rebuild_of_the_tableau(env, which):
    """Long solves on a rebuildable tableau check |A x - b|_inf every `residual_every` pivots and rebuild T = B^-1 A_N (and x_B) from A
    when it exceeds the tolerance.  Forced here with a zero tolerance: the rebuilds happen, the solve ends on the oracle's optimum,
    and with the default tolerance (1e-9 relative) a healthy solve triggers none."""
    O, S, N, ctx = env["O"], env["S"], env["N"], env["ctx"]
    m, n = 520, 700
    A, b, c = _dense_lp(31, m, n)
    if which == "dual":
        Af, cf, kind, lb, ub, start = _gte_dual_start(A, b, c)
        cls, oid = S.GpuDualSimplexSolver, O.DUAL
    else:
        Af, cf, kind, lb, ub, x0, B0, N0, Ns0 = _slack_start_primal(A, b, c)
        start = [x0, B0, N0, Ns0]
        cls, oid = S.GpuPrimalSimplexSolver, O.PRIMAL
    st = [a.copy() for a in start]
    ref = O.solve_with_initial(oid, m, n + m, Af, cf, b, kind, lb, ub, *st, max_iter=None, **({"mode": O.MODE_CANONICAL} if which == "primal" else {}))
    out = {}
    for tag, tol in (("forced", 0), ("default", 1000)):
        ctx.set_tuning("residual_every", 100)
        ctx.set_tuning("residual_tol_1e12", tol)
        try:
            sg = [a.copy() for a in start]
            sol = cls.new(None, ctx=ctx, engine=N.ENGINE_TABLEAU, block_k=16, check_every=32, tie_rule=N.TIES_CANONICAL)
            res, _ = sol.solve_with_initial(m, n + m, Af, cf, b, kind, lb, ub, *sg)
        finally:
            ctx.set_tuning("residual_every", -1)
            ctx.set_tuning("residual_tol_1e12", 1000)
        assert res.status == ref.status == O.OPTIMAL
        assert _rel(res.obj, ref.obj) < 1e-9
        np.testing.assert_allclose(sg[0], st[0], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(Af @ sg[0], b, rtol=1e-10, atol=1e-9)
        out[tag] = (int(res.refactors), int(res.iters))
    # refactors counts the initial build of the tableau (1) + the triggered rebuilds; a check happens at the first host read-back
    # (every 32 pivots) after 100 pivots since the last check, i.e. at most every 128 pivots
    assert out["forced"][0] >= 1 + out["forced"][1] // 160 and out["forced"][0] > 1, out
    assert out["default"][0] == 1, out


def test_small_cases_of_every_kernel_family(env):
    """tools/sanitize_cases.py (written for compute-sanitizer, which this pool keeps closed) as a plain regression: one small,
    self-checking invocation of every kernel family -- K6 batch, fused primal / dual pivot kernels + flushes, the three rank-k
    kernels at ragged shapes, blocked LU, single-CTA dual / Gauss-Jordan, rank-1."""
    import importlib, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    sc = importlib.import_module("sanitize_cases")
    for name, fn in sc.CASES.items():
        print(name, fn(env["ctx"]))


def test_dual_solve_that_starts_optimal_needs_no_device_work(env):
    """Loop head of the dual (dual :188-246): a starting point without any bound violation is Optimal before the first LU; the
    boundary answers from the host data (0 launches), with the reference's objective (dual_obj(y, d))."""
    O, S, N = env["O"], env["S"], env["N"]
    A = np.asfortranarray(np.array([[1.0, 2.0, 1.0, 0.0], [3.0, 1.0, 0.0, 1.0]])); c = np.array([1.0, 1.0, 0.0, 0.0]); b = np.array([4.0, 5.0])
    kind = np.array([N.LOWER] * 4, dtype=np.uint8); lb = np.zeros(4); ub = np.zeros(4)
    start = [np.array([0.0, 0.0, 4.0, 5.0]), np.array([2, 3], dtype=np.int32), np.array([0, 1], dtype=np.int32), np.array([0, 0], dtype=np.uint8), np.zeros(2), c.copy()]
    ref = O.solve_with_initial(O.DUAL, 2, 4, A, c, b, kind, lb, ub, *[a.copy() for a in start], max_iter=10)
    for engine in (N.ENGINE_REVISED, N.ENGINE_TABLEAU):
        sg = [a.copy() for a in start]
        res, _ = S.GpuDualSimplexSolver.new(10, ctx=env["ctx"], engine=engine, block_k=4).solve_with_initial(2, 4, A, c, b, kind, lb, ub, *sg)
        assert res.status == ref.status == O.OPTIMAL and res.iters == 0 and res.launches == 0
        assert res.obj == ref.obj
        for g, s0 in zip(sg, start):
            np.testing.assert_array_equal(g, s0)
    # max_iter = 0 still reports MaxIter first (dual :191-194)
    sg = [a.copy() for a in start]
    res, _ = S.GpuDualSimplexSolver.new(0, ctx=env["ctx"], engine=N.ENGINE_REVISED).solve_with_initial(2, 4, A, c, b, kind, lb, ub, *sg)
    assert res.status == N.MAXITER


@pytest.mark.parametrize("engine,bk", [("revised", 0), ("tableau", 0), ("tableau", 16)])
@pytest.mark.parametrize("name", P.NETLIB + ["small_prob_1", "small_prob_2", "small_prob_5", "beale_cycle", "small_prob_unbounded_1"])
def test_primal_harris_ratio_test_reaches_the_same_verdict(env, name, engine, bk):
    """ELLP_RATIO_HARRIS on the primal side (opt-in, no reference counterpart): same status as the oracle, same objective to 1e-9."""
    prob, exp = P.netlib(name) if name in P.NETLIB else getattr(P, name)()
    O, N = env["O"], env["N"]
    res = _solver(env, "primal", engine=N.ENGINE_REVISED if engine == "revised" else N.ENGINE_TABLEAU, block_k=bk, ratio=N.RATIO_HARRIS,
                  tie_rule=N.TIES_CANONICAL).solve(prob)
    ref = O.solve(prob, O.PRIMAL, 1000, O.MODE_EXACT)
    assert res.kind == ref.status_name
    if res.is_optimal:
        assert _rel(res.solution.obj(), ref.obj) < 1e-9
        assert prob.is_feasible(list(res.solution.x() + 0.0)) or np.allclose(res.solution.x(), ref.x, atol=1e-7)


def test_primal_harris_on_a_dense_lp_matches_highs(env):
    import bench_lp
    S, N = env["S"], env["N"]
    m, ns = 512, 1024
    fx = _highs_fixture()[f"{m}x{ns}_seed0_variant0"]
    lp = bench_lp.dense_lp(m, ns, 0, 0)
    for engine, bk in ((N.ENGINE_REVISED, 0), (N.ENGINE_TABLEAU, 32)):
        st = [lp[k].copy() for k in ("x", "B", "N", "N_side")]
        sol = S.GpuPrimalSimplexSolver.new(None, ctx=env["ctx"], engine=engine, block_k=bk, ratio=N.RATIO_HARRIS, tie_rule=N.TIES_CANONICAL)
        res, _ = sol.solve_with_initial(m, m + ns, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], *st)
        assert res.status == N.OPTIMAL
        obj = float(lp["c"] @ st[0])
        assert _rel(obj, fx["obj"]) < 1e-9
        assert np.abs(lp["A"] @ st[0] - lp["b"]).max() <= 1e-8 * np.abs(lp["b"]).max()
