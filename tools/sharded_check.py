"""Run under torchrun with N ranks (one GPU each; N = 1 works too): the column-sharded tableau engines must pivot exactly
like the oracle (order-free tie rule) and finish with the same point.  Checks both sharded engines:
  block_k = 0   the NCCL path (full tableau split by column, 3 collectives per pivot);
  block_k > 1   the peer-memory engine (condensed tableau split by nonbasic position, exchange fused into the pivot kernel),
and, on a size the oracle cannot reach (one LU per pivot), the peer engine against the single-GPU blocked engine.
Prints SHARDED_CHECK_OK on rank 0."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_lp  # noqa: E402
from ellp_b200 import _native as N  # noqa: E402
from ellp_b200 import sharded  # noqa: E402
from oracle import binding as O  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = N.Context(local)
sharded.init_comm(ctx, rank, world)
ok = True
for bk in (0, 8, 5):
    for (m, ns, seed, K) in [(64, 192, 3, 10**6), (256, 768, 4, 300)]:
        n = m + ns
        o = N.default_opts(K, engine=N.ENGINE_TABLEAU, tie_rule=N.TIES_CANONICAL, check_every=8, block_k=bk)
        tr = np.zeros(20000, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr); o.trace_cap = len(tr)
        # (1) generated in HBM, sharded
        ctx.check(N.lib.ellp_b200_sharded_generate_dense(ctx.h, m, ns, seed, C.byref(o)))
        res = N.Result()
        ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
        lp = bench_lp.dense_lp(m, ns, seed)
        x, B, Nv, Ns = lp["x"].copy(), lp["B"].copy(), lp["N"].copy(), lp["N_side"].copy()
        pt = N.Point(N.ptr(x), N.ptr(B), N.ptr(Nv), N.ptr(Ns), None, None, m, ns)
        ctx.check(N.lib.ellp_b200_download(ctx.h, C.byref(pt)))
        xo, Bo, No, Nso = lp["x"].copy(), lp["B"].copy(), lp["N"].copy(), lp["N_side"].copy()
        ref = O.solve_with_initial(O.PRIMAL, m, n, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], xo, Bo, No, Nso,
                                   max_iter=K, mode=O.MODE_CANONICAL, trace_cap=20000)
        k = len(ref.trace)
        good = (res.status == ref.status and res.iters == k and (tr["entering"][:k] == ref.trace["entering"]).all()
                and (tr["leaving"][:k] == ref.trace["leaving"]).all() and np.array_equal(B, Bo)
                and np.allclose(x, xo, rtol=1e-9, atol=1e-9) and set(Nv.tolist()) == set(No.tolist()))
        if bk > 1:  # the peer engine keeps the N list in position order like the reference (and the oracle)
            good = good and np.array_equal(Nv, No) and np.array_equal(Ns, Nso)
        # (2) the same LP uploaded from host buffers, each rank passing ITS block
        x2, B2, N2, Ns2 = lp["x"].copy(), lp["B"].copy(), lp["N"].copy(), lp["N_side"].copy()
        pt2 = N.Point(N.ptr(x2), N.ptr(B2), N.ptr(N2), N.ptr(Ns2), None, None, m, ns)
        tr2 = np.zeros(20000, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr2)
        if bk > 1:  # nonbasic columns of this rank's positions, in N order
            plo, phi = sharded.shard_range(ns, world, rank)
            Aloc = np.asfortranarray(lp["A"][:, lp["N"][plo:phi]])
            sf = N.StdForm(m, n, N.ptr(Aloc), N.ptr(lp["c"]), N.ptr(lp["b"]), N.ptr(lp["kind"]), N.ptr(lp["lb"]), N.ptr(lp["ub"]))
            ctx.check(N.lib.ellp_b200_sharded_upload_nonbasic(ctx.h, C.byref(sf), C.byref(pt2), C.byref(o)))
        else:
            lo, hi = sharded.shard_range(n, world, rank)
            Aloc = np.asfortranarray(lp["A"][:, lo:hi])
            sf = N.StdForm(m, n, N.ptr(Aloc), N.ptr(lp["c"]), N.ptr(lp["b"]), N.ptr(lp["kind"]), N.ptr(lp["lb"]), N.ptr(lp["ub"]))
            ctx.check(N.lib.ellp_b200_sharded_upload(ctx.h, C.byref(sf), C.byref(pt2), C.byref(o)))
        res2 = N.Result()
        ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res2)))
        ctx.check(N.lib.ellp_b200_download(ctx.h, C.byref(pt2)))
        good2 = (res2.iters == k and (tr2["entering"][:k] == ref.trace["entering"]).all() and np.array_equal(B2, Bo)
                 and np.allclose(x2, xo, rtol=1e-9, atol=1e-9))
        print(f"rank {rank}: block_k={bk} m={m} n={n} pivots={res.iters} oracle={k} status={res.status}/{ref.status} "
              f"generated_ok={good} uploaded_ok={good2}", flush=True)
        ok = ok and good and good2

# (2b) near-ties ACROSS ranks: every structural column appears twice, once in each half of N, so the two copies have the
# same Dantzig key on different ranks (world >= 2) -- the second mailbox round (order-free rule: largest variable index)
# decides, and must decide like the oracle's canonical mode.  Also exercises ties in the ratio test (duplicate rows).
for bk in (8, 64):
    m, half, seed, K = 64, 96, 11, 400
    base = bench_lp.dense_lp(m, half, seed)
    ns = 2 * half
    n = m + ns
    A = np.zeros((m, n), order="F")
    A[:, :half] = base["A"][:, :half]; A[:, half:ns] = base["A"][:, :half]; A[:, ns:] = np.eye(m)
    A[m // 2:, :ns] = A[:m - m // 2, :ns]                       # duplicate rows => ties in the ratio test too
    c = np.concatenate([base["c"][:half], base["c"][:half], np.zeros(m)])
    b = base["b"].copy(); b[m // 2:] = b[:m - m // 2]
    kind = np.ones(n, dtype=np.uint8); lb = np.zeros(n); ub = np.zeros(n)
    x0 = np.concatenate([np.zeros(ns), b]); B0 = np.arange(ns, n, dtype=np.int32); N0 = np.arange(ns, dtype=np.int32); Ns0 = np.zeros(ns, dtype=np.uint8)
    xo, Bo, No, Nso = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
    ref = O.solve_with_initial(O.PRIMAL, m, n, A, c, b, kind, lb, ub, xo, Bo, No, Nso, max_iter=K, mode=O.MODE_CANONICAL, trace_cap=K)
    o = N.default_opts(K, engine=N.ENGINE_TABLEAU, tie_rule=N.TIES_CANONICAL, check_every=8, block_k=bk)
    tr = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr); o.trace_cap = K
    plo, phi = sharded.shard_range(ns, world, rank)
    Aloc = np.asfortranarray(A[:, N0[plo:phi]])
    sf = N.StdForm(m, n, N.ptr(Aloc), N.ptr(c), N.ptr(b), N.ptr(kind), N.ptr(lb), N.ptr(ub))
    xg, Bg, Ng, Nsg = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
    pt = N.Point(N.ptr(xg), N.ptr(Bg), N.ptr(Ng), N.ptr(Nsg), None, None, m, ns)
    ctx.check(N.lib.ellp_b200_sharded_upload_nonbasic(ctx.h, C.byref(sf), C.byref(pt), C.byref(o)))
    res = N.Result()
    ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
    ctx.check(N.lib.ellp_b200_download(ctx.h, C.byref(pt)))
    k = len(ref.trace)
    good = (res.status == ref.status and res.iters == k and (tr["entering"][:k] == ref.trace["entering"]).all()
            and (tr["leaving"][:k] == ref.trace["leaving"]).all() and np.array_equal(Bg, Bo) and np.array_equal(Ng, No)
            and np.allclose(xg, xo, rtol=1e-9, atol=1e-9))
    print(f"rank {rank}: duplicated columns/rows (cross-rank near-ties) block_k={bk} pivots={res.iters} oracle={k} status={res.status}/{ref.status} ok={good}", flush=True)
    ok = ok and good

# (3) a size with many CTAs per rank (grid barriers, last-block tickets, many column words on the wire): the peer engine
# against the single-GPU blocked engine (itself checked against the oracle at small sizes), same LP, same tie rule.
m, ns, seed, K, bk = 2048, 4096, 5, 160, 32
o = N.default_opts(K, engine=N.ENGINE_TABLEAU, tie_rule=N.TIES_CANONICAL, check_every=32, block_k=bk)
tr_s = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr_s); o.trace_cap = K
single = N.Context(local)
single.check(N.lib.ellp_b200_generate_dense(single.h, m, ns, seed, C.byref(o)))
rs = N.Result()
single.check(N.lib.ellp_b200_run(single.h, C.byref(o), C.byref(rs)))
xs = np.zeros(m + ns); Bs = np.zeros(m, dtype=np.int32); Nvs = np.zeros(ns, dtype=np.int32); Nss = np.zeros(ns, dtype=np.uint8)
single.check(N.lib.ellp_b200_download(single.h, C.byref(N.Point(N.ptr(xs), N.ptr(Bs), N.ptr(Nvs), N.ptr(Nss), None, None, m, ns))))
single.close()
tr_p = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr_p)
ctx.check(N.lib.ellp_b200_sharded_generate_dense(ctx.h, m, ns, seed, C.byref(o)))
rp = N.Result()
ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(rp)))
xp = np.zeros(m + ns); Bp = np.zeros(m, dtype=np.int32); Nvp = np.zeros(ns, dtype=np.int32); Nsp = np.zeros(ns, dtype=np.uint8)
ctx.check(N.lib.ellp_b200_download(ctx.h, C.byref(N.Point(N.ptr(xp), N.ptr(Bp), N.ptr(Nvp), N.ptr(Nsp), None, None, m, ns))))
good3 = (rp.status == rs.status and rp.iters == rs.iters == K and (tr_p["entering"] == tr_s["entering"]).all()
         and (tr_p["leaving"] == tr_s["leaving"]).all() and np.array_equal(Bp, Bs) and np.array_equal(Nvp, Nvs)
         and np.array_equal(Nsp, Nss) and xp.tobytes() == xs.tobytes() and rp.obj == rs.obj)
print(f"rank {rank}: peer vs single-GPU blocked engine m={m} n={m + ns} pivots={rp.iters}/{rs.iters} identical={good3} "
      f"peer_ms={rp.ms_device:.2f} single_ms={rs.ms_device:.2f}", flush=True)
ok = ok and good3

# (4) DUAL simplex on the peer engine (dual_blocked.cuh): generated in HBM and uploaded from host buffers (slack basis -I passed as
# basis_diag), against the oracle's DualSimplexSolver loop: same pivots, x, y, d, B, N
for bk in (8, 5):
    for (m, ns, seed, K) in [(64, 192, 3, 10**6), (256, 768, 4, 300)]:
        n = m + ns
        lp = bench_lp.dense_lp(m, ns, seed, 1)
        st = [lp[k].copy() for k in ("x", "B", "N", "N_side", "y", "d")]
        ref = O.solve_with_initial(O.DUAL, m, n, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], *st, max_iter=K, trace_cap=20000)
        k = len(ref.trace)
        o = N.default_opts(K, engine=N.ENGINE_TABLEAU, check_every=8, block_k=bk)
        oks = []
        for mode in ("generated", "uploaded"):
            tr = np.zeros(20000, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr); o.trace_cap = len(tr)
            g = [lp[q].copy() for q in ("x", "B", "N", "N_side", "y", "d")]
            pt = N.Point(N.ptr(g[0]), N.ptr(g[1]), N.ptr(g[2]), N.ptr(g[3]), N.ptr(g[4]), N.ptr(g[5]), m, ns)
            if mode == "generated":
                ctx.check(N.lib.ellp_b200_sharded_generate_dense_ex(ctx.h, m, ns, seed, 1, C.byref(o)))
            else:
                plo, phi = sharded.shard_range(ns, world, rank)
                Aloc = np.asfortranarray(lp["A"][:, lp["N"][plo:phi]])
                sf = N.StdForm(m, n, N.ptr(Aloc), N.ptr(lp["c"]), N.ptr(lp["b"]), N.ptr(lp["kind"]), N.ptr(lp["lb"]), N.ptr(lp["ub"]))
                diag = -np.ones(m)
                ctx.check(N.lib.ellp_b200_sharded_upload_nonbasic_ex(ctx.h, C.byref(sf), C.byref(pt), N.DUAL, N.ptr(diag), C.byref(o)))
            res = N.Result()
            ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
            ctx.check(N.lib.ellp_b200_download(ctx.h, C.byref(pt)))
            good = (res.status == ref.status and res.iters == k and (tr["entering"][:k] == ref.trace["entering"]).all()
                    and (tr["leaving"][:k] == ref.trace["leaving"]).all() and np.array_equal(g[1], st[1]) and np.array_equal(g[2], st[2])
                    and np.array_equal(g[3], st[3]) and np.allclose(g[0], st[0], rtol=1e-9, atol=1e-9)
                    and np.allclose(g[4], st[4], rtol=1e-9, atol=1e-9) and np.allclose(g[5], st[5], rtol=1e-9, atol=1e-9)
                    and abs(res.obj - ref.obj) <= 1e-9 * max(1.0, abs(ref.obj)))
            oks.append(good)
        print(f"rank {rank}: DUAL peer engine block_k={bk} m={m} n={n} pivots={res.iters} oracle={k} status={res.status}/{ref.status} "
              f"generated_ok={oks[0]} uploaded_ok={oks[1]}", flush=True)
        ok = ok and all(oks)

# (5) the dual at a size with many CTAs per rank: peer engine vs the single-GPU dual tableau engine, bit for bit
m, ns, seed, K, bk = 2048, 4096, 5, 160, 32
o = N.default_opts(K, engine=N.ENGINE_TABLEAU, check_every=32, block_k=bk)
tr_s = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr_s); o.trace_cap = K
single = N.Context(local)
single.check(N.lib.ellp_b200_generate_dense_ex(single.h, m, ns, seed, 1, C.byref(o)))
rs = N.Result()
single.check(N.lib.ellp_b200_run(single.h, C.byref(o), C.byref(rs)))


def _dl(c_):
    x = np.zeros(m + ns); B = np.zeros(m, dtype=np.int32); Nv = np.zeros(ns, dtype=np.int32); Ns = np.zeros(ns, dtype=np.uint8)
    y = np.zeros(m); d = np.zeros(m + ns)
    c_.check(N.lib.ellp_b200_download(c_.h, C.byref(N.Point(N.ptr(x), N.ptr(B), N.ptr(Nv), N.ptr(Ns), N.ptr(y), N.ptr(d), m, ns))))
    return x, B, Nv, Ns, y, d


ds = _dl(single)
single.close()
tr_p = np.zeros(K, dtype=N.TRACE_DTYPE); o.trace = N.ptr(tr_p)
ctx.check(N.lib.ellp_b200_sharded_generate_dense_ex(ctx.h, m, ns, seed, 1, C.byref(o)))
rp = N.Result()
ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(rp)))
dp = _dl(ctx)
good5 = (rp.status == rs.status and rp.iters == rs.iters == K and (tr_p["entering"] == tr_s["entering"]).all()
         and (tr_p["leaving"] == tr_s["leaving"]).all() and rp.obj == rs.obj and all(a.tobytes() == b.tobytes() for a, b in zip(dp, ds)))
print(f"rank {rank}: DUAL peer vs single-GPU dual tableau engine m={m} n={m + ns} pivots={rp.iters}/{rs.iters} identical={good5} "
      f"peer_ms={rp.ms_device:.2f} single_ms={rs.ms_device:.2f}", flush=True)
ok = ok and good5

flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0 and int(flag.item()) == 1:
    print("SHARDED_CHECK_OK", flush=True)
dist.barrier()
ctx.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
