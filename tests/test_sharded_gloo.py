"""world_size-2 CPU test (gloo) of the N > 1 path's host logic: the column partition and the arg-select / pivot-column
exchange protocol of the sharded tableau (ellp_b200/sharded.py), emulated on numpy shards and checked pivot for pivot
against the oracle's order-free mode.  The CUDA kernels themselves are covered by tools/sharded_check.py on GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = 1e-10


def _worker(rank, world, port, m, ns, seed, K, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench_lp
    from ellp_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lp = bench_lp.dense_lp(m, ns, seed)
    n = m + ns
    lo, hi = sharded.shard_range(n, world, rank)
    assert sharded.owner_of(lo, n, world) == rank and sharded.owner_of(hi - 1, n, world) == rank
    T = lp["A"][:, lo:hi].copy()                       # local column block of the tableau (identity basis => T = A)
    dj = lp["c"][lo:hi].copy()                         # local slice of the reduced-cost row (c_B = 0)
    stat = np.where(np.arange(lo, hi) < ns, 0, 3).astype(np.uint8)  # 0 = at lower, 3 = basic
    x, Bv = lp["x"].copy(), lp["B"].copy()
    trace = []
    for it in range(K):
        key = np.where((stat == 0) & (dj < 0) & (np.abs(dj) >= EPS), -dj, -1.0)
        kmax_loc = torch.tensor([key.max() if len(key) else -1.0], dtype=torch.float64)
        g1 = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(g1, kmax_loc)
        kmax = max(float(t) for t in g1)
        cand = np.nonzero((key != -1.0) & (kmax - key < EPS))[0]
        mine = (-1.0, 0.0, 0.0) if len(cand) == 0 else (float(lo + cand.max()), float(dj[cand.max()]), float(stat[cand.max()]))
        g2 = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(g2, torch.tensor(mine, dtype=torch.float64))
        win = sharded.pick_entering([float(t) for t in g1], [tuple(t.tolist()) for t in g2])
        if win is None:
            break
        q, rq = int(win[0]), win[1]
        col = torch.zeros(m, dtype=torch.float64)
        if lo <= q < hi:
            col = torch.from_numpy(T[:, q - lo].copy())
        dist.all_reduce(col)                           # owner contributes the pivot column, the others zeros
        alpha = col.numpy()
        d = -alpha                                     # entering from its lower bound
        lam = np.full(m, np.inf)
        act = (np.abs(d) >= EPS) & (d < 0)
        lam[act] = np.where(x[Bv[act]] > 0, (0 - x[Bv[act]]) / d[act], 0.0)
        lmin = lam.min()
        if not np.isfinite(lmin):
            break
        tie = np.nonzero(lam - lmin < EPS)[0]
        r = tie[np.argmin(Bv[tie])]
        lam_r = lam[r]
        x[Bv] += lam_r * d
        x[q] += lam_r
        leave = int(Bv[r])
        trace.append((q, leave))
        prow = T[r, :] / alpha[r]
        T -= np.outer(alpha, prow)
        T[r, :] = prow
        dj -= rq * prow
        Bv[r] = q
        if lo <= q < hi:
            stat[q - lo] = 3
        if lo <= leave < hi:
            stat[leave - lo] = 0
    ret[rank] = (trace, x, Bv)
    dist.destroy_process_group()


@pytest.mark.parametrize("m,ns,seed,K", [(16, 48, 1, 40), (32, 96, 2, 60)])
def test_sharded_protocol_matches_oracle_world2(m, ns, seed, K):
    sys.path.insert(0, ROOT)
    import bench_lp
    from oracle import binding as O
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, m, ns, seed, K, ret), nprocs=world, join=True)
    lp = bench_lp.dense_lp(m, ns, seed)
    xo, Bo = lp["x"].copy(), lp["B"].copy()
    ref = O.solve_with_initial(O.PRIMAL, m, m + ns, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], xo, Bo,
                               lp["N"].copy(), lp["N_side"].copy(), max_iter=K, mode=O.MODE_CANONICAL, trace_cap=K)
    want = list(zip(ref.trace["entering"].tolist(), ref.trace["leaving"].tolist()))
    for rank in range(world):
        trace, x, Bv = ret[rank]
        assert trace == want
        np.testing.assert_array_equal(Bv, Bo)
        np.testing.assert_allclose(x, xo, rtol=1e-9, atol=1e-9)


def _peer_worker(rank, world, port, m, ns, seed, K, bk, ret):
    """Peer-memory engine protocol (ellp_b200/csrc/peer.cuh) on numpy shards over gloo: condensed tableau split by
    nonbasic POSITION, deferred rank-k updates (U replicated, V local), pricing mailbox = all_gather of 4 doubles per
    rank, near-tie second round, pivot column 'stored into every rank' = broadcast from the owner."""
    sys.path.insert(0, ROOT)
    import bench_lp
    from ellp_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lp = bench_lp.dense_lp(m, ns, seed)
    n = m + ns
    plo, phi = sharded.shard_range(ns, world, rank)
    nT = phi - plo
    Nv, Ns = lp["N"].copy(), lp["N_side"].copy()     # replicated N list (position order)
    T = lp["A"][:, Nv[plo:phi]].copy()                # local positions of the condensed tableau (identity basis => T = A_N)
    dj = lp["c"][Nv[plo:phi]].copy()
    x, Bv, c = lp["x"].copy(), lp["B"].copy(), lp["c"]
    U = np.zeros((m, bk)); V = np.zeros((bk, nT)); fill = 0
    trace = []
    for it in range(K):
        side = Ns[plo:phi]
        key = np.where((np.abs(dj) >= EPS) & (dj < 0) & (side == 0), -dj, -1.0)
        order = np.argsort(-key, kind="stable")
        a1 = key[order[0]]
        a2 = key[order[1]] if nT > 1 else -1.0
        msg = torch.tensor([a1, a2, float(plo + order[0]) if a1 != -1.0 else -1.0, dj[order[0]] if a1 != -1.0 else 0.0], dtype=torch.float64)
        box = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(box, msg)
        status, q_pos, rq = sharded.merge_pricing([tuple(b.tolist()) for b in box])
        if status == "optimal":
            break
        if status == "near_tie":
            kmax = max(float(b[0]) for b in box)
            cand = np.nonzero((key != -1.0) & (kmax - key < EPS))[0]
            mine = (-1.0, -1.0, 0.0)
            if len(cand):
                j = cand[np.argmax(Nv[plo + cand])]
                mine = (float(Nv[plo + j]), float(plo + j), float(dj[j]))
            box2 = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(box2, torch.tensor(mine, dtype=torch.float64))
            q_pos, rq = sharded.merge_near_tie([tuple(b.tolist()) for b in box2])
        q_var = int(Nv[q_pos])
        owner = q_pos // nT
        col = torch.zeros(m, dtype=torch.float64)
        if owner == rank:                              # stale column + pending corrections, in pivot order
            a = T[:, q_pos - plo].copy()
            for j in range(fill):
                a = a - U[:, j] * V[j, q_pos - plo]
            col = torch.from_numpy(a)
        dist.broadcast(col, src=owner)
        alpha = col.numpy()
        d = -alpha                                     # entering from its lower bound
        lam = np.full(m, np.inf)
        act = (np.abs(d) >= EPS) & (d < 0)
        lam[act] = np.where(x[Bv[act]] > 0, (0 - x[Bv[act]]) / d[act], 0.0)
        lmin = lam.min()
        if not np.isfinite(lmin):
            break
        tie = np.nonzero(lam - lmin < EPS)[0]
        r = tie[np.argmin(Bv[tie])]
        lam_r = lam[r]
        x[Bv] += lam_r * d
        x[q_var] += lam_r
        leave = int(Bv[r])
        trace.append((q_var, leave))
        # current pivot row of the local positions (stale row + pending corrections), scaled
        e = T[r, :].copy()
        for j in range(fill):
            e = e - U[r, j] * V[j, :]
        p = e / alpha[r]
        if owner == rank:                              # the entering position is handed to the leaving variable (column e_r)
            ql = q_pos - plo
            p[ql] = 1.0 / alpha[r]
            dj[ql] = 0.0
            T[:, ql] = 0.0; T[r, ql] = 1.0
            V[:fill, ql] = 0.0
        dj -= rq * p
        u = alpha.copy(); u[r] -= 1.0
        U[:, fill] = u; V[fill, :] = p; fill += 1
        Bv[r] = q_var; Nv[q_pos] = leave; Ns[q_pos] = 0 if d[r] <= 0 else 1
        if fill == bk:                                 # flush: T -= U V (k_blk_flush*)
            T -= U @ V
            U[:] = 0.0; V[:] = 0.0; fill = 0
    ret[rank] = (trace, x, Bv, Nv)
    dist.destroy_process_group()


@pytest.mark.parametrize("m,ns,seed,K,bk", [(16, 48, 1, 40, 4), (32, 96, 2, 60, 8)])
def test_peer_protocol_matches_oracle_world2(m, ns, seed, K, bk):
    sys.path.insert(0, ROOT)
    import bench_lp
    from oracle import binding as O
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_peer_worker, args=(world, port, m, ns, seed, K, bk, ret), nprocs=world, join=True)
    lp = bench_lp.dense_lp(m, ns, seed)
    xo, Bo, No = lp["x"].copy(), lp["B"].copy(), lp["N"].copy()
    ref = O.solve_with_initial(O.PRIMAL, m, m + ns, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], xo, Bo,
                               No, lp["N_side"].copy(), max_iter=K, mode=O.MODE_CANONICAL, trace_cap=K)
    want = list(zip(ref.trace["entering"].tolist(), ref.trace["leaving"].tolist()))
    for rank in range(world):
        trace, x, Bv, Nv = ret[rank]
        assert trace == want
        np.testing.assert_array_equal(Bv, Bo)
        np.testing.assert_array_equal(Nv, No)          # the N list keeps the reference's position order
        np.testing.assert_allclose(x, xo, rtol=1e-9, atol=1e-9)


def _peer_dual_worker(rank, world, port, m, ns, seed, K, bk, ret):
    """DUAL simplex on the peer layout (ellp_b200/csrc/dual_blocked.cuh) on numpy shards over gloo: leaving row from the
    replicated x (no exchange), pivot row + ratios per shard, mailbox merge = all_gather of 4 doubles per rank
    (sharded.merge_dual_entering), entering column broadcast from its owner AFTER the local part of the update, deferred rank-k
    updates, cancellation snap of the reduced costs."""
    sys.path.insert(0, ROOT)
    import bench_lp
    from ellp_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lp = bench_lp.dense_lp(m, ns, seed, 1)
    plo, phi = sharded.shard_range(ns, world, rank)
    nT = phi - plo
    Nv, Ns = lp["N"].copy(), lp["N_side"].copy()
    T = -lp["A"][:, Nv[plo:phi]].copy()               # slack basis -I: T = B^-1 A_N = -A_N
    dj = lp["d"][Nv[plo:phi]].copy()
    x, Bv = lp["x"].copy(), lp["B"].copy()
    U = np.zeros((m, bk)); V = np.zeros((bk, nT)); fill = 0
    trace = []
    status = "maxiter"
    for it in range(K):
        viol = np.nonzero(x[Bv] < 0 - EPS)[0]          # every variable is Lower(0): first infeasible position (dual :200-236)
        if len(viol) == 0:
            status = "optimal"
            break
        r = int(viol[0])
        delta = x[Bv[r]] - 0.0
        neg = delta < 0
        a_raw = T[r, :].copy()                         # pivot row of the local positions: stale row + pending corrections
        for j in range(fill):
            a_raw = a_raw - U[r, j] * V[j, :]
        a = -a_raw if neg else a_raw
        side = Ns[plo:phi]
        keep = np.where(side == 0, a > EPS, a < -EPS)
        mine = (0.0, 0.0, -1.0, 0.0)
        if keep.any():
            ratio = np.full(nT, np.inf)
            ratio[keep] = dj[keep] / a[keep]
            ratio[ratio == 0.0] = 0.0
            jl = int(np.lexsort((np.arange(nT), ratio))[0])
            mine = (float(ratio[jl]), float(np.isnan(ratio[keep]).any()), float(plo + jl), float(a_raw[jl]))
        box = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(box, torch.tensor(mine, dtype=torch.float64))
        what, q_pos, ratio_q, alpha_rq = sharded.merge_dual_entering([tuple(b.tolist()) for b in box])
        if what != "pick":
            status = what
            break
        theta_d = -ratio_q if neg else ratio_q
        theta_p = delta / alpha_rq
        owner = q_pos // nT
        ql = q_pos - plo
        # E: local part of the update, before the column is needed
        p = a_raw / alpha_rq
        prod = theta_d * a_raw
        dnew = dj - prod
        dnew[np.abs(dnew) <= 16 * 2.220446049250313e-16 * np.maximum(np.abs(dj), np.abs(prod))] = 0.0
        col = torch.zeros(m, dtype=torch.float64)
        if owner == rank:
            acol = T[:, ql].copy()
            for j in range(fill):
                acol = acol - U[:, j] * V[j, ql]
            col = torch.from_numpy(acol)
            p[ql] = 1.0 / alpha_rq
            dnew[ql] = -theta_d
            T[:, ql] = 0.0; T[r, ql] = 1.0
            V[:fill, ql] = 0.0
        dj = dnew
        dist.broadcast(col, src=owner)
        alpha_q = col.numpy()
        assert alpha_q[r] == alpha_rq                  # the row-derived and the column-derived pivot element are the same bits
        x[Bv] = x[Bv] - theta_p * alpha_q
        q_var, leave = int(Nv[q_pos]), int(Bv[r])
        x[q_var] = x[q_var] + theta_p
        trace.append((q_var, leave))
        u = alpha_q.copy(); u[r] -= 1.0
        U[:, fill] = u; V[fill, :] = p; fill += 1
        Bv[r] = q_var; Nv[q_pos] = leave; Ns[q_pos] = 0 if neg else 1
        if fill == bk:
            T -= U @ V
            U[:] = 0.0; V[:] = 0.0; fill = 0
    ret[rank] = (trace, x, Bv, Nv, status)
    dist.destroy_process_group()


@pytest.mark.parametrize("m,ns,seed,K,bk", [(16, 48, 1, 400, 4), (32, 96, 2, 60, 8)])
def test_peer_dual_protocol_matches_oracle_world2(m, ns, seed, K, bk):
    sys.path.insert(0, ROOT)
    import bench_lp
    from oracle import binding as O
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29700 + (os.getpid() % 90)
    mp.spawn(_peer_dual_worker, args=(world, port, m, ns, seed, K, bk, ret), nprocs=world, join=True)
    lp = bench_lp.dense_lp(m, ns, seed, 1)
    st = [lp[k].copy() for k in ("x", "B", "N", "N_side", "y", "d")]
    ref = O.solve_with_initial(O.DUAL, m, m + ns, lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"], lp["ub"], *st, max_iter=K, trace_cap=K)
    want = list(zip(ref.trace["entering"].tolist(), ref.trace["leaving"].tolist()))
    assert len(want) > 5
    for rank in range(world):
        trace, x, Bv, Nv, status = ret[rank]
        assert trace == want
        assert (status == "optimal") == (ref.status == O.OPTIMAL)
        np.testing.assert_array_equal(Bv, st[1])
        np.testing.assert_array_equal(Nv, st[2])
        np.testing.assert_allclose(x, st[0], rtol=1e-9, atol=1e-9)


def test_dual_entering_merge_rules():
    from ellp_b200 import sharded
    none = (0.0, 0.0, -1.0, 0.0)
    assert sharded.merge_dual_entering([none, none])[0] == "infeasible"
    assert sharded.merge_dual_entering([(0.5, 0.0, 7.0, 2.0), (0.25, 0.0, 40.0, -1.0)]) == ("pick", 40, 0.25, -1.0)
    # equal ratios: the smaller POSITION wins whatever the rank order (Iterator::min_by keeps the first minimum)
    assert sharded.merge_dual_entering([(0.5, 0.0, 41.0, 2.0), (0.5, 0.0, 7.0, 3.0)]) == ("pick", 7, 0.5, 3.0)
    assert sharded.merge_dual_entering([(0.5, 0.0, 7.0, 3.0), (0.5, 0.0, 41.0, 2.0)]) == ("pick", 7, 0.5, 3.0)
    assert sharded.merge_dual_entering([(0.5, 1.0, 7.0, 3.0), none])[0] == "nan"
    assert sharded.merge_dual_entering([(float("inf"), 0.0, 3.0, 0.0), none]) == ("pick", 3, float("inf"), 0.0)


def test_pricing_merge_rules():
    from ellp_b200 import sharded
    none = (-1.0, -1.0, -1.0, 0.0)
    assert sharded.merge_pricing([none, none])[0] == "optimal"
    assert sharded.merge_pricing([(3.0, 1.0, 5.0, -3.0), (2.0, 0.5, 70.0, -2.0)]) == ("pick", 5, -3.0)
    assert sharded.merge_pricing([(2.0, 0.5, 5.0, -2.0), (3.0, 1.0, 70.0, -3.0)]) == ("pick", 70, -3.0)
    # equal maxima on two ranks, or a second best within 2 EPS on the same rank: second round
    assert sharded.merge_pricing([(3.0, 1.0, 5.0, -3.0), (3.0, 0.5, 70.0, -3.0)])[0] == "near_tie"
    assert sharded.merge_pricing([(3.0, 3.0 - 1e-11, 5.0, -3.0), none])[0] == "near_tie"
    assert sharded.merge_near_tie([(12.0, 5.0, -3.0), (40.0, 70.0, -3.0), (-1.0, -1.0, 0.0)]) == (70, -3.0)


def test_shard_range_partition():
    from ellp_b200 import sharded
    for world in (1, 2, 4, 8):
        covered = []
        for r in range(world):
            lo, hi = sharded.shard_range(65536, world, r)
            covered += [(lo, hi)]
        assert covered[0][0] == 0 and covered[-1][1] == 65536
        assert all(covered[i][1] == covered[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        sharded.shard_range(10, 4, 0)
