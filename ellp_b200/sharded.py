"""Column-sharded tableau: one process per GPU, launched with torchrun (reference config: BASELINE.json configs[4]).

torch.distributed is only the plumbing (rendezvous, shipping the NCCL unique id, max-over-ranks of the timings);
the per-pivot exchange -- two 8..24-byte all-gathers for the arg-select and one all-reduce that broadcasts the pivot
column from its owner -- is issued by libellp_b200.so on its own stream through NCCL (include/ellp_b200.h).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from typing import Tuple

import numpy as np

from . import _native as N


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Global column range [lo, hi) stored by `rank` (equal blocks; n must be divisible by world)."""
    if n % world != 0:
        raise ValueError(f"column sharding needs n ({n}) divisible by the number of ranks ({world})")
    w = n // world
    return rank * w, (rank + 1) * w


def owner_of(col: int, n: int, world: int) -> int:
    return col // (n // world)


def pick_entering(local_max_keys, candidates, eps: float = 1e-10):
    """Host restatement of the device protocol (k_shard_pick / k_shard_stage_column), used by the CPU tests.

    local_max_keys[g]: largest Dantzig key on rank g (-1 when it has no candidate).
    candidates[g]: (var, rq, side) = rank g's largest variable index whose key is within eps of the GLOBAL maximum,
    or (-1, 0, 0).  Returns (var, rq, side) of the entering variable or None when no rank has a candidate.
    """
    kmax = max(local_max_keys)
    if kmax == -1.0:
        return None
    best = max(candidates, key=lambda c: c[0])
    return best if best[0] >= 0 else None


# ---- peer-memory engine (ellp_b200/csrc/peer.cuh): host restatement of the per-pivot protocol, used by the CPU tests ----
def top2_merge_max(t, o):
    """(best, second, index) merge of k_blk_pivots_fused phase B (top2_merge<true>): `o` wins only when strictly larger;
    an equal best from another source becomes the SECOND, which is what flags the near-tie."""
    a1, a2, i1 = t
    b1, b2, j1 = o
    if b1 > a1:
        return (b1, a1 if a1 > b2 else b2, j1)
    return (a1, b1 if b1 > a2 else a2, i1)


def merge_pricing(msgs, eps: float = 1e-10):
    """msgs[g] = (best key, second best key, global position of the best, reduced cost of the best) from rank g, or
    (-1, -1, -1, 0) when rank g has no candidate.  Returns (status, position, reduced cost):
    'optimal' (no candidate anywhere), 'pick' (isolated maximum) or 'near_tie' (second round needed)."""
    t = (-1.0, -1.0, -1)
    rq = 0.0
    for (a1, a2, pos, r) in msgs:
        if a1 > t[0]:
            rq = r
        t = top2_merge_max(t, (a1, a2, int(pos)))
    if t[0] == -1.0:
        return "optimal", -1, 0.0
    if t[0] - t[1] < 2.0 * eps:
        return "near_tie", t[2], rq
    return "pick", t[2], rq


def merge_near_tie(cands):
    """Second round: cands[g] = (variable index, global position, reduced cost) of rank g's largest variable index whose
    key lies within EPS of the global maximum, or (-1, -1, 0).  The largest variable index wins."""
    best = max(cands, key=lambda c: c[0])
    return int(best[1]), best[2]


def merge_dual_entering(msgs):
    """Entering position of the sharded DUAL (dual_blocked.cuh, phase R): msgs[g] = (ratio, nan flag, global position or -1,
    pivot-row entry) of rank g's lexicographic minimum of (d_j / alpha~_j, position) over its eligible positions.  Returns
    ('nan' | 'infeasible' | 'pick', position, ratio, pivot-row entry): Iterator::min_by's first minimum over the whole N list
    (dual_simplex_solver.rs:270-279) -- order-free, so merging the ranks in any order gives the same answer."""
    best = None
    nan = False
    for (v, f, p, al) in msgs:
        if f != 0.0:
            nan = True
        p = int(p)
        if p >= 0 and (best is None or v < best[0] or (v == best[0] and p < best[1])):
            best = (v, p, al)
    if nan:
        return "nan", -1, 0.0, 0.0
    if best is None:
        return "infeasible", -1, 0.0, 0.0
    return "pick", best[1], best[0], best[2]


def nccl_library_path() -> str:
    try:
        import nvidia.nccl as pkg  # torch's bundled NCCL
        p = os.path.join(os.path.dirname(pkg.__file__), "lib", "libnccl.so.2")
        if os.path.exists(p):
            return p
    except Exception:
        pass
    return "libnccl.so.2"


def init_comm(ctx: N.Context, rank: int, world: int):
    """Creates the library's NCCL communicator; the 128-byte unique id travels over torch.distributed."""
    import torch
    import torch.distributed as dist
    path = nccl_library_path().encode()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        rc = N.lib.ellp_b200_comm_unique_id(path, C.cast(buf, C.c_void_p))
        if rc != N.OK:
            raise N.NativeError(rc, "ncclGetUniqueId failed")
    t = torch.tensor(list(buf), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    idbuf = (C.c_ubyte * 128).from_buffer_copy(raw)
    ctx.check(N.lib.ellp_b200_comm_init(ctx.h, path, C.cast(idbuf, C.c_void_p), rank, world))


def _emit(line: dict) -> None:
    """bench.py's emit(): the ONE JSON line goes to the descriptor bench.py saved before pointing fd 1 at stderr."""
    import sys
    sys.stdout.flush()
    os.write(int(os.environ.get("ELLP_BENCH_STDOUT_FD", "1")), (json.dumps(line) + "\n").encode())


def bench_main(args, wl, name, METRIC, UNIT, SEED, ClockSampler, measured_peak, cpu_reference_sample):
    """bench.py body for N > 1 (strong scaling: the same LP, columns split over the ranks)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = N.Context(local_rank)
    init_comm(ctx, rank, world)
    if getattr(args, "owner_ratio", -1) >= 0:
        ctx.set_tuning("owner_ratio", args.owner_ratio)
    m, ns, P = wl["m"], wl["ns"], (args.pivots or wl["pivots"])
    n = m + ns
    # the SAME block_k and pivots per step at every N (so that N = 1, 2, 4, 8 run the same pivots and end on the same objective);
    # a per-N tuned block length is reported separately under "tuned"
    bk = wl.get("block_k", 0) if getattr(args, "block_k", -1) < 0 else args.block_k
    dual = bool(wl.get("dual")) and bool(wl.get("tableau"))
    variant = 1 if dual else 0
    peer = bk > 1  # peer-memory engine: condensed tableau split by nonbasic position, exchange fused into the pivot kernel
    lo, hi = shard_range(ns, world, rank) if peer else shard_range(n, world, rank)
    # order-free tie rule: the arg-select is a reduction over ranks; the dual's rules are order-free as they are
    o = N.default_opts(P, engine=N.ENGINE_TABLEAU, tie_rule=N.TIES_CANONICAL, check_every=min(P, max(16, bk)), profile=True, block_k=bk)
    ctx.check(N.lib.ellp_b200_sharded_generate_dense_ex(ctx.h, m, ns, SEED, variant, C.byref(o)))

    def step():
        res = N.Result()
        ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
        assert res.status == N.MAXITER and res.iters == P, (res.status, res.iters)
        return res

    def barrier():
        ctx.check(N.lib.ellp_b200_sync(ctx.h))
        torch.cuda.synchronize()
        dist.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = ctx.launch_count()
    t0 = time.perf_counter()
    dev_ms = rank1_ms = 0.0
    n_rank1 = 0
    for _ in range(args.steps):
        r = step()
        dev_ms += r.ms_device; rank1_ms += r.ms_rank1; n_rank1 += r.n_rank1
    ctx.check(N.lib.ellp_b200_sync(ctx.h))
    torch.cuda.synchronize()
    dt_local = time.perf_counter() - t0
    launches = ctx.launch_count() - launches0
    clk = clocks.stop()
    tt = torch.tensor([dt_local, dev_ms, rank1_ms / max(n_rank1, 1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt, dev_ms_max, k3_ms = [float(v) for v in tt.cpu()]
    try:  # version of the row reduction that launch_rankk's auto rule picked for this shard shape (read before the parity LP below runs)
        flush_version = int(N.lib.ellp_b200_last_flush_kernel(ctx.h))
    except Exception:
        flush_version = 0
    value = args.steps * P / dt
    obj_resident = float(r.obj)

    # tuned extra (narrow shards: the pivot kernel's latency dominates and the longest block amortises the flush best)
    tuned = None
    if peer and world >= 4 and getattr(args, "block_k", -1) < 0 and bk != 64:
        Pt = (P + 63) // 64 * 64
        ot = N.default_opts(Pt, engine=N.ENGINE_TABLEAU, tie_rule=N.TIES_CANONICAL, check_every=64, block_k=64)
        ctx.check(N.lib.ellp_b200_sharded_generate_dense_ex(ctx.h, m, ns, SEED, variant, C.byref(ot)))
        rt = N.Result()
        for _ in range(3):
            ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(ot), C.byref(rt)))
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(ot), C.byref(rt)))
        ctx.check(N.lib.ellp_b200_sync(ctx.h))
        torch.cuda.synchronize()
        tq = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(tq, op=dist.ReduceOp.MAX)
        tuned = {"block_k": 64, "pivots_per_step": Pt, "value": args.steps * Pt / float(tq.cpu()[0]), "unit": UNIT}

    # parity check (every run with N > 1): 64 pivots of a 2048 x 6144 LP on the peer engine (all ranks) and on this rank's own
    # single-GPU blocked engine must give the same trace, x, B, N, objective -- bit for bit; a mismatch fails the bench (rc != 0)
    parity = None
    if peer:
        pm, pns, pK, pbk = 2048, 4096, 64, 32
        po = N.default_opts(pK, engine=N.ENGINE_TABLEAU, tie_rule=N.TIES_CANONICAL, check_every=32, block_k=pbk)

        def run_and_fetch(c_, gen):
            tr = np.zeros(pK, dtype=N.TRACE_DTYPE); po.trace = N.ptr(tr); po.trace_cap = pK
            c_.check(gen(c_.h, pm, pns, SEED + 5, variant, C.byref(po)))
            rr = N.Result()
            c_.check(N.lib.ellp_b200_run(c_.h, C.byref(po), C.byref(rr)))
            x = np.zeros(pm + pns); B = np.zeros(pm, dtype=np.int32); Nv = np.zeros(pns, dtype=np.int32); Ns = np.zeros(pns, dtype=np.uint8)
            y = np.zeros(pm); d = np.zeros(pm + pns)
            c_.check(N.lib.ellp_b200_download(c_.h, C.byref(N.Point(N.ptr(x), N.ptr(B), N.ptr(Nv), N.ptr(Ns), N.ptr(y) if dual else None,
                                                                  N.ptr(d) if dual else None, pm, pns))))
            return rr, tr, (x, B, Nv, Ns, y, d)

        single = N.Context(local_rank)
        rs, trs, ds = run_and_fetch(single, N.lib.ellp_b200_generate_dense_ex)
        single.close()
        rp, trp, dp = run_and_fetch(ctx, N.lib.ellp_b200_sharded_generate_dense_ex)
        same = (rp.status == rs.status and rp.iters == rs.iters == pK and rp.obj == rs.obj and trp.tobytes() == trs.tobytes()
                and all(a.tobytes() == b.tobytes() for a, b in zip(dp, ds)))
        flag = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = "ok" if int(flag.item()) == 1 else "MISMATCH"
        if parity != "ok":
            if rank == 0:
                _emit({"metric": METRIC, "parity_check": parity, "n_gpus": world, "error": "peer engine and single-GPU engine disagree"})
            ctx.close()
            dist.destroy_process_group()
            raise SystemExit(3)

    # e2e: every rank uploads ITS column block from pinned host memory, pivots, downloads the point
    e2e = None
    if not args.no_e2e:
        nloc = hi - lo
        A_h = torch.empty(m * nloc, dtype=torch.float64, pin_memory=True).numpy()
        c_h = np.zeros(n); b_h = np.zeros(m); lb_h = np.zeros(n); ub_h = np.zeros(n); kind_h = np.zeros(n, dtype=np.uint8)
        ctx.check(N.lib.ellp_b200_sharded_generate_dense_ex(ctx.h, m, ns, SEED, variant, C.byref(o)))
        if dual:  # the resident slice is the tableau -A_N: the host copy of A_N comes from the numpy twin of the generator
            import bench_lp
            idx = (np.arange(lo, hi, dtype=np.uint64)[:, None] * np.uint64(m) + np.arange(m, dtype=np.uint64)[None, :]).reshape(-1)
            A_h[:] = bench_lp._uniform01(SEED, idx)
            ctx.check(N.lib.ellp_b200_download_std_form(ctx.h, None, N.ptr(c_h), N.ptr(b_h), N.ptr(kind_h), N.ptr(lb_h), N.ptr(ub_h)))
        else:
            ctx.check(N.lib.ellp_b200_download_std_form(ctx.h, N.ptr(A_h), N.ptr(c_h), N.ptr(b_h), N.ptr(kind_h), N.ptr(lb_h), N.ptr(ub_h)))
        x0 = np.zeros(n); x0[ns:] = -b_h if dual else b_h
        y0 = np.zeros(m); d0 = c_h.copy(); diag = -np.ones(m)
        B0 = np.arange(ns, n, dtype=np.int32); N0 = np.arange(ns, dtype=np.int32); Ns0 = np.zeros(ns, dtype=np.uint8)
        sf = N.StdForm(m, n, N.ptr(A_h), N.ptr(c_h), N.ptr(b_h), N.ptr(kind_h), N.ptr(lb_h), N.ptr(ub_h))
        oe = N.default_opts(P, engine=N.ENGINE_TABLEAU, tie_rule=N.TIES_CANONICAL, check_every=min(P, max(16, bk)), block_k=bk)
        upload = N.lib.ellp_b200_sharded_upload_nonbasic if peer else N.lib.ellp_b200_sharded_upload

        def e2e_step():
            x, B, Nv, Ns = x0.copy(), B0.copy(), N0.copy(), Ns0.copy()
            y, d = y0.copy(), d0.copy()
            pt = N.Point(N.ptr(x), N.ptr(B), N.ptr(Nv), N.ptr(Ns), N.ptr(y) if dual else None, N.ptr(d) if dual else None, m, ns)
            if dual:
                ctx.check(N.lib.ellp_b200_sharded_upload_nonbasic_ex(ctx.h, C.byref(sf), C.byref(pt), N.DUAL, N.ptr(diag), C.byref(oe)))
            else:
                ctx.check(upload(ctx.h, C.byref(sf), C.byref(pt), C.byref(oe)))
            res = N.Result()
            ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(oe), C.byref(res)))
            ctx.check(N.lib.ellp_b200_download(ctx.h, C.byref(pt)))
            assert res.status == N.MAXITER and res.iters == P
            return float(res.obj) if dual else float(np.dot(c_h, x))

        for _ in range(min(args.warmup, 3)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            obj = e2e_step()
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dte = float(te.cpu()[0])
        h2d = 8 * m * nloc + 8 * (3 * n + m) + n + 8 * n + 4 * m + n
        d2h = 8 * n + 4 * m + n
        e2e = {"value": args.steps * P / dte, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "ms_per_step": 1e3 * dte / args.steps, "api": "ellp_b200_sharded_upload + ellp_b200_run + ellp_b200_download (host buffers)",
               "objective_after_step": obj}

    if rank == 0:
        peak, peak_src = measured_peak()
        nloc = hi - lo
        if peer:
            alg_bytes = 16.0 * m * nloc + 8.0 * bk * (m + nloc)
            fk = {1: "k_blk_flush", 3: "k_blk_flush3", 4: "k_blk_flush4", 5: "k_blk_flush5<2>", 6: "k_blk_flush5<4>", 7: "k_blk_flush6<1>",
                  8: "k_blk_flush6<2>", 9: "k_blk_flush4r<3>"}.get(flush_version, "k_blk_flush")
            kernel = "%s (rank-%d row reduction of the local %d x %d slice of the condensed tableau, fp64 DMMA)" % (fk, bk, m, nloc)
        else:
            alg_bytes = 16.0 * m * nloc + 8.0 * (m + nloc)
            kernel = "k_rank1<true> on the local column shard"
        achieved = alg_bytes / (k3_ms * 1e-3) / 1e9 if k3_ms > 0 else None
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src, "ms_per_launch": k3_ms,
                    "algorithmic_bytes_per_launch": alg_bytes, "note": "per GPU; max over ranks of the mean launch time"}
        roofline["algorithmic_bytes_per_launch"] = alg_bytes
        if peer and achieved:
            import bench as _bench
            roofline = _bench.blocked_roofline(roofline, m, nloc, bk, k3_ms)
        engine = ("dual simplex, " if dual else "") + ("condensed tableau split by nonbasic position, blocked (block_k=%d), per-pivot exchange fused into the cooperative "
                  "pivot kernel over NVLink peer memory (k_blk_pivots_peer)" % bk) if peer else "tableau, column-sharded, rank-1 update per pivot"
        exchange = ("per pivot: %d x 64 B pricing words + m x 16 B column words stored into every peer (LL protocol, no NCCL call)" % world) if peer \
            else "2 x ncclAllGather (8 B, 24 B per rank) + 1 x ncclAllReduce (m doubles) per pivot"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": name, "m": m, "n": n, "pivots_per_step": P, "engine": engine, "block_k": bk,
                           "columns_per_gpu": nloc, "tie_rule": "order-free (canonical)", "owner_ratio": getattr(args, "owner_ratio", -1),
                           "exchange": exchange,
                           "l2": f"local shard {8.0 * m * nloc / 1e9:.2f} GB >> 126 MB L2"},
                "device_ms_per_step": dev_ms_max / args.steps, "gpu_launches": int(launches) * world, "clocks": clk,
                "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "parity_check": parity,
                "parity_check_what": "64 pivots of a 2048 x 6144 LP: peer engine on all ranks vs every rank's single-GPU blocked engine, bit-identical trace / x / B / N / objective" if parity else None,
                "objective_after_timed_steps": obj_resident, "tuned": tuned}
        _emit(line)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
