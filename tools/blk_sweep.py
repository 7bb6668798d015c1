"""Sweeps the blocked tableau engine on cuda:0:
  (1) the rank-k flush kernel alone (K3b, k_blk_flush) over k and columns-per-CTA: ms, algorithmic GB/s, DMMA TFLOP/s;
  (2) the whole pivot loop (pivots/s) over block_k, same LP, same pivot budget.
Usage: python tools/blk_sweep.py [m ns] ...   Output: one JSON line per point."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ellp_b200 import _native as N  # noqa: E402

PEAK = 6541.1
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def flush_point(ctx, R, Cc, k, col_steps, reps=5, warm=2):
    dE, dU, dV = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx.check(N.lib.ellp_b200_dev_alloc(ctx.h, R * Cc * 8, C.byref(dE)))
    ctx.check(N.lib.ellp_b200_dev_alloc(ctx.h, R * 64 * 8, C.byref(dU)))
    ctx.check(N.lib.ellp_b200_dev_alloc(ctx.h, Cc * 64 * 8, C.byref(dV)))
    try:
        ctx.check(N.lib.ellp_b200_dev_fill_uniform(ctx.h, dE, R * Cc, 1, 0, 0.0, 1.0))
        ctx.check(N.lib.ellp_b200_dev_fill_uniform(ctx.h, dU, R * 64, 2, 0, -1e-3, 1e-3))
        ctx.check(N.lib.ellp_b200_dev_fill_uniform(ctx.h, dV, Cc * 64, 3, 0, -1e-3, 1e-3))
        ctx.set_tuning("flush_col_steps", col_steps)
        ms = C.c_float()
        ctx.check(N.lib.ellp_b200_rankk_update_dev(ctx.h, dE, R, Cc, R, dU, dV, Cc, k, warm, C.byref(ms)))
        ctx.check(N.lib.ellp_b200_rankk_update_dev(ctx.h, dE, R, Cc, R, dU, dV, Cc, k, reps, C.byref(ms)))
        bytes_alg = 16.0 * R * Cc + 8.0 * k * (R + Cc)
        gbs = bytes_alg / (ms.value * 1e-3) / 1e9
        return dict(kind="flush", R=R, C=Cc, k=k, col_steps=col_steps, ms=round(ms.value, 4), GBs=round(gbs, 1), frac_measured_peak=round(gbs / PEAK, 4),
                    TFLOPs=round(2.0 * R * Cc * k / (ms.value * 1e-3) / 1e12, 2), ms_per_pivot=round(ms.value / k, 4))
    finally:
        for d in (dE, dU, dV):
            N.lib.ellp_b200_dev_free(ctx.h, d)


def loop_point(ctx, m, ns, bk, pivots, col_steps=8):
    o = N.default_opts(pivots, engine=N.ENGINE_TABLEAU, block_k=bk, check_every=min(pivots, 32), profile=True)
    ctx.set_tuning("flush_col_steps", col_steps)
    ctx.check(N.lib.ellp_b200_generate_dense(ctx.h, m, ns, 0, C.byref(o)))
    res = N.Result()
    ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))  # warm-up (also builds the reduced-cost row)
    tot_ms = 0.0; tot_piv = 0; k_ms = 0.0; k_n = 0
    for _ in range(3):
        ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
        assert res.status == N.MAXITER and res.iters == pivots, (res.status, res.iters)
        tot_ms += res.ms_device; tot_piv += res.iters; k_ms += res.ms_rank1; k_n += res.n_rank1
    return dict(kind="loop", m=m, n=m + ns, block_k=bk, pivots_per_run=pivots, pivots_per_s=round(tot_piv / (tot_ms * 1e-3), 1),
                ms_per_pivot=round(tot_ms / tot_piv, 4), row_reduction_ms=round(k_ms / max(k_n, 1), 4),
                row_reduction_share=round(k_ms / tot_ms, 3), obj=res.obj, launches=int(res.launches))


if __name__ == "__main__":
    ctx = N.Context(0)
    sizes = [(16384, 16384)]
    if len(sys.argv) >= 3:
        a = list(map(int, sys.argv[1:]))
        sizes = list(zip(a[0::2], a[1::2]))
    for m, ns in sizes:
        n = m + ns
        for k in (8, 16, 32, 40, 48, 64):
            for cs in ((16,) if k not in (32, 48) else (4, 8, 16, 32, 64)):
                print(json.dumps(flush_point(ctx, m, n, k, cs)), flush=True)
        for bk in (16, 32, 48, 64):
            print(json.dumps(loop_point(ctx, m, ns, bk, 192 if bk else 40)), flush=True)
