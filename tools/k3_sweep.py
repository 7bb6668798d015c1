"""Times the rank-1 row-reduction kernel (K3) alone at several sizes / tunings on cuda:0 (CUDA events inside
the library, on its own stream).  Usage: python tools/k3_sweep.py [R C] ...  Output: one JSON line per point."""
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


def run(ctx, R, Cc, cpc, stream_mb, reps=20, warm=3):
    dE, da = C.c_void_p(), C.c_void_p()
    ctx.check(N.lib.ellp_b200_dev_alloc(ctx.h, R * Cc * 8, C.byref(dE)))
    ctx.check(N.lib.ellp_b200_dev_alloc(ctx.h, R * 8, C.byref(da)))
    try:
        ctx.check(N.lib.ellp_b200_dev_fill_uniform(ctx.h, dE, R * Cc, 1, 0, 0.0, 1.0))
        ctx.check(N.lib.ellp_b200_dev_fill_uniform(ctx.h, da, R, 2, 0, -1e-3, 1e-3))
        one = (C.c_double * 1)(1.0)
        r = R // 3
        ctx.check(N.lib.ellp_b200_h2d(ctx.h, C.c_void_p(da.value + r * 8), C.cast(one, C.c_void_p), 8))
        ctx.set_tuning("rank1_cols_per_cta", cpc)
        ctx.set_tuning("rank1_stream_min_mb", stream_mb)
        ms = C.c_float()
        ctx.check(N.lib.ellp_b200_rank1_update_dev(ctx.h, dE, R, Cc, R, da, r, warm, C.byref(ms)))
        ctx.check(N.lib.ellp_b200_rank1_update_dev(ctx.h, dE, R, Cc, R, da, r, reps, C.byref(ms)))
        bytes_alg = 16.0 * R * Cc + 8.0 * (R + Cc)
        gbs = bytes_alg / (ms.value * 1e-3) / 1e9
        return dict(R=R, C=Cc, cols_per_cta=cpc, stream=stream_mb == 0, ms=round(ms.value, 4), GBs=round(gbs, 1),
                    frac_measured_peak=round(gbs / PEAK, 4), frac_8TBs=round(gbs / 8000.0, 4))
    finally:
        N.lib.ellp_b200_dev_free(ctx.h, dE)
        N.lib.ellp_b200_dev_free(ctx.h, da)


if __name__ == "__main__":
    ctx = N.Context(0)
    sizes = [(16384, 32768)]
    if len(sys.argv) >= 3:
        a = list(map(int, sys.argv[1:]))
        sizes = list(zip(a[0::2], a[1::2]))
    for R, Cc in sizes:
        for stream_mb in (0, 1 << 30):
            for cpc in (8, 16, 32, 64, 128, 256, 1024):
                print(json.dumps(run(ctx, R, Cc, cpc, stream_mb)), flush=True)
