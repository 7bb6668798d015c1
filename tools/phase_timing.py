"""Phase timing of k_blk_pivots_fused on cuda:0 (block 0 / thread 0 clock64 stamps): mean microseconds per phase."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ellp_b200 import _native as N  # noqa: E402

NAMES = ["A+ticket(send)", "B(mailbox wait+merge)", "C1(owner column)", "C2(recv col+ratios+blockred)", "grid.sync", "D(decide)", "E+A'(row,price)", "blockred+partA"]


def run(ctx, m, ns, bk, pivots, mhz=1965.0):
    o = N.default_opts(pivots, engine=N.ENGINE_TABLEAU, block_k=bk, check_every=pivots)
    ctx.check(N.lib.ellp_b200_generate_dense(ctx.h, m, ns, 0, C.byref(o)))
    res = N.Result()
    ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))  # warm-up
    ctx.set_tuning("phase_timing", pivots)
    ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
    log = np.zeros((pivots, 10), dtype=np.int64)
    ctx.check(N.lib.ellp_b200_phase_log(ctx.h, log.ctypes.data_as(C.POINTER(C.c_int64)), pivots))
    ctx.set_tuning("phase_timing", 0)
    d = np.diff(log[:, :9], axis=1) / mhz  # us
    ok = (log[:, 8] > 0)
    mean = d[ok].mean(axis=0)
    out = {"m": m, "n": m + ns, "block_k": bk, "pivots": int(ok.sum()), "us_per_pivot_kernel": round(float(mean.sum()), 2),
           "ms_device_per_pivot_incl_flush": round(res.ms_device / res.iters * 1e3, 2)}
    for nme, v in zip(NAMES, mean):
        out[nme] = round(float(v), 2)
    return out


if __name__ == "__main__":
    ctx = N.Context(0)
    for thr, cps in ((512, 1), (256, 1), (128, 1)):
        ctx.set_tuning("coop_threads", thr)
        ctx.set_tuning("coop_ctas_per_sm", cps)
        for (m, ns) in ((1024, 2048), (4096, 8192), (32768, 32768)):
            d = run(ctx, m, ns, 32, 256)
            d["coop_threads"] = thr; d["ctas_per_sm"] = cps
            print(json.dumps(d), flush=True)
