"""Sweep of the rank-k row reduction (K3b) on the SHARD shapes of the sharded 32768 x 65536 LP (columns per GPU at N = 2, 4, 8): column
steps per CTA with the wave heuristic of launch_rankk switched off, k = 56 / 64.  One JSON line per point."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ellp_b200 import _native as N
import blk_sweep

ctx = N.Context(0)
for Cc in (4096, 8192, 16384):
    for k in (56, 64):
        ctx.set_tuning("flush_waves", 6)
        d = blk_sweep.flush_point(ctx, 32768, Cc, k, 32)
        d["heuristic"] = "default (6 waves)"
        print(json.dumps(d), flush=True)
        ctx.set_tuning("flush_waves", 0)
        for cs in (2, 4, 8, 16, 32):
            d = blk_sweep.flush_point(ctx, 32768, Cc, k, cs)
            d["heuristic"] = "off"
            print(json.dumps(d), flush=True)
ctx.set_tuning("flush_waves", 6)
ctx.close()
