"""Round-2 sweep of the rank-k row reduction (K3b): k_blk_flush4 (one register tile per warp, loaded and waited for as a whole) against
k_blk_flush5<2> / <4> (tuning key flush_kernel = 5 / 6: the same tile pipelined in 2 / 4 parts inside the warp) on the full 32768 x 32768
condensed tableau and on the shard shapes, then the whole pivot loop of BASELINE.json configs[4] with each kernel.  One JSON line per point."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ellp_b200 import _native as N
import blk_sweep

ctx = N.Context(0)
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
for k in (56, 64, 48) if not quick else (56,):
    for kern in (4, 5, 6):
        for cs in (32, 16) if not quick else (32,):
            ctx.set_tuning("flush_kernel", kern)
            d = blk_sweep.flush_point(ctx, 32768, 32768, k, cs, reps=8, warm=3)
            d["flush_kernel"] = kern
            print(json.dumps(d), flush=True)
for Cc in (4096, 8192):
    for kern in (4, 5, 6):
        ctx.set_tuning("flush_kernel", kern)
        d = blk_sweep.flush_point(ctx, 32768, Cc, 56, 32, reps=8, warm=3)
        d["flush_kernel"] = kern
        print(json.dumps(d), flush=True)
for bk in (56, 64):
    for kern in (4, 5, 6):
        ctx.set_tuning("flush_kernel", kern)
        d = blk_sweep.loop_point(ctx, 32768, 32768, bk, 12 * bk, 32)
        d["flush_kernel"] = kern
        print(json.dumps(d), flush=True)
ctx.set_tuning("flush_kernel", 0)
ctx.close()
