"""Round-2 sweep of the rank-k row reduction (K3b), tuning key flush_kernel: 4 = k_blk_flush4 (16 consumer warps + producer warp, one
register tile per warp), 5 = k_blk_flush5<2> (tile pipelined in two halves inside the warp), 7 / 8 = k_blk_flush6<1 / 2> (no producer
warp: 512 threads, 128 registers, no spills; schedule of version 4 / 5) on the full 32768 x 32768 condensed tableau and on the shard
shapes, then the whole pivot loop of BASELINE.json configs[4] with each kernel.  One JSON line per point."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ellp_b200 import _native as N
import blk_sweep

ctx = N.Context(0)
kerns = (4, 5, 8, 9)
for k in (56, 64, 48, 40):
    for kern in kerns:
        ctx.set_tuning("flush_kernel", kern)
        d = blk_sweep.flush_point(ctx, 32768, 32768, k, 32, reps=8, warm=3)
        d["flush_kernel"] = kern
        print(json.dumps(d), flush=True)
for Cc in (4096, 8192):
    for kern in kerns:
        ctx.set_tuning("flush_kernel", kern)
        d = blk_sweep.flush_point(ctx, 32768, Cc, 56, 32, reps=8, warm=3)
        d["flush_kernel"] = kern
        print(json.dumps(d), flush=True)
for bk in (56, 64):
    for kern in (4, 8, 9):
        ctx.set_tuning("flush_kernel", kern)
        d = blk_sweep.loop_point(ctx, 32768, 32768, bk, 12 * bk, 32)
        d["flush_kernel"] = kern
        print(json.dumps(d), flush=True)
ctx.set_tuning("flush_kernel", 0)
ctx.close()
