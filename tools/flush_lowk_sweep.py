"""k_blk_flush3 (8 warps, two register tiles) against k_blk_flush6<2> (16 warps, pipelined halves, no producer warp) below the tensor-bound
band (k = 16 .. 40) on the full tableau and on the 4096 x 8192 tableau of BASELINE.json configs[2].  One JSON line per point."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ellp_b200 import _native as N
import blk_sweep

ctx = N.Context(0)
for (R, Cc) in ((32768, 32768), (4096, 8192), (32768, 4096)):
    for k in (16, 24, 32, 40):
        for kern in (3, 8):
            ctx.set_tuning("flush_kernel", kern)
            d = blk_sweep.flush_point(ctx, R, Cc, k, 32, reps=8, warm=3)
            d["flush_kernel"] = kern
            print(json.dumps(d), flush=True)
ctx.set_tuning("flush_kernel", 0)
ctx.close()
