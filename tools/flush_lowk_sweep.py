"""Round-2 sweeps that fix the auto rule of launch_rankk: (1) shard shapes 32768 x {4096, 8192, 16384} at k = 56 / 64 with versions 3 / 8 /
9 (flush_kernel); (2) below the tensor-bound band (k = 16 .. 40) on the full tableau and on the 4096 x 8192 tableau of BASELINE.json
configs[2]; (3) column steps per CTA of version 9 on the narrowest shard.  One JSON line per point."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ellp_b200 import _native as N
import blk_sweep

ctx = N.Context(0)
def pt(R, Cc, k, kern, cs=32, **kw):
    ctx.set_tuning("flush_kernel", kern)
    d = blk_sweep.flush_point(ctx, R, Cc, k, cs, reps=8, warm=3)
    d["flush_kernel"] = kern
    d.update(kw)
    print(json.dumps(d), flush=True)
for Cc in (4096, 8192, 16384):
    for k in (56, 64):
        for kern in (3, 8, 9):
            pt(32768, Cc, k, kern)
for (R, Cc) in ((32768, 32768), (4096, 8192)):
    for k in (16, 24, 32, 40, 48):
        for kern in (3, 8, 9):
            pt(R, Cc, k, kern)
ctx.set_tuning("flush_waves", 0)
for cs in (4, 8, 16, 32):
    for k in (56, 64):
        pt(32768, 4096, k, 9, cs, heuristic="off")
ctx.set_tuning("flush_waves", 6)
ctx.set_tuning("flush_kernel", 0)
ctx.close()
