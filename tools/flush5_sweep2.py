"""Second sweep of the 16-warp row-reduction kernels: tile access mode (tuning flush_ld: 1 = evict-first .cs, 2 = .cg loads that bypass
L1 + .cs stores, 3 = .cg both, 0 = default caching) and ring depth (flush_stages), k_blk_flush4 / k_blk_flush5<2>, full 32768^2 tableau."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ellp_b200 import _native as N
import blk_sweep

ctx = N.Context(0)
pts = []
for k in (56, 64):
    for kern in (4, 5):
        for mode in (1, 2, 3, 0):
            pts.append((k, kern, mode, 0))
pts += [(48, 4, 1, 3), (48, 4, 2, 3), (48, 5, 2, 3), (56, 4, 1, 1), (56, 4, 2, 1), (40, 4, 2, 0), (40, 4, 1, 0)]
for k, kern, mode, st in pts:
    ctx.set_tuning("flush_kernel", kern)
    ctx.set_tuning("flush_ld", mode)
    ctx.set_tuning("flush_stages", st)
    d = blk_sweep.flush_point(ctx, 32768, 32768, k, 32, reps=8, warm=3)
    d.update(flush_kernel=kern, flush_ld=mode, flush_stages=st)
    print(json.dumps(d), flush=True)
ctx.set_tuning("flush_kernel", 0); ctx.set_tuning("flush_ld", -1); ctx.set_tuning("flush_stages", 0)
ctx.close()
