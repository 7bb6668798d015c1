"""Clocks / power while the rank-k flush kernels run back to back (is the fp64 tensor pipe power-limited?)."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench as BM
from ellp_b200 import _native as N
import blk_sweep as B
ctx = N.Context(0)
for fk, k, cs in ((3, 64, 32), (4, 56, 32), (4, 48, 32), (1, 24, 8)):
    ctx.set_tuning("flush_kernel", fk)
    B.flush_point(ctx, 32768, 32768, k, cs, reps=3)
    s = BM.ClockSampler(0); s.Q = s.Q; s.start(); time.sleep(0.3)
    d = B.flush_point(ctx, 32768, 32768, k, cs, reps=300, warm=2)
    clk = s.stop()
    d.update(flush_kernel=fk, clocks=clk)
    print(json.dumps(d), flush=True)
