"""What bounds the rank-k row reduction at 0.75 of the DGEMM peak?  (1) sustained back-to-back launches with clocks / power sampled;
(2) the same kernels on L2-resident tableaus (no HBM round trip per tile): if the fp64 tensor rate rises there, the limiter is the
HBM side (latency / power), not the DMMA issue structure of the kernel."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench as BM
from ellp_b200 import _native as N
import blk_sweep as B
ctx = N.Context(0)
ctx.set_tuning("flush_waves", 0)
for R, Cc, cs in ((4096, 1024, 2), (2048, 2048, 2), (4736, 2048, 4), (18944, 4096, 32)):
    for fk in (4, 5, 3):
        for k in (56, 40):
            ctx.set_tuning("flush_kernel", fk)
            d = B.flush_point(ctx, R, Cc, k, cs, reps=50, warm=5)
            ctas = ((R + 127) // 128) * (((Cc + (127 if fk >= 4 else 63)) // (128 if fk >= 4 else 64) + cs - 1) // cs)
            d.update(flush_kernel=fk, ctas=ctas, MB=R * Cc * 8 / 1e6, TFLOPs_per_busy_SM_x148=round(d["TFLOPs"] * 148 / min(ctas, 148), 2) if ctas <= 148 else None)
            print(json.dumps(d), flush=True)
ctx.set_tuning("flush_waves", 6)
for fk, k in ((4, 56), (5, 64)):
    ctx.set_tuning("flush_kernel", fk)
    B.flush_point(ctx, 32768, 32768, k, 32, reps=3)
    s = BM.ClockSampler(0); s.start(); time.sleep(0.2)
    d = B.flush_point(ctx, 32768, 32768, k, 32, reps=250, warm=2)
    clk = s.stop()
    d.update(flush_kernel=fk, clocks=clk)
    print(json.dumps(d), flush=True)
ctx.set_tuning("flush_kernel", 0)
