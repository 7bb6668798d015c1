#!/usr/bin/env python
"""bench.py -- simplex pivots/s of the B200 pivot loop on the dense LP of BASELINE.json configs[4].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--pivots P] [--block-k k]

A "step" = one pass of the hot path over one batch of synthetic input = P simplex pivots of the same dense LP
(`min -c.x, Ax <= b, x >= 0`, A ~ U(0,1); standard form m x (n_struct + m), slack starting basis; built in HBM by a
counter-based generator).  `value` times steps with the tableau resident in HBM (each step continues pivoting where
the previous one stopped); `e2e` times the reference-facing C-ABI call ellp_b200_primal_solve_with_initial on HOST
buffers (H2D of the standard form -- only its nonbasic columns when the starting basis is verified to be the identity --
+ P pivots + D2H of the point, every step).  N > 1: the condensed tableau is split by nonbasic position, one rank per GPU
(torchrun), with the per-pivot exchange fused into the pivot kernel over NVLink peer memory (ellp_b200/sharded.py,
ellp_b200/csrc/peer.cuh).

Other workloads (--workload): the north_star target size 16384x32768, 4096x12288, the dual simplex on the revised engine
(configs[2] shape, reference rules and steepest edge + Harris), the batch of 65536 small LPs (configs[3]) and the netlib
LPs through the public API (configs[0] / configs[1]).

`--impl reference` times the reference's own CPU algorithm (the oracle port of ellp's PrimalSimplexSolver::
solve_with_initial: a fresh dense LU per pivot, single thread like the reference) on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: keep NCCL's "NCCL version ..." banner (printed to stdout at NCCL_DEBUG=VERSION) out of it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

# Anything a native library prints to fd 1 (NCCL's "NCCL version ..." banner, a stray printf) must not reach the driver: main()
# points fd 1 at stderr for the whole run and emit() writes the ONE JSON line to the saved descriptor.
def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    fd = int(os.environ.get("ELLP_BENCH_STDOUT_FD", "1"))
    sys.stdout.flush()
    os.write(fd, data)


def _protect_stdout() -> None:
    if "ELLP_BENCH_STDOUT_FD" in os.environ:
        return
    sys.stdout.flush()
    saved = os.dup(1)
    os.set_inheritable(saved, False)
    os.dup2(2, 1)
    os.environ["ELLP_BENCH_STDOUT_FD"] = str(saved)


WORKLOADS = {
    # BASELINE.json configs[4]: "synthetic dense LP 32768x65536 fp64, tableau column-sharded ... at 1/2/4/8 B200"
    "dense_tableau_32768x65536": dict(m=32768, ns=32768, pivots=704, block_k=64, sample_m=1024, sample_pivots=3),
    # north_star target size: "for a 16384x32768 dense LP, the row-reduction kernel sustains >= 70% of HBM bandwidth"
    "dense_tableau_16384x32768": dict(m=16384, ns=16384, pivots=704, block_k=64, sample_m=1024, sample_pivots=3),
    "dense_tableau_4096x12288": dict(m=4096, ns=8192, pivots=960, block_k=48, sample_m=512, sample_pivots=6),
    # BASELINE.json configs[2] shape: dense 4096x8192 (Gte rows => standard form 4096x12288), DUAL simplex, revised engine
    # (explicit basis inverse), the reference's entering / leaving rules.
    "dense_revised_dual_4096x12288": dict(m=4096, ns=8192, pivots=200, sample_m=512, sample_pivots=6, dual=True),
    # the same LP and rules on the blocked condensed tableau (dual_blocked.cuh): one cooperative launch per block of pivots
    "dense_tableau_dual_4096x12288": dict(m=4096, ns=8192, pivots=960, block_k=48, sample_m=512, sample_pivots=6, dual=True, tableau=True),
    # ... and with dual Devex pricing (reference weights updated from the entering column, no extra pass): a step = one whole solve
    "dense_tableau_dual_devex_4096x12288": dict(m=4096, ns=8192, pivots=200, block_k=48, sample_m=512, sample_pivots=6, dual=True, tableau=True, devex=True),
    "dense_tableau_dual_16384x32768": dict(m=16384, ns=16384, pivots=704, block_k=64, sample_m=1024, sample_pivots=3, dual=True, tableau=True),
    "dense_tableau_dual_32768x65536": dict(m=32768, ns=32768, pivots=704, block_k=64, sample_m=1024, sample_pivots=3, dual=True, tableau=True),
    # the same LP with dual steepest edge (exact weights from the rank-1 update's epilogue) + Harris ratio test
    "dense_revised_dual_dse_4096x12288": dict(m=4096, ns=8192, pivots=200, sample_m=512, sample_pivots=6, dual=True, dse=True),
    # BASELINE.json configs[0] / configs[1]: netlib LPs of the reference's tests/benchmark_problems (fixtures in tests/golden/),
    # PrimalSimplexSolver::solve and DualSimplexSolver::solve through the public API, wall time per solve vs the CPU port
    "netlib_afiro": dict(netlib="afiro"),
    "netlib_adlittle": dict(netlib="adlittle"),
    "netlib_blend": dict(netlib="blend"),
    "dense_tableau_tiny": dict(m=256, ns=256, pivots=32, block_k=8, sample_m=128, sample_pivots=4),
    # BASELINE.json configs[3]: "batch of 65536 independent small LPs (64x128), sharded one shard per GPU at 1/2/4/8 B200"
    "batch_small_lps_65536x64x128": dict(batch=True, nlp=65536, m=64, ns=128, sample_lps=150),
    "batch_small_lps_tiny": dict(batch=True, nlp=600, m=64, ns=128, sample_lps=20),
}
DEFAULT_WORKLOAD = "dense_tableau_32768x65536"
SEED = 0
METRIC = "simplex pivots/sec (fp64)"
UNIT = "pivots/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# fp64 tensor pipe (DMMA m8n8k4) peak of one B200: MEASURED_PEAKS.json carries no fp64 figure, so the denominator is derived
# from ncu: sm__pipe_tensor_subpipe_dmma_cycles_active = 62.1 % at 22.9 TFLOP/s (profiles/r01_ncu_full_blk_flush_k32_summary.txt)
# => 36.9 TFLOP/s = 148 SMs x 64 fp64 FMA/clk x 1.965 GHz.
DMMA_NOMINAL_TFLOPS = 36.9


def _fp64_peak():
    """Denominator of the fp64 tensor roofline: cuBLAS DGEMM 8192^3 MEASURED on this pool's B200 (tools/fp64_peaks.py ->
    profiles/r02_fp64_peaks.json, burst = best of 10 for a kernel timed alone); the 148 x 64 x 1.965 GHz figure stays as nominal."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_fp64_peaks.json")) as f:
            d = json.load(f)["dgemm_8192"]
        return float(d["tflops_burst"]), ("measured cuBLAS DGEMM 8192^3 on this pool's B200 (profiles/r02_fp64_peaks.json: burst %.2f, sustained %.2f TFLOP/s); "
                                          "nominal 148 SMs x 64 fp64 FMA/clk x 1.965 GHz = %.1f" % (d["tflops_burst"], d["tflops_sustained"], DMMA_NOMINAL_TFLOPS))
    except Exception:
        return DMMA_NOMINAL_TFLOPS, "nominal 148 SMs x 64 fp64 FMA/clk x 1.965 GHz (profiles/r02_fp64_peaks.json missing)"


DMMA_PEAK_TFLOPS, DMMA_PEAK_SOURCE = _fp64_peak()


def blocked_roofline(roofline: dict, m: int, cols: int, bk: int, ms: float) -> dict:
    """The rank-k flush moves 16 m cols bytes and does 2 m cols k flops: HBM-bound below k ~ 45, fp64-tensor-bound above.
    Reports the fraction of BOTH ceilings and names the binding one in `bound` / `frac`."""
    tfl = 2.0 * m * cols * bk / (ms * 1e-3) / 1e12
    hbm = {k: roofline[k] for k in ("achieved", "peak", "frac", "peak_source")}
    t_hbm = roofline["algorithmic_bytes_per_launch"] / (roofline["peak"] * 1e9)
    t_dmma = 2.0 * m * cols * bk / (DMMA_PEAK_TFLOPS * 1e12)
    roofline["hbm"] = hbm
    roofline["tensor_fp64"] = {"achieved": tfl, "peak": DMMA_PEAK_TFLOPS, "frac": tfl / DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "peak_source": DMMA_PEAK_SOURCE,
                               "nominal": DMMA_NOMINAL_TFLOPS, "frac_of_nominal": tfl / DMMA_NOMINAL_TFLOPS}
    roofline["pivots_per_launch"] = bk
    roofline["roofline_ms"] = 1e3 * max(t_hbm, t_dmma)
    roofline["frac_of_binding_roofline"] = 1e3 * max(t_hbm, t_dmma) / ms
    if t_dmma > t_hbm:
        roofline.update(bound="tensor", achieved=tfl, peak=DMMA_PEAK_TFLOPS, unit="TFLOP/s", frac=tfl / DMMA_PEAK_TFLOPS, peak_source=DMMA_PEAK_SOURCE)
        roofline.pop("frac_of_8TBs_nominal", None)
    return roofline


def k3_target_leg(ctx, N, peak, peak_src, R=16384, Cc=32768, reps=20):
    """north_star's kernel target on a driver-visible record: K3 (k_rank1, the rank-1 row reduction that replaces the reference's
    per-pivot LU) on a 16384 x 32768 fp64 matrix that never leaves HBM; algorithmic bytes 16 R C + 8 (R + C) per launch, timed
    with CUDA events on the library's stream (ellp_b200_rank1_update_dev)."""
    E = C.c_void_p(); al = C.c_void_p()
    ctx.check(N.lib.ellp_b200_dev_alloc(ctx.h, 8 * R * Cc, C.byref(E)))
    ctx.check(N.lib.ellp_b200_dev_alloc(ctx.h, 8 * R, C.byref(al)))
    try:
        ctx.check(N.lib.ellp_b200_dev_fill_uniform(ctx.h, E, R * Cc, 11, 0, 0.0, 1.0))
        ctx.check(N.lib.ellp_b200_dev_fill_uniform(ctx.h, al, R, 12, 0, 0.5, 1.5))
        ms = C.c_float(0)
        ctx.check(N.lib.ellp_b200_rank1_update_dev(ctx.h, E, R, Cc, R, al, R // 3, 3, C.byref(ms)))      # warm-up
        ctx.check(N.lib.ellp_b200_rank1_update_dev(ctx.h, E, R, Cc, R, al, R // 3, reps, C.byref(ms)))
    finally:
        ctx.check(N.lib.ellp_b200_dev_free(ctx.h, E))
        ctx.check(N.lib.ellp_b200_dev_free(ctx.h, al))
    alg = 16.0 * R * Cc + 8.0 * (R + Cc)
    gbs = alg / (ms.value * 1e-3) / 1e9
    return {"kernel": "k_rank1<STREAM> (K3) on a %d x %d fp64 matrix (8.59 GB >> L2)" % (R, Cc), "ms": float(ms.value), "launches_timed": reps,
            "algorithmic_bytes_per_launch": alg, "GB/s": gbs, "peak": peak, "frac": gbs / peak, "frac_of_8TBs_nominal": gbs / 8000.0,
            "peak_source": peak_src, "target": ">= 0.70 of HBM bandwidth (BASELINE.json north_star)", "met": gbs / peak >= 0.70}


def other_config_legs(ctx, N) -> dict:
    """Short legs of the OTHER BASELINE.json configs on the default (driver-run) bench line, so that one record carries all of them:
    configs[2] (dense 4096x8192 dual: reference rules, pivots/s; Devex, whole solve), configs[3] (batch of 65536 small LPs, one GPU)
    and configs[0] (netlib AFIRO through the public API).  Each leg is a few timed repetitions after a warm-up; failures are reported
    in place and never lose the main line."""
    out = {}
    try:  # configs[2]: dual simplex on the blocked tableau
        m, ns, P, bk = 4096, 8192, 960, 48
        o = N.default_opts(P, engine=N.ENGINE_TABLEAU, check_every=bk, block_k=bk)
        ctx.check(N.lib.ellp_b200_generate_dense_ex(ctx.h, m, ns, SEED, 1, C.byref(o)))
        res = N.Result()
        for _ in range(2):
            ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
        t0 = time.perf_counter(); piv = 0
        for _ in range(5):
            ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res))); piv += res.iters
        ctx.check(N.lib.ellp_b200_sync(ctx.h))
        dt = time.perf_counter() - t0
        leg = {"workload": "dense_tableau_dual_4096x12288", "pivots_per_s": piv / dt, "pivots_timed": int(piv), "rules": "reference (first infeasible row, first minimum ratio)"}
        od = N.default_opts(None, engine=N.ENGINE_TABLEAU, check_every=96, block_k=bk, pricing=N.PRICE_DEVEX)
        od.max_iter = N.U64_MAX
        ts = []
        for _ in range(4):
            ctx.check(N.lib.ellp_b200_generate_dense_ex(ctx.h, m, ns, SEED, 1, C.byref(od)))
            t0 = time.perf_counter()
            ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(od), C.byref(res)))
            ts.append(time.perf_counter() - t0)
        leg["devex_complete_solve"] = {"status": int(res.status), "pivots": int(res.iters), "ms_per_solve": 1e3 * min(ts[1:]), "objective": float(res.obj)}
        # the same complete solve END TO END through the boundary: host buffers in (pinned A), point out
        import torch
        n = m + ns
        A_h = torch.empty(m * n, dtype=torch.float64, pin_memory=True).numpy()
        c_h = np.zeros(n); b_h = np.zeros(m); lb_h = np.zeros(n); ub_h = np.zeros(n); kind_h = np.zeros(n, dtype=np.uint8)
        ctx.check(N.lib.ellp_b200_generate_dense_ex(ctx.h, m, ns, SEED, 1, C.byref(N.default_opts(1, engine=N.ENGINE_REVISED))))
        ctx.check(N.lib.ellp_b200_download_std_form(ctx.h, N.ptr(A_h), N.ptr(c_h), N.ptr(b_h), N.ptr(kind_h), N.ptr(lb_h), N.ptr(ub_h)))
        sf = N.StdForm(m, n, N.ptr(A_h), N.ptr(c_h), N.ptr(b_h), N.ptr(kind_h), N.ptr(lb_h), N.ptr(ub_h))
        te = []
        for _ in range(3):
            x = np.zeros(n); x[ns:] = -b_h
            B = np.arange(ns, n, dtype=np.int32); Nv = np.arange(ns, dtype=np.int32); Ns = np.zeros(ns, dtype=np.uint8)
            y = np.zeros(m); d = c_h.copy()
            pt = N.Point(N.ptr(x), N.ptr(B), N.ptr(Nv), N.ptr(Ns), N.ptr(y), N.ptr(d), m, ns)
            t0 = time.perf_counter()
            ctx.check(N.lib.ellp_b200_dual_solve_with_initial(ctx.h, C.byref(sf), C.byref(pt), C.byref(od), C.byref(res)))
            te.append(time.perf_counter() - t0)
        leg["devex_complete_solve"]["e2e_ms_per_solve_host_buffers"] = 1e3 * min(te[1:])
        leg["devex_complete_solve"]["e2e_objective"] = float(res.obj)
        leg["devex_complete_solve"]["e2e_primal_residual_rel"] = float(np.abs(A_h.reshape((n, m)).T @ x - b_h).max() / np.abs(b_h).max())
        del A_h
        out["configs[2]"] = leg
    except Exception as e:
        out["configs[2]"] = {"error": str(e)}
    try:  # configs[3]: batch of independent small LPs, this GPU
        nlp, m, ns = 65536, 64, 128
        ctx.check(N.lib.ellp_b200_batch_generate(ctx.h, nlp, m, ns, SEED, 0, 0))
        o = N.default_opts(None)
        br = N.BatchResult()
        ctx.check(N.lib.ellp_b200_batch_run(ctx.h, C.byref(o), C.byref(br)))
        it = np.zeros((nlp, 2), dtype=np.int32); stt = np.zeros(nlp, dtype=np.int32)
        got = N.BatchResult(N.ptr(stt), None, None, N.ptr(it), None, None, 0, None, 0.0, 0, 0)
        t0 = time.perf_counter()
        for _ in range(2):
            ctx.check(N.lib.ellp_b200_batch_run(ctx.h, C.byref(o), C.byref(br)))
        ctx.check(N.lib.ellp_b200_sync(ctx.h))
        dt = (time.perf_counter() - t0) / 2
        ctx.check(N.lib.ellp_b200_batch_download(ctx.h, C.byref(got)))
        out["configs[3]"] = {"workload": "batch_small_lps_65536x64x128 (1 GPU)", "pivots_per_s": float(got.pivots) / dt, "lps_per_s": nlp / dt,
                             "all_optimal": bool((stt == N.OPTIMAL).all()), "pivots_per_batch": int(got.pivots)}
    except Exception as e:
        out["configs[3]"] = {"error": str(e)}
    try:  # configs[0]: netlib AFIRO, whole two-phase solves through the public API
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import problems as P
        from ellp_b200.solver import GpuDualSimplexSolver, GpuPrimalSimplexSolver
        prob, exp = P.netlib("afiro")
        leg = {"workload": "netlib_afiro"}
        for tag, cls in (("primal", GpuPrimalSimplexSolver), ("dual", GpuDualSimplexSolver)):
            sol = cls.default(ctx=ctx)
            r = sol.solve(prob)
            ts = []
            for _ in range(20):
                t0 = time.perf_counter(); r = sol.solve(prob); ts.append(time.perf_counter() - t0)
            leg[tag] = {"status": r.kind, "objective": r.solution.obj() if r.is_optimal else None, "ms_per_solve": 1e3 * float(np.median(ts)),
                        "pivots": int(sum(r.iters)), "launches": int(r.launches)}
        out["configs[0]"] = leg
    except Exception as e:
        out["configs[0]"] = {"error": str(e)}
    return out


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Samples NVML in-process every
    10 ms (the timed region of a step can be shorter than one `nvidia-smi -lms` period); falls back to an nvidia-smi
    subprocess when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nv, self.stop_flag, self.thread = index, [], None, None, False, None

    def _nvml_loop(self):
        nv, h = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((sm, mx, pw, int(rs)))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates all GPUs of the box regardless of CUDA_VISIBLE_DEVICES; LOCAL_RANK == physical index under torchrun
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            self.nv = (nv, nv.nvmlDeviceGetHandleByIndex(phys))
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self) -> dict:
        if self.nv:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            nv = self.nv[0]
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            reasons = sorted({name for (_, _, _, rs) in self.rows for name, bit in bits.items() if rs & bit})
            sm = [r[0] for r in self.rows]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": float(min(sm)) if sm else None,
                    "sm_max_mhz": float(max(r[1] for r in self.rows)) if sm else None, "power_w_max": max((r[2] for r in self.rows), default=None),
                    "reasons": reasons, "samples": len(sm), "source": "nvml, 10 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_reference_sample(wl: dict, pivots: int):
    """ellp's own algorithm (oracle port) on the same generator at size sample_m x 2*sample_m, `pivots` pivots."""
    import bench_lp
    from oracle import binding as O
    ms = wl["sample_m"]
    ratio = wl["ns"] / wl["m"]
    dual = bool(wl.get("dual"))
    lp = bench_lp.dense_lp(ms, int(ms * ratio), SEED, 1 if dual else 0)
    O.lib()
    t0 = time.perf_counter()
    r = O.solve_with_initial(O.DUAL if dual else O.PRIMAL, lp["m"], lp["n"], lp["A"], lp["c"], lp["b"], lp["kind"], lp["lb"],
                             lp["ub"], lp["x"], lp["B"], lp["N"], lp["N_side"], lp.get("y"), lp.get("d"), max_iter=pivots)
    dt = time.perf_counter() - t0
    assert r.status == O.MAXITER, r.status_name
    return pivots / dt, dt, f"{pivots} pivots of the same generator at {lp['m']}x{lp['n']} (a full-size pivot needs a " \
                            f"{wl['m']}^3 dense LU: ~{(wl['m'] / ms) ** 3 * dt / pivots:.0f} s extrapolated)"


def cpu_scaling_fields(wl: dict, value: float, budget_s: float = 25.0) -> dict:
    """Structured form of the CPU sample: its size, the m^3 extrapolation to the full workload, and measured seconds per pivot
    at 2x (and, budget permitting, 4x) the sample size that back the cubic law (one dense LU per pivot, primal :173 / dual :241)."""
    ms = wl["sample_m"]
    out = {"sample_config": {"m": ms, "n": ms + int(ms * wl["ns"] / wl["m"])}, "same_config_as_gpu_arm": ms == wl["m"],
           "seconds_per_pivot": {str(ms): 1.0 / value}}
    t_used = 0.0
    for mult in (2, 4):
        mm = ms * mult
        est = (mult ** 3) / value
        if mm > wl["m"] or t_used + est > budget_s:
            break
        v, dtc, _ = cpu_reference_sample(dict(wl, sample_m=mm), 1)
        out["seconds_per_pivot"][str(mm)] = 1.0 / v
        t_used += dtc
    sizes = sorted(int(k) for k in out["seconds_per_pivot"])
    if len(sizes) >= 2:
        a, b = sizes[0], sizes[-1]
        out["measured_exponent"] = float(np.log(out["seconds_per_pivot"][str(b)] / out["seconds_per_pivot"][str(a)]) / np.log(b / a))
    big = sizes[-1]
    out["extrapolated_full_size_seconds_per_pivot"] = out["seconds_per_pivot"][str(big)] * (wl["m"] / big) ** 3
    out["extrapolated_full_size_pivots_per_s"] = 1.0 / out["extrapolated_full_size_seconds_per_pivot"]
    out["extrapolation"] = f"m^3 law from the largest timed size ({big}) to m = {wl['m']}"
    return out


def run_reference(args, wl, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P = wl["sample_pivots"]
    for _ in range(args.warmup):
        cpu_reference_sample(wl, P)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, sample = cpu_reference_sample(wl, P)
    dt = time.perf_counter() - t0
    value = args.steps * P / dt
    cpu = {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "host_cores_available": os.cpu_count(),
           "threads_note": "ellp's loop is single-threaded (nalgebra LU, no rayon): 1 core is all the reference uses"}
    cpu.update(cpu_scaling_fields(wl, value))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "m": wl["m"], "n": wl["m"] + wl["ns"], "timed_sample": cpu["sample_config"],
                       "note": "one reference pivot at the full size is a %d^3 dense LU (~%.0f s extrapolated): the timed steps run the same generator at the sample size" % (wl["m"], cpu["extrapolated_full_size_seconds_per_pivot"])},
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ batch of small LPs
def cpu_batch_sample(wl: dict, nlps: int):
    """ellp's PrimalSimplexSolver::solve (oracle port, 1 thread) on the first `nlps` LPs of the batch."""
    import bench_lp
    from oracle import binding as O
    O.lib()
    probs = [bench_lp.batch_lp_problem_arrays(wl["m"], wl["ns"], SEED, k) for k in range(nlps)]
    t0 = time.perf_counter()
    piv = 0
    for p in probs:
        r = O.solve(p, O.PRIMAL, None, O.MODE_EXACT)
        assert r.status == O.OPTIMAL
        piv += sum(r.iters)
    dt = time.perf_counter() - t0
    return piv / dt, dt, f"PrimalSimplexSolver::solve on the first {nlps} LPs of the same batch ({piv} pivots, {nlps / dt:.1f} LPs/s)"


def run_batch(args, wl, name):
    """configs[3]: every rank solves its shard of the batch with the shared-memory kernel (K6); no collective."""
    import torch
    from ellp_b200 import _native as N
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        for _ in range(min(args.warmup, 1)):
            cpu_batch_sample(wl, 5)
        t0 = time.perf_counter(); piv = 0.0
        per_step = max(4, wl["sample_lps"] // 8)
        for _ in range(args.steps):
            v, dts, sample = cpu_batch_sample(wl, per_step)
            piv += v * dts
        dt = time.perf_counter() - t0
        value = piv / dt
        emit(({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": name, "nlp": wl["nlp"], "m": wl["m"], "n_struct": wl["ns"]},
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = N.Context(local_rank)
    m, ns = wl["m"], wl["ns"]
    nlp = wl["nlp"] // world
    first = rank * nlp
    n0 = ns + m
    o = N.default_opts(None)
    ctx.check(N.lib.ellp_b200_batch_generate(ctx.h, nlp, m, ns, SEED, first, 0))
    iters = np.zeros((nlp, 2), dtype=np.int32); status = np.zeros(nlp, dtype=np.int32)

    def step():
        res = N.BatchResult()
        ctx.check(N.lib.ellp_b200_batch_run(ctx.h, C.byref(o), C.byref(res)))
        return res.ms_device

    def sync_all():
        ctx.check(N.lib.ellp_b200_sync(ctx.h)); torch.cuda.synchronize()
        if dist:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    out = N.BatchResult(N.ptr(status), None, None, N.ptr(iters), None, None, 0, None, 0.0, 0, 0)
    ctx.check(N.lib.ellp_b200_batch_download(ctx.h, C.byref(out)))
    assert (status == N.OPTIMAL).all()
    pivots_local = int(iters.sum())
    sync_all()
    clocks = ClockSampler(local_rank); clocks.start()
    t0 = time.perf_counter(); dev_ms = 0.0
    for _ in range(args.steps):
        dev_ms += step()
    ctx.check(N.lib.ellp_b200_sync(ctx.h)); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    # e2e: host buffers in, results out (ellp_b200_primal_solve_batch)
    A_h = torch.empty(nlp * m * n0, dtype=torch.float64, pin_memory=True).numpy()
    c_h = np.zeros(nlp * n0); b_h = np.zeros(nlp * m)
    e2e = None
    if not args.no_e2e:
        kind_h = np.ones(nlp * n0, dtype=np.uint8); lb_h = np.zeros(nlp * n0); ub_h = np.zeros(nlp * n0)
        ctx.check(N.lib.ellp_b200_batch_download_all(ctx.h, N.ptr(A_h), N.ptr(c_h), N.ptr(b_h)))
        bt = N.Batch(nlp, m, n0, N.ptr(A_h), N.ptr(c_h), N.ptr(b_h), N.ptr(kind_h), N.ptr(lb_h), N.ptr(ub_h))
        obj = np.zeros(nlp); st2 = np.zeros(nlp, dtype=np.int32); it2 = np.zeros((nlp, 2), dtype=np.int32); err2 = np.zeros(nlp, dtype=np.int32)
        res = N.BatchResult(N.ptr(st2), N.ptr(obj), None, N.ptr(it2), N.ptr(err2), None, 0, None, 0.0, 0, 0)

        def e2e_step():
            ctx.check(N.lib.ellp_b200_primal_solve_batch(ctx.h, C.byref(bt), C.byref(o), C.byref(res)))
            return int(res.pivots)

        e2e_step()
        sync_all()
        t0 = time.perf_counter(); pe = 0
        for _ in range(args.steps):
            pe += e2e_step()
        dte = time.perf_counter() - t0
        e2e = dict(dt=dte, pivots=pe, h2d=8 * nlp * (m * n0 + 3 * n0 + m) + nlp * n0, d2h=nlp * (8 + 4 + 8 + 4))
    vals = torch.tensor([dt, dev_ms, e2e["dt"] if e2e else 0.0], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(pivots_local), float(e2e["pivots"]) if e2e else 0.0], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dt, dev_ms, dte = [float(v) for v in vals.cpu()]
    pivots_total, pe_total = [float(v) for v in tot.cpu()]
    if rank == 0:
        value = args.steps * pivots_total / dt
        cpu = None
        if not args.no_cpu:
            v, dtc, sample = cpu_batch_sample(wl, wl["sample_lps"])
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "seconds": dtc, "host_cores_available": os.cpu_count()}
        peak, peak_src = measured_peak()
        hbm_bytes = nlp * (8.0 * (m * n0 + 3 * n0 + m) + n0 + 8.0 * (n0 + m) + 4.0 * (m + n0) + n0 + 24)
        smem_bytes_per_pivot = 16.0 * m * n0
        roofline = {"bound": "hbm", "kernel": "k_batch_primal (tableau resident in shared memory; HBM only for loading the LP and storing the point)",
                    "achieved": hbm_bytes * args.steps / (dev_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "traffic": None, "peak_source": peak_src,
                    "frac": hbm_bytes * args.steps / (dev_ms * 1e-3) / 1e9 / peak,
                    "note": "this kernel is shared-memory bound, not HBM bound: %.0f KB of shared-memory traffic per pivot, %.2f TB/s aggregate per GPU" % (
                        smem_bytes_per_pivot / 1e3, smem_bytes_per_pivot * pivots_total / world * args.steps / (dev_ms * 1e-3) / 1e12)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": name, "nlp": wl["nlp"], "lps_per_gpu": nlp, "m": m, "n_struct": ns, "std_form": f"{m}x{n0} (+{m} artificial columns in phase 1)",
                           "pivots_per_step": pivots_total, "lps_per_s": args.steps * wl["nlp"] / dt, "engine": "K6 shared-memory condensed tableau, one CTA per LP",
                           "l2": f"{8.0 * nlp * m * n0 / 1e9:.2f} GB of LP data per GPU per step (> 126 MB L2)"},
                "device_ms_per_step": dev_ms / args.steps, "gpu_launches": args.steps * world, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
                "e2e": None if not e2e else {"value": pe_total / dte, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"] * world, "d2h_bytes_per_step": e2e["d2h"] * world,
                                             "ms_per_step": 1e3 * dte / args.steps, "api": "ellp_b200_primal_solve_batch (host buffers, pinned)"}}
        emit(line)
    if dist:
        dist.barrier(); dist.destroy_process_group()
    ctx.close()


# ------------------------------------------------------------------------------------------------ netlib (configs[0], configs[1])
def run_netlib(args, wl, name):
    """A step = `reps` complete solves (Problem -> standard form -> both phases -> Solution) of one netlib LP through the public
    API, host data in, solution out: every number here is end to end.  value = pivots/s of the primal solver."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import problems as P
    from oracle import binding as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob, exp = P.netlib(wl["netlib"])
    reps = 20

    def cpu_leg(which):
        O.lib()
        ts = []
        for _ in range(max(3, reps)):
            t0 = time.perf_counter(); r = O.solve(prob, which, 1000, O.MODE_EXACT); ts.append(time.perf_counter() - t0)
        return float(np.median(ts)), int(sum(r.iters)), r

    if args.impl == "reference":
        t0 = time.perf_counter(); piv = 0
        for _ in range(args.warmup):
            cpu_leg(O.PRIMAL)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for _ in range(reps):
                r = O.solve(prob, O.PRIMAL, 1000, O.MODE_EXACT); piv += sum(r.iters)
        dt = time.perf_counter() - t0
        v = piv / dt
        emit(({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "netlib fixture (tests/golden)", "config": {"workload": name, "solves_per_step": reps},
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"{args.steps * reps} complete PrimalSimplexSolver::solve calls"},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    from ellp_b200 import _native as N
    from ellp_b200.solver import GpuDualSimplexSolver, GpuPrimalSimplexSolver
    torch.cuda.set_device(0)
    ctx = N.Context(0)
    out = {}
    for cls, which, tag in ((GpuPrimalSimplexSolver, O.PRIMAL, "primal"), (GpuDualSimplexSolver, O.DUAL, "dual")):
        sol = cls.default(ctx=ctx)
        for _ in range(max(args.warmup, 3)):
            res = sol.solve(prob)
        ts, dev, launches = [], [], 0
        clocks = ClockSampler(0); clocks.start()
        t_all = time.perf_counter()
        for _ in range(args.steps * reps):
            t0 = time.perf_counter(); res = sol.solve(prob); ts.append(time.perf_counter() - t0)
            dev.append(res.ms_device); launches += res.launches
        wall = time.perf_counter() - t_all
        clk = clocks.stop()
        cpu_t, cpu_piv, ref = cpu_leg(which)
        piv = int(sum(res.iters))
        obj = res.solution.obj()
        assert res.kind == ref.status_name == "Optimal" and abs(obj - ref.obj) <= 1e-9 * max(1.0, abs(ref.obj)), (res.kind, obj, ref.obj)
        out[tag] = dict(wall_ms_median=1e3 * float(np.median(ts)), device_ms_median=float(np.median(dev)), pivots=piv, launches_per_solve=launches / len(ts),
                        pivots_per_s_wall=piv / float(np.median(ts)), cpu_ms_median=1e3 * cpu_t, cpu_pivots=cpu_piv, cpu_pivots_per_s=cpu_piv / cpu_t,
                        objective=obj, oracle_objective=ref.obj, wall_total_s=wall, clocks=clk, launches_total=launches)
    pr = out["primal"]
    m, n = len(prob.constraints), len(prob.variables)
    alg_bytes = 8.0 * (m * (n + m) + 3 * (n + m) + m)
    line = {"metric": METRIC, "value": pr["pivots"] / (pr["device_ms_median"] * 1e-3) if pr["device_ms_median"] > 0 else pr["pivots_per_s_wall"], "unit": UNIT,
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * pr["wall_total_s"] / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "netlib fixture (tests/golden)",
            "config": {"workload": name, "solves_per_step": reps, "rows": m, "structural_columns": n, "engine": "auto: whole two-phase primal solve in one launch of k_batch_primal (K6); dual on the revised engine",
                       "l2": "working set < 100 KB: the L2 flush rule does not apply (latency-bound, one CTA)"},
            "gpu_launches": int(pr["launches_total"]), "clocks": pr["clocks"],
            "roofline": {"bound": "hbm", "kernel": "k_batch_primal (one CTA, tableau in shared memory)", "achieved": alg_bytes / (pr["device_ms_median"] * 1e-3) / 1e9 if pr["device_ms_median"] > 0 else None,
                         "peak": measured_peak()[0], "unit": "GB/s", "frac": (alg_bytes / (pr["device_ms_median"] * 1e-3) / 1e9 / measured_peak()[0]) if pr["device_ms_median"] > 0 else None,
                         "traffic": None, "note": "latency-bound by construction (27..74 rows): the figure of merit is wall time per solve, not a roofline fraction"},
            "cpu_baseline": {"value": pr["cpu_pivots_per_s"], "unit": UNIT, "cores": 1, "kind": "port", "sample": "median of complete PrimalSimplexSolver::solve calls (oracle port)", "ms_per_solve": pr["cpu_ms_median"]},
            "e2e": {"value": pr["pivots_per_s_wall"], "unit": UNIT, "h2d_bytes_per_step": int(reps * alg_bytes), "d2h_bytes_per_step": int(reps * 8 * (n + m + 4)), "ms_per_solve": pr["wall_ms_median"],
                    "api": "GpuPrimalSimplexSolver.default().solve(Problem)"},
            "solvers": out}
    emit(line)
    ctx.close()


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, wl, name):
    import torch
    from ellp_b200 import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from ellp_b200 import sharded
        return sharded.bench_main(args, wl, name, METRIC, UNIT, SEED, ClockSampler, measured_peak, cpu_reference_sample)

    torch.cuda.set_device(local_rank)
    ctx = N.Context(local_rank)
    m, ns, P = wl["m"], wl["ns"], (args.pivots or wl["pivots"])
    n = m + ns
    dual = bool(wl.get("dual"))
    tab = (not dual) or bool(wl.get("tableau"))
    engine = N.ENGINE_TABLEAU if tab else N.ENGINE_REVISED
    rules = dict(pricing=N.PRICE_STEEPEST_EDGE, ratio=N.RATIO_HARRIS) if wl.get("dse") else (dict(pricing=N.PRICE_DEVEX) if wl.get("devex") else {})
    bk = 0 if not tab else (wl.get("block_k", 0) if args.block_k < 0 else args.block_k)
    if bk > 1:
        rules["block_k"] = bk
    # revised (dual) engine: the timed steps run WITHOUT per-launch events so that its iterations replay from a CUDA graph; one
    # extra profiled step after the timed region supplies the row-reduction timing of the roofline object
    o = N.default_opts(P, engine=engine, check_every=min(P, max(16, bk)), profile=tab, **rules)
    ctx.check(N.lib.ellp_b200_generate_dense_ex(ctx.h, m, ns, SEED, 1 if dual else 0, C.byref(o)))

    full_solve = bool(wl.get("dse") or wl.get("devex"))  # steepest edge / Devex reach the optimum within a few hundred pivots: a step = one whole solve
    if full_solve:
        o.max_iter = N.U64_MAX

    def step():
        if full_solve:  # rebuild the LP in HBM (0.1 ms) and solve it to optimality
            ctx.check(N.lib.ellp_b200_generate_dense_ex(ctx.h, m, ns, SEED, 1 if dual else 0, C.byref(o)))
        res = N.Result()
        ctx.check(N.lib.ellp_b200_run(ctx.h, C.byref(o), C.byref(res)))
        if full_solve:
            assert res.status == N.OPTIMAL, res.status
        else:
            assert res.status == N.MAXITER and res.iters == P, (res.status, res.iters)
        return res

    for _ in range(args.warmup):
        step()
    ctx.check(N.lib.ellp_b200_sync(ctx.h))
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = ctx.launch_count()
    t0 = time.perf_counter()
    dev_ms = rank1_ms = 0.0
    n_rank1 = 0
    pivots_done = 0
    for _ in range(args.steps):
        r = step()
        dev_ms += r.ms_device; rank1_ms += r.ms_rank1; n_rank1 += r.n_rank1; pivots_done += r.iters
    ctx.check(N.lib.ellp_b200_sync(ctx.h))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = ctx.launch_count() - launches0
    clk = clocks.stop()
    value = pivots_done / dt
    obj_resident = float(r.obj)  # same pivots at every N (same block_k, same pivots per step) => the same number in the N > 1 lines
    P = pivots_done // args.steps
    dev_ms_timed = dev_ms
    if not tab:  # profiled pass (direct launches, CUDA events around every k_rank1) for the roofline numbers only
        o.profile = 1
        r = step()
        dev_ms, rank1_ms, n_rank1 = r.ms_device, r.ms_rank1, r.n_rank1
        o.profile = 0

    # roofline of the dominant kernel = the row reduction.  Rank-1 engine: k_rank1 over the m x n tableau (revised engine:
    # over the m x m basis inverse).  Blocked engine (block_k > 1): k_blk_flush, the deferred rank-k form T -= U V over the
    # condensed m x (n - m) tableau (only nonbasic columns are stored), once every block_k pivots.
    peak, peak_src = measured_peak()
    if bk > 1:
        k3_cols = ns
        alg_bytes = 16.0 * m * k3_cols + 8.0 * bk * (m + k3_cols)
        try:  # the version launch_rankk's auto rule launched on this shape
            flush_version = int(N.lib.ellp_b200_last_flush_kernel(ctx.h))
        except Exception:
            flush_version = 0
        fk = {1: "k_blk_flush", 3: "k_blk_flush3", 4: "k_blk_flush4", 5: "k_blk_flush5<2>", 6: "k_blk_flush5<4>", 7: "k_blk_flush6<1>", 8: "k_blk_flush6<2>",
              9: "k_blk_flush4r<3>"}.get(flush_version, "k_blk_flush")
        kernel = "%s (rank-%d row reduction T -= U V of the condensed %d x %d tableau, fp64 DMMA)" % (fk, bk, m, ns)
        traffic_key = name + ":k_blk_flush"
    else:
        k3_cols = ns if tab else m  # revised engine: K3 updates the m x m basis inverse; tableau engine: the condensed tableau
        alg_bytes = 16.0 * m * k3_cols + 8.0 * (m + k3_cols)
        kernel = "k_rank1 (rank-1 row reduction of the %s)" % ("tableau" if tab else "basis inverse")
        traffic_key = name + ":k_rank1"
    k3_ms = rank1_ms / max(n_rank1, 1)
    achieved = alg_bytes / (k3_ms * 1e-3) / 1e9 if n_rank1 else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(traffic_key)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "frac_of_8TBs_nominal": (achieved / 8000.0) if achieved else None,
                "traffic": traffic, "peak_source": peak_src, "ms_per_launch": k3_ms, "launches_timed": n_rank1,
                "algorithmic_bytes_per_launch": alg_bytes, "share_of_step_device_time": (rank1_ms / dev_ms) if dev_ms else None}
    if bk > 1 and n_rank1:
        roofline = blocked_roofline(roofline, m, k3_cols, bk, k3_ms)
    if name == DEFAULT_WORKLOAD or getattr(args, "k3_target", False):
        try:
            roofline["k3_target"] = k3_target_leg(ctx, N, peak, peak_src)
        except Exception as e:  # never lose the bench line over the extra leg
            roofline["k3_target"] = {"error": str(e)}

    # ---- e2e: the C-ABI boundary on HOST buffers (H2D + pivots + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        A_h = torch.empty(m * n, dtype=torch.float64, pin_memory=True)
        small = torch.empty(5 * n + m, dtype=torch.float64, pin_memory=True).numpy()
        c_h, b_h, lb_h, ub_h, x0 = small[:n], small[n:n + m], small[n + m:2 * n + m], small[2 * n + m:3 * n + m], small[3 * n + m:4 * n + m]
        kind_h = np.zeros(n, dtype=np.uint8)
        A_np = A_h.numpy()
        ctx.check(N.lib.ellp_b200_generate_dense_ex(ctx.h, m, ns, SEED, 1 if dual else 0, C.byref(o)))
        ctx.check(N.lib.ellp_b200_download_std_form(ctx.h, N.ptr(A_np), N.ptr(c_h), N.ptr(b_h), N.ptr(kind_h), N.ptr(lb_h), N.ptr(ub_h)))
        x0[:] = 0.0
        x0[ns:] = -b_h if dual else b_h
        y0 = np.zeros(m); d0 = c_h.copy()
        B0 = np.arange(ns, n, dtype=np.int32); N0 = np.arange(ns, dtype=np.int32); Ns0 = np.zeros(ns, dtype=np.uint8)
        sf = N.StdForm(m, n, N.ptr(A_np), N.ptr(c_h), N.ptr(b_h), N.ptr(kind_h), N.ptr(lb_h), N.ptr(ub_h))
        oe = N.default_opts(P, engine=engine, check_every=min(P, max(16, bk)), **rules)
        xs = torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()

        def e2e_step():
            xs[:] = x0
            B, Nv, Ns = B0.copy(), N0.copy(), Ns0.copy()
            y, d = y0.copy(), d0.copy()
            pt = N.Point(N.ptr(xs), N.ptr(B), N.ptr(Nv), N.ptr(Ns), N.ptr(y) if dual else None, N.ptr(d) if dual else None, m, ns)
            res = N.Result()
            fn = N.lib.ellp_b200_dual_solve_with_initial if dual else N.lib.ellp_b200_primal_solve_with_initial
            ctx.check(fn(ctx.h, C.byref(sf), C.byref(pt), C.byref(oe), C.byref(res)))
            assert (res.status == N.OPTIMAL) if full_solve else (res.status == N.MAXITER and res.iters == P)
            e2e_pivots[0] += res.iters
            return res.obj

        e2e_pivots = [0]
        if full_solve:
            oe.max_iter = N.U64_MAX
        for _ in range(min(args.warmup, 3)):
            e2e_step()
        torch.cuda.synchronize()
        e2e_pivots[0] = 0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            obj = e2e_step()
        torch.cuda.synchronize()
        dte = time.perf_counter() - t0
        # tableau engine, >= 32 MB, slack basis: only the nonbasic columns of A cross PCIe (the basis columns are verified to be
        # unit vectors on the host while the DMA runs, ellp_b200_upload); otherwise the whole matrix is uploaded
        cols_up = ns if (tab and 8.0 * m * n >= 32 * 1048576) else n
        h2d = 8 * m * cols_up + 8 * (3 * n + m) + n + 8 * n + 4 * m + 4 * ns + ns
        d2h = 8 * n + 4 * m + 4 * ns + ns + 120 * ((P + 15) // 16)
        e2e = {"value": e2e_pivots[0] / dte, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * dte / args.steps, "api": "ellp_b200_%s_solve_with_initial (host buffers, pinned)" % ("dual" if dual else "primal"),
               "objective_after_step": obj}
        del A_h

    cpu = None
    if not args.no_cpu:
        v, dtc, sample = cpu_reference_sample(wl, max(6, wl["sample_pivots"] * 20))
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "seconds": dtc,
               "host_cores_available": os.cpu_count()}
        cpu.update(cpu_scaling_fields(wl, v, budget_s=14.0))

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "m": m, "n": n, "pivots_per_step": P, "engine": "revised (explicit B^-1), dual simplex" if not tab else (("dual simplex, " if dual else "") + "condensed tableau (nonbasic columns of B^-1 A resident), " + (f"blocked: one cooperative launch per {bk} pivots + one rank-{bk} flush" if bk > 1 else "rank-1 update per pivot")),
                       "block_k": bk,
                       "tie_rule": "reference folds", "dual_rules": "steepest edge + Harris" if wl.get("dse") else ("Devex leaving row, reference ratio test" if wl.get("devex") else "reference (first infeasible / first min ratio)"), "l2": (f"A_N {8.0 * m * ns / 1e6:.0f} MB + B^-1 {8.0 * m * m / 1e6:.0f} MB streamed every pivot (> 126 MB L2, no flush)" if not tab
                              else f"condensed tableau {8.0 * m * ns / 1e9:.1f} GB >> 126 MB L2 (no L2 flush needed)"),
                       "baseline_config": "BASELINE.json configs[4]" if name == DEFAULT_WORKLOAD else "north_star / smaller variant"},
            "device_ms_per_step": dev_ms_timed / args.steps, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e, "objective_after_timed_steps": obj_resident}
    if name == DEFAULT_WORKLOAD and not args.no_other_configs:
        line["other_configs"] = other_config_legs(ctx, N)
    emit(line)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--k3-target", action="store_true", help="also time K3 (k_rank1) at the north_star target size 16384x32768 (always on for the default workload)")
    ap.add_argument("--no-other-configs", action="store_true", help="default workload: skip the short legs of configs[0], [2], [3] appended to the line")
    ap.add_argument("--owner-ratio", type=int, default=-1, help="sharded primal: 1 = only the owner of the entering column runs the ratio test, 0 = every rank, -1 = library default")
    ap.add_argument("--pivots", type=int, default=0, help="pivots per step (0 = workload default)")
    ap.add_argument("--block-k", type=int, default=-1, help="tableau engine: pivots per deferred rank-k row reduction (-1 = workload default, 0 = rank-1 engine)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    _protect_stdout()
    wl = WORKLOADS[args.workload]
    if wl.get("netlib"):
        return run_netlib(args, wl, args.workload)
    if wl.get("batch"):
        return run_batch(args, wl, args.workload)
    if args.impl == "reference":
        run_reference(args, wl, args.workload)
    else:
        run_ours(args, wl, args.workload)


if __name__ == "__main__":
    main()
