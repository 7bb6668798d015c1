"""Measured fp64 denominators and library comparators on this box.

  * cuBLAS DGEMM 8192^3: burst (best of 10) and sustained (back to back for ~4 s)  -> the fp64 tensor-pipe roofline peak
  * cuBLAS DGER / DGEMV at 16384 x 32768                                        -> comparators of K3 (k_rank1) / K1 (k_gemv_t)
  * cuSOLVER DGETRF (+ DGETRS on the identity = explicit inverse) at 8192, 16384 -> comparators of K4 (refactor)

cuBLAS / cuSOLVER are called through ctypes on torch-allocated device buffers (torch is only the allocator and the
event timer here).  Library numbers are comparison baselines, never the product path.
Writes one JSON object to the path given as argv[1] (default gpurun_out/fp64_peaks.json).
"""
import ctypes as C
import json
import os
import sys
import time

import torch


def _lib(names):
    last = None
    for n in names:
        try:
            return C.CDLL(n)
        except OSError as e:  # noqa: PERF203
            last = e
    raise last


def _ev_time(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best, tot = 1e30, 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = min(best, ms)
        tot += ms
    return best, tot / reps


def _clocks():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return {"sm_mhz": pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                "sm_max_mhz": pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM),
                "power_w": pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0}
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/fp64_peaks.json"
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    torch.zeros(1, device=dev)
    cublas = _lib(["libcublas.so.12", "libcublas.so"])
    cusolver = _lib(["libcusolver.so.11", "libcusolver.so"])
    stream = torch.cuda.current_stream().cuda_stream
    hb = C.c_void_p()
    assert cublas.cublasCreate_v2(C.byref(hb)) == 0
    assert cublas.cublasSetStream_v2(hb, C.c_void_p(stream)) == 0
    hs = C.c_void_p()
    assert cusolver.cusolverDnCreate(C.byref(hs)) == 0
    assert cusolver.cusolverDnSetStream(hs, C.c_void_p(stream)) == 0
    one, mone, zero = C.c_double(1.0), C.c_double(-1.0), C.c_double(0.0)
    res = {"gpu": torch.cuda.get_device_name(0), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
           "how": "cuBLAS/cuSOLVER via ctypes on torch buffers, CUDA events on torch's current stream"}

    # ---- DGEMM 8192^3 -------------------------------------------------------------------------------------------
    n = 8192
    A = torch.rand(n, n, device=dev, dtype=torch.float64)
    B = torch.rand(n, n, device=dev, dtype=torch.float64)
    Cm = torch.zeros(n, n, device=dev, dtype=torch.float64)
    pA, pB, pC = (C.c_void_p(t.data_ptr()) for t in (A, B, Cm))

    def gemm():
        rc = cublas.cublasDgemm_v2(hb, 0, 0, n, n, n, C.byref(one), pA, n, pB, n, C.byref(zero), pC, n)
        assert rc == 0, rc
    best, _ = _ev_time(gemm, 10)
    flop = 2.0 * n ** 3
    res["dgemm_8192"] = {"ms_best": best, "tflops_burst": flop / best / 1e9}
    # sustained: back to back for ~4 s
    torch.cuda.synchronize()
    reps = max(4, int(4000.0 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks_mid = None
    e0.record()
    for k in range(reps):
        gemm()
        if k == reps // 2:
            clocks_mid = _clocks()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res["dgemm_8192"].update({"ms_sustained": ms, "tflops_sustained": flop / ms / 1e9, "sustained_reps": reps, "clocks_mid_run": clocks_mid})
    # the shape of the rank-k flush: (32768 x 56) x (56 x 32768) accumulate, C = C - U V
    for k in (32, 56, 64):
        mm = 32768 if torch.cuda.mem_get_info()[0] > 20e9 else 16384
        U = torch.rand(mm, k, device=dev, dtype=torch.float64)
        V = torch.rand(k, mm, device=dev, dtype=torch.float64)
        T = torch.rand(mm, mm, device=dev, dtype=torch.float64)
        pU, pV, pT = (C.c_void_p(t.data_ptr()) for t in (U, V, T))

        def rk():  # column-major view: T (mm x mm) -= U (mm x k, ld mm) * V (k x mm, ld k)
            rc = cublas.cublasDgemm_v2(hb, 0, 0, mm, mm, k, C.byref(mone), pU, mm, pV, k, C.byref(one), pT, mm)
            assert rc == 0, rc
        best, avg = _ev_time(rk, 5)
        res[f"dgemm_rank{k}_update_{mm}"] = {"ms_best": best, "ms_avg": avg, "tflops": 2.0 * mm * mm * k / best / 1e9,
                                             "gbs_algorithmic": (16.0 * mm * mm + 8.0 * k * 2 * mm) / best / 1e6}
        del U, V, T
    del A, B, Cm
    torch.cuda.empty_cache()

    # ---- DGER / DGEMV at 16384 x 32768 ---------------------------------------------------------------------------
    R, Cc = 16384, 32768
    E = torch.rand(Cc, R, device=dev, dtype=torch.float64)  # column-major R x Cc
    x = torch.rand(R, device=dev, dtype=torch.float64)
    y = torch.rand(Cc, device=dev, dtype=torch.float64)
    pE, px, py = (C.c_void_p(t.data_ptr()) for t in (E, x, y))

    def ger():
        rc = cublas.cublasDger_v2(hb, R, Cc, C.byref(mone), px, 1, py, 1, pE, R)
        assert rc == 0, rc
    best, avg = _ev_time(ger, 20)
    res["dger_16384x32768"] = {"ms_best": best, "ms_avg": avg, "gbs": (16.0 * R * Cc + 8.0 * (R + Cc)) / best / 1e6}

    def gemv_t():
        rc = cublas.cublasDgemv_v2(hb, 1, R, Cc, C.byref(one), pE, R, px, 1, C.byref(zero), py, 1)
        assert rc == 0, rc
    best, avg = _ev_time(gemv_t, 20)
    res["dgemv_t_16384x32768"] = {"ms_best": best, "ms_avg": avg, "gbs": (8.0 * R * Cc) / best / 1e6}

    def gemv_n():
        rc = cublas.cublasDgemv_v2(hb, 0, R, Cc, C.byref(one), pE, R, py, 1, C.byref(zero), px, 1)
        assert rc == 0, rc
    best, avg = _ev_time(gemv_n, 20)
    res["dgemv_n_16384x32768"] = {"ms_best": best, "ms_avg": avg, "gbs": (8.0 * R * Cc) / best / 1e6}
    del E, x, y
    torch.cuda.empty_cache()

    # ---- DGETRF (+ DGETRS on I = explicit inverse) ---------------------------------------------------------------
    for m in (4096, 8192, 16384):
        M0 = torch.rand(m, m, device=dev, dtype=torch.float64) + torch.eye(m, device=dev, dtype=torch.float64) * 4.0
        M = M0.clone()
        ipiv = torch.zeros(m, device=dev, dtype=torch.int32)
        info = torch.zeros(1, device=dev, dtype=torch.int32)
        lwork = C.c_int(0)
        assert cusolver.cusolverDnDgetrf_bufferSize(hs, m, m, C.c_void_p(M.data_ptr()), m, C.byref(lwork)) == 0
        work = torch.zeros(max(1, lwork.value), device=dev, dtype=torch.float64)
        I = torch.eye(m, device=dev, dtype=torch.float64)

        def getrf():
            M.copy_(M0)
            rc = cusolver.cusolverDnDgetrf(hs, m, m, C.c_void_p(M.data_ptr()), m, C.c_void_p(work.data_ptr()), C.c_void_p(ipiv.data_ptr()),
                                           C.c_void_p(info.data_ptr()))
            assert rc == 0, rc

        def copy_only():
            M.copy_(M0)

        def getrs():
            rc = cusolver.cusolverDnDgetrs(hs, 0, m, m, C.c_void_p(M.data_ptr()), m, C.c_void_p(ipiv.data_ptr()), C.c_void_p(I.data_ptr()), m,
                                           C.c_void_p(info.data_ptr()))
            assert rc == 0, rc
        t_copy, _ = _ev_time(copy_only, 3)
        t_f, _ = _ev_time(getrf, 3, warm=1)
        t_f -= t_copy
        t_s, _ = _ev_time(getrs, 1, warm=0)  # the first call inverts; later calls would multiply by the inverse again (same cost)
        res[f"getrf_{m}"] = {"ms_getrf": t_f, "tflops_getrf": (2.0 / 3.0) * m ** 3 / t_f / 1e9, "ms_getrs_identity": t_s,
                             "ms_inverse_total": t_f + t_s, "tflops_inverse_total": 2.0 * m ** 3 / (t_f + t_s) / 1e9,
                             "info": int(info.item())}
        del M0, M, I, work
        torch.cuda.empty_cache()

    res["clocks_end"] = _clocks()
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
