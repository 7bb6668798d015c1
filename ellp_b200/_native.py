"""ctypes binding of libellp_b200.so (include/ellp_b200.h).

The library is the product: there is no Python or CPU fallback.  If the shared object is missing the
import of this module raises; if no CUDA device is present ``Context()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libellp_b200.so")

OK, E_ELLP, E_PANIC, E_CUDA, E_ARG = 0, -1, -2, -3, -4
OPTIMAL, INFEASIBLE, UNBOUNDED, MAXITER = 0, 1, 2, 3
PRIMAL, DUAL = 0, 1
TIES_REFERENCE, TIES_CANONICAL = 0, 1
ENGINE_AUTO, ENGINE_REVISED, ENGINE_TABLEAU = 0, 1, 2
PRICE_REFERENCE, PRICE_STEEPEST_EDGE, PRICE_DEVEX = 0, 1, 2
RATIO_REFERENCE, RATIO_HARRIS = 0, 1
U64_MAX = 2**64 - 1
FREE, LOWER, UPPER, TWOSIDED, FIXED = 0, 1, 2, 3, 4   # ellp_bound_kind (problem.rs:190-197)
NB_LOWER, NB_UPPER, NB_FREE = 0, 1, 2               # ellp_nb_side (standard_form.rs NonbasicBound)

TRACE_DTYPE = np.dtype([("phase", "<i4"), ("iter", "<i4"), ("entering", "<i4"), ("leaving", "<i4"),
                        ("step", "<f8"), ("obj", "<f8")])


class StdForm(C.Structure):
    _fields_ = [("m", C.c_int32), ("n", C.c_int32), ("A", C.c_void_p), ("c", C.c_void_p), ("b", C.c_void_p),
                ("kind", C.c_void_p), ("lb", C.c_void_p), ("ub", C.c_void_p)]


class Point(C.Structure):
    _fields_ = [("x", C.c_void_p), ("B", C.c_void_p), ("N", C.c_void_p), ("N_side", C.c_void_p),
                ("y", C.c_void_p), ("d", C.c_void_p), ("nB", C.c_int32), ("nN", C.c_int32)]


class Opts(C.Structure):
    _fields_ = [("max_iter", C.c_uint64), ("tie_rule", C.c_int32), ("engine", C.c_int32),
                ("refactor_every", C.c_int32), ("check_every", C.c_int32), ("phase_tag", C.c_int32),
                ("profile", C.c_int32), ("trace", C.c_void_p), ("trace_cap", C.c_int64), ("pricing", C.c_int32), ("ratio", C.c_int32),
                ("block_k", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("iters", C.c_uint64), ("obj", C.c_double), ("trace_len", C.c_int64),
                ("launches", C.c_uint64), ("ms_device", C.c_double), ("ms_rank1", C.c_double),
                ("n_rank1", C.c_uint64), ("refactors", C.c_uint64)]


class Batch(C.Structure):
    _fields_ = [("nlp", C.c_int32), ("m", C.c_int32), ("n", C.c_int32), ("A", C.c_void_p), ("c", C.c_void_p), ("b", C.c_void_p),
                ("kind", C.c_void_p), ("lb", C.c_void_p), ("ub", C.c_void_p)]


class BatchResult(C.Structure):
    _fields_ = [("status", C.c_void_p), ("obj", C.c_void_p), ("x", C.c_void_p), ("iters", C.c_void_p), ("err", C.c_void_p),
                ("trace", C.c_void_p), ("trace_cap", C.c_int32), ("trace_len", C.c_void_p), ("ms_device", C.c_double),
                ("launches", C.c_uint64), ("pivots", C.c_uint64)]


class ProblemDesc(C.Structure):
    _fields_ = [("nvars", C.c_int32), ("ncons", C.c_int32), ("obj", C.c_void_p), ("kind", C.c_void_p),
                ("lb", C.c_void_p), ("ub", C.c_void_p), ("var_id", C.c_void_p), ("row_ptr", C.c_void_p),
                ("col_id", C.c_void_p), ("coef", C.c_void_p), ("op", C.c_void_p), ("rhs", C.c_void_p)]


class Solution(C.Structure):
    _fields_ = [("status", C.c_int32), ("obj", C.c_double), ("x", C.c_void_p), ("iters", C.c_uint64 * 4),
                ("used_primal_fallback", C.c_int32), ("trace_len", C.c_int64), ("launches", C.c_uint64),
                ("ms_device", C.c_double)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(ellp_b200/csrc/build.sh).  ellp_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    sig = {
        "ellp_b200_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "ellp_b200_destroy": (None, [vp]),
        "ellp_b200_last_error": (C.c_char_p, [vp]),
        "ellp_b200_version": (C.c_char_p, []),
        "ellp_b200_launch_count": (u64, [vp]),
        "ellp_b200_reset_launch_count": (None, [vp]),
        "ellp_b200_set_tuning": (C.c_int, [vp, C.c_char_p, C.c_int]),
        "ellp_b200_default_opts": (None, [C.POINTER(Opts)]),
        "ellp_b200_primal_solve_with_initial": (C.c_int, [vp, C.POINTER(StdForm), C.POINTER(Point), C.POINTER(Opts), C.POINTER(Result)]),
        "ellp_b200_dual_solve_with_initial": (C.c_int, [vp, C.POINTER(StdForm), C.POINTER(Point), C.POINTER(Opts), C.POINTER(Result)]),
        "ellp_b200_upload": (C.c_int, [vp, C.POINTER(StdForm), C.POINTER(Point), C.c_int, C.POINTER(Opts)]),
        "ellp_b200_run": (C.c_int, [vp, C.POINTER(Opts), C.POINTER(Result)]),
        "ellp_b200_download": (C.c_int, [vp, C.POINTER(Point)]),
        "ellp_b200_solve": (C.c_int, [vp, C.POINTER(ProblemDesc), C.c_int, C.POINTER(Opts), C.POINTER(Solution)]),
        "ellp_b200_parse_mps": (C.c_int, [C.c_char_p, C.POINTER(vp), C.c_char_p]),
        "ellp_b200_model_free": (None, [vp]),
        "ellp_b200_model_desc": (None, [vp, C.POINTER(ProblemDesc)]),
        "ellp_b200_stage_new": (C.c_int, [C.POINTER(ProblemDesc), C.c_int, C.POINTER(vp), C.POINTER(C.c_int), C.c_char_p]),
        "ellp_b200_stage_free": (None, [vp]),
        "ellp_b200_stage_dims": (None, [vp] + [C.POINTER(i32)] * 7),
        "ellp_b200_stage_copy": (None, [vp] + [vp] * 12),
        "ellp_b200_dev_alloc": (C.c_int, [vp, u64, C.POINTER(vp)]),
        "ellp_b200_dev_free": (C.c_int, [vp, vp]),
        "ellp_b200_h2d": (C.c_int, [vp, vp, vp, u64]),
        "ellp_b200_d2h": (C.c_int, [vp, vp, vp, u64]),
        "ellp_b200_sync": (C.c_int, [vp]),
        "ellp_b200_dev_fill_uniform": (C.c_int, [vp, vp, u64, u64, u64, C.c_double, C.c_double]),
        "ellp_b200_primal_solve_batch": (C.c_int, [vp, C.POINTER(Batch), C.POINTER(Opts), C.POINTER(BatchResult)]),
        "ellp_b200_batch_generate": (C.c_int, [vp, i32, i32, i32, u64, i64, i32]),
        "ellp_b200_batch_upload": (C.c_int, [vp, C.POINTER(Batch), i32]),
        "ellp_b200_batch_run": (C.c_int, [vp, C.POINTER(Opts), C.POINTER(BatchResult)]),
        "ellp_b200_batch_download": (C.c_int, [vp, C.POINTER(BatchResult)]),
        "ellp_b200_batch_download_lp": (C.c_int, [vp, i32, vp, vp, vp]),
        "ellp_b200_batch_download_all": (C.c_int, [vp, vp, vp, vp]),
        "ellp_b200_comm_unique_id": (C.c_int, [C.c_char_p, vp]),
        "ellp_b200_comm_init": (C.c_int, [vp, C.c_char_p, vp, C.c_int, C.c_int]),
        "ellp_b200_sharded_generate_dense": (C.c_int, [vp, i32, i32, u64, C.POINTER(Opts)]),
        "ellp_b200_sharded_generate_dense_ex": (C.c_int, [vp, i32, i32, u64, i32, C.POINTER(Opts)]),
        "ellp_b200_sharded_upload_nonbasic_ex": (C.c_int, [vp, C.POINTER(StdForm), C.POINTER(Point), C.c_int, vp, C.POINTER(Opts)]),
        "ellp_b200_sharded_upload": (C.c_int, [vp, C.POINTER(StdForm), C.POINTER(Point), C.POINTER(Opts)]),
        "ellp_b200_phase_log": (C.c_int, [vp, C.POINTER(C.c_int64), C.c_int32]),
        "ellp_b200_last_flush_kernel": (C.c_int, [vp]),
        "ellp_b200_sharded_upload_nonbasic": (C.c_int, [vp, C.POINTER(StdForm), C.POINTER(Point), C.POINTER(Opts)]),
        "ellp_b200_generate_dense_ex": (C.c_int, [vp, i32, i32, u64, i32, C.POINTER(Opts)]),
        "ellp_b200_generate_dense": (C.c_int, [vp, i32, i32, u64, C.POINTER(Opts)]),
        "ellp_b200_download_std_form": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
        "ellp_b200_rank1_update_dev": (C.c_int, [vp, vp, i64, i64, i64, vp, i64, i32, C.POINTER(C.c_float)]),
        "ellp_b200_rank1_update": (C.c_int, [vp, vp, i64, i64, i64, vp, i64]),
        "ellp_b200_rankk_update_dev": (C.c_int, [vp, vp, i64, i64, i64, vp, vp, i64, i32, i32, C.POINTER(C.c_float)]),
        "ellp_b200_rankk_update": (C.c_int, [vp, vp, i64, i64, i64, vp, vp, i32]),
        "ellp_b200_gemv_t": (C.c_int, [vp, vp, i64, i64, i64, vp, i64, vp, vp]),
        "ellp_b200_gemv_n": (C.c_int, [vp, vp, i64, i64, i64, vp, vp]),
        "ellp_b200_invert": (C.c_int, [vp, vp, i64, vp]),
        "ellp_b200_refactor_bench": (C.c_int, [vp, i32, u64, i32, i32, C.POINTER(C.c_float)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here == the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTED = _load()


def ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ellp_b200 rc={code}: {msg}")
        self.code, self.msg = code, msg


class Context:
    """One (host thread, CUDA device) context; not thread-safe."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        rc = lib.ellp_b200_create(device, C.byref(h))
        if rc != OK:
            raise NativeError(rc, f"ellp_b200_create(device={device}) failed: no usable CUDA device (no CPU fallback exists)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib.ellp_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != OK:
            raise NativeError(rc, lib.ellp_b200_last_error(self.h).decode(errors="replace"))

    def launch_count(self) -> int:
        return int(lib.ellp_b200_launch_count(self.h))

    def set_tuning(self, key: str, value: int):
        self.check(lib.ellp_b200_set_tuning(self.h, key.encode(), int(value)))


def default_opts(max_iter: Optional[int] = 1000, tie_rule: int = TIES_REFERENCE, refactor_every: int = 0,
                 check_every: int = 0, profile: bool = False, engine: int = ENGINE_AUTO, pricing: int = 0, ratio: int = 0,
                 block_k: int = 0) -> Opts:
    o = Opts()
    lib.ellp_b200_default_opts(C.byref(o))
    o.max_iter = U64_MAX if max_iter is None else int(max_iter)
    o.tie_rule = tie_rule
    o.refactor_every = refactor_every
    o.check_every = check_every
    o.profile = 1 if profile else 0
    o.engine = engine
    o.pricing = pricing
    o.ratio = ratio
    o.block_k = block_k
    return o


def problem_desc(arr: dict):
    """arr = Problem.to_arrays(); returns (ProblemDesc, keepalive)."""
    d = ProblemDesc(arr["nvars"], arr["ncons"], ptr(arr["obj"]), ptr(arr["kind"]), ptr(arr["lb"]), ptr(arr["ub"]),
                    ptr(arr["var_id"]), ptr(arr["row_ptr"]), ptr(arr["col_id"]), ptr(arr["coef"]), ptr(arr["op"]),
                    ptr(arr["rhs"]))
    return d, arr
