//! `extern "C"` mirror of the hot-path part of `include/ellp_b200.h`.
//! Field order and types are those of the header; `tests/test_abi_cpu.py` of the engine's repository pins the header's
//! `sizeof` / `offsetof` (std_form 56, point 56, trace_rec 32, opts 64, result 72 bytes on LP64).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct ellp_b200_ctx {
    _private: [u8; 0],
}

// Bound kinds (problem.rs:190-197), nonbasic sides (standard_form.rs:205-210), SolutionStatus (solver.rs:27-33)
pub const ELLP_FREE: u8 = 0;
pub const ELLP_LOWER: u8 = 1;
pub const ELLP_UPPER: u8 = 2;
pub const ELLP_TWOSIDED: u8 = 3;
pub const ELLP_FIXED: u8 = 4;
pub const ELLP_NB_LOWER: u8 = 0;
pub const ELLP_NB_UPPER: u8 = 1;
pub const ELLP_NB_FREE: u8 = 2;
pub const ELLP_OPTIMAL: i32 = 0;
pub const ELLP_INFEASIBLE: i32 = 1;
pub const ELLP_UNBOUNDED: i32 = 2;
pub const ELLP_MAXITER: i32 = 3;
// return codes
pub const ELLP_OK: c_int = 0;
pub const ELLP_E_ELLP: c_int = -1; // the reference returns Err(EllPError(msg))
pub const ELLP_E_PANIC: c_int = -2; // the reference panics (message preserved)
// engines
pub const ELLP_ENGINE_AUTO: i32 = 0;
pub const ELLP_ENGINE_REVISED: i32 = 1;
pub const ELLP_ENGINE_TABLEAU: i32 = 2;

/// StandardForm{c, A, b, bounds} (standard_form.rs:27-34); A column-major, lda = m (= DMatrix::as_slice()).
#[repr(C)]
pub struct ellp_std_form {
    pub m: i32,
    pub n: i32,
    pub a: *const f64,
    pub c: *const f64,
    pub b: *const f64,
    pub kind: *const u8,
    pub lb: *const f64,
    pub ub: *const f64,
}

/// Point{x, N, B} (+ y, d of DualFeasiblePoint) (standard_form.rs:20-25, dual_problem.rs:11-16); all in/out.
#[repr(C)]
pub struct ellp_point {
    pub x: *mut f64,
    pub b: *mut i32,
    pub n: *mut i32,
    pub n_side: *mut u8,
    pub y: *mut f64,
    pub d: *mut f64,
    pub n_b: i32,
    pub n_n: i32,
}

#[repr(C)]
pub struct ellp_trace_rec {
    pub phase: i32,
    pub iter: i32,
    pub entering: i32,
    pub leaving: i32,
    pub step: f64,
    pub obj: f64,
}

#[repr(C)]
pub struct ellp_opts {
    pub max_iter: u64,
    pub tie_rule: i32,
    pub engine: i32,
    pub refactor_every: i32,
    pub check_every: i32,
    pub phase_tag: i32,
    pub profile: i32,
    pub trace: *mut ellp_trace_rec,
    pub trace_cap: i64,
    pub pricing: i32,
    pub ratio: i32,
    pub block_k: i32,
}

#[repr(C)]
pub struct ellp_result {
    pub status: i32,
    pub iters: u64,
    pub obj: f64,
    pub trace_len: i64,
    pub launches: u64,
    pub ms_device: f64,
    pub ms_rank1: f64,
    pub n_rank1: u64,
    pub refactors: u64,
}

extern "C" {
    pub fn ellp_b200_create(device: c_int, out: *mut *mut ellp_b200_ctx) -> c_int;
    pub fn ellp_b200_destroy(ctx: *mut ellp_b200_ctx);
    pub fn ellp_b200_last_error(ctx: *const ellp_b200_ctx) -> *const c_char;
    pub fn ellp_b200_default_opts(o: *mut ellp_opts);
    /// replaces PrimalSimplexSolver::solve_with_initial (primal_simplex_solver.rs:95-236)
    pub fn ellp_b200_primal_solve_with_initial(
        ctx: *mut ellp_b200_ctx,
        sf: *const ellp_std_form,
        pt: *mut ellp_point,
        o: *const ellp_opts,
        r: *mut ellp_result,
    ) -> c_int;
    /// replaces DualSimplexSolver::solve_with_initial (dual_simplex_solver.rs:110-335)
    pub fn ellp_b200_dual_solve_with_initial(
        ctx: *mut ellp_b200_ctx,
        sf: *const ellp_std_form,
        pt: *mut ellp_point,
        o: *const ellp_opts,
        r: *mut ellp_result,
    ) -> c_int;
}
