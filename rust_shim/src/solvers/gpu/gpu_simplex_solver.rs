//! `GpuPrimalSimplexSolver` / `GpuDualSimplexSolver`: same inherent methods as the CPU solvers
//! (`default()`, `new(Option<u64>)`, `solve(Problem) -> EllPResult`, `solve_with_initial`), so they slot into
//! `generate_tests!` (tests/integration_tests.rs:29-49) unchanged.  The `solve` drivers below are the reference's
//! (primal_simplex_solver.rs:32-93, dual_simplex_solver.rs:33-108); only `solve_with_initial` differs: one FFI call.
#![allow(non_snake_case)]

use super::ffi;
use crate::error::EllPError;
use crate::problem::{Bound, Problem};
use crate::solver::{EllPResult, OptimalPoint, Solution, SolutionStatus, SolverResult};
use crate::solvers::dual::dual_problem::{DualFeasiblePoint, DualPhase1, DualPhase2};
use crate::solvers::primal::primal_problem::{PrimalFeasiblePoint, PrimalPhase1, PrimalPhase2};
use crate::standard_form::{BasicPoint, Nonbasic, NonbasicBound, Point, StandardForm, StandardizedProblem};
use crate::util::EPS;

use std::ffi::CStr;
use std::sync::{Mutex, Once};

// ---- one device context per process (the C library serialises nothing by itself) ------------------------------
struct Ctx(*mut ffi::ellp_b200_ctx);
unsafe impl Send for Ctx {}
static INIT: Once = Once::new();
static mut CTX: Option<Mutex<Ctx>> = None;

fn with_ctx<T>(f: impl FnOnce(*mut ffi::ellp_b200_ctx) -> T) -> T {
    INIT.call_once(|| {
        let mut raw: *mut ffi::ellp_b200_ctx = std::ptr::null_mut();
        let device = std::env::var("ELLP_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let rc = unsafe { ffi::ellp_b200_create(device, &mut raw) };
        // no CPU fallback: the GPU solver without a GPU is an error, not a silent detour
        assert!(rc == ffi::ELLP_OK && !raw.is_null(), "ellp_b200_create failed ({}): no CUDA device?", rc);
        unsafe { CTX = Some(Mutex::new(Ctx(raw))) };
    });
    let guard = unsafe { CTX.as_ref().unwrap().lock().unwrap() };
    f(guard.0)
}

fn last_error(ctx: *mut ffi::ellp_b200_ctx) -> String {
    unsafe { CStr::from_ptr(ffi::ellp_b200_last_error(ctx)) }.to_string_lossy().into_owned()
}

fn split_bounds(bounds: &[Bound]) -> (Vec<u8>, Vec<f64>, Vec<f64>) {
    let mut kind = Vec::with_capacity(bounds.len());
    let mut lb = vec![0.; bounds.len()];
    let mut ub = vec![0.; bounds.len()];
    for (i, b) in bounds.iter().enumerate() {
        kind.push(match *b {
            Bound::Free => ffi::ELLP_FREE,
            Bound::Lower(l) => { lb[i] = l; ffi::ELLP_LOWER }
            Bound::Upper(u) => { ub[i] = u; ffi::ELLP_UPPER }
            Bound::TwoSided(l, u) => { lb[i] = l; ub[i] = u; ffi::ELLP_TWOSIDED }
            Bound::Fixed(v) => { lb[i] = v; ub[i] = v; ffi::ELLP_FIXED }
        });
    }
    (kind, lb, ub)
}

fn side_to_u8(s: NonbasicBound) -> u8 {
    match s {
        NonbasicBound::Lower => ffi::ELLP_NB_LOWER,
        NonbasicBound::Upper => ffi::ELLP_NB_UPPER,
        NonbasicBound::Free => ffi::ELLP_NB_FREE,
    }
}

fn side_from_u8(s: u8) -> NonbasicBound {
    match s {
        ffi::ELLP_NB_LOWER => NonbasicBound::Lower,
        ffi::ELLP_NB_UPPER => NonbasicBound::Upper,
        _ => NonbasicBound::Free,
    }
}

fn status_from(code: i32) -> SolutionStatus {
    match code {
        ffi::ELLP_OPTIMAL => SolutionStatus::Optimal,
        ffi::ELLP_INFEASIBLE => SolutionStatus::Infeasible,
        ffi::ELLP_UNBOUNDED => SolutionStatus::Unbounded,
        _ => SolutionStatus::MaxIter,
    }
}

/// The one FFI call behind both solvers.  `yd` = Some((y, d)) for the dual.
fn device_solve_with_initial(
    sf: &StandardForm,
    pt: &mut Point,
    yd: Option<(&mut nalgebra::DVector<f64>, &mut nalgebra::DVector<f64>)>,
    max_iter: u64,
) -> Result<SolutionStatus, EllPError> {
    let (kind, lb, ub) = split_bounds(&sf.bounds);
    let mut b_idx: Vec<i32> = pt.B.iter().map(|b| b.index as i32).collect();
    let mut n_idx: Vec<i32> = pt.N.iter().map(|n| n.index as i32).collect();
    let mut n_side: Vec<u8> = pt.N.iter().map(|n| side_to_u8(n.bound)).collect();
    let (n_b, n_n) = (b_idx.len() as i32, n_idx.len() as i32);
    // rows() == 0: solve_trivial_problem rewrites N with cols() entries (solve_trivial_problem.rs:5-96)
    n_idx.resize(sf.cols().max(n_idx.len()), 0);
    n_side.resize(n_idx.len(), 0);
    let c_sf = ffi::ellp_std_form {
        m: sf.rows() as i32,
        n: sf.cols() as i32,
        a: sf.A.as_slice().as_ptr(), // DMatrix: column-major, lda = nrows
        c: sf.c.as_slice().as_ptr(),
        b: sf.b.as_slice().as_ptr(),
        kind: kind.as_ptr(),
        lb: lb.as_ptr(),
        ub: ub.as_ptr(),
    };
    let is_dual = yd.is_some();
    let (y_ptr, d_ptr) = match yd {
        Some((y, d)) => (y.as_mut_slice().as_mut_ptr(), d.as_mut_slice().as_mut_ptr()),
        None => (std::ptr::null_mut(), std::ptr::null_mut()),
    };
    let mut c_pt = ffi::ellp_point {
        x: pt.x.as_mut_slice().as_mut_ptr(),
        b: b_idx.as_mut_ptr(),
        n: n_idx.as_mut_ptr(),
        n_side: n_side.as_mut_ptr(),
        y: y_ptr,
        d: d_ptr,
        n_b,
        n_n,
    };
    let mut o = unsafe { std::mem::zeroed::<ffi::ellp_opts>() };
    unsafe { ffi::ellp_b200_default_opts(&mut o) };
    o.max_iter = max_iter;
    let mut r = unsafe { std::mem::zeroed::<ffi::ellp_result>() };
    let (rc, msg) = with_ctx(|ctx| {
        let rc = unsafe {
            if is_dual {
                ffi::ellp_b200_dual_solve_with_initial(ctx, &c_sf, &mut c_pt, &o, &mut r)
            } else {
                ffi::ellp_b200_primal_solve_with_initial(ctx, &c_sf, &mut c_pt, &o, &mut r)
            }
        };
        (rc, if rc != ffi::ELLP_OK { last_error(ctx) } else { String::new() })
    });
    if rc == ffi::ELLP_E_ELLP {
        return Err(EllPError::new(msg)); // "invalid B, has {} elements but {} expected", "invalid B, A_B is not invertible", ...
    }
    if rc != ffi::ELLP_OK {
        panic!("{}", msg); // the reference panics at the same places with the same text
    }
    for (b, i) in pt.B.iter_mut().zip(&b_idx) {
        b.index = *i as usize;
    }
    pt.N.clear();
    for k in 0..c_pt.n_n as usize {
        pt.N.push(Nonbasic::new(n_idx[k] as usize, side_from_u8(n_side[k])));
    }
    Ok(status_from(r.status))
}

// ---- primal ---------------------------------------------------------------------------------------------------
pub struct GpuPrimalSimplexSolver {
    max_iter: u64,
}

impl std::default::Default for GpuPrimalSimplexSolver {
    fn default() -> Self {
        Self { max_iter: 1000 } // primal_simplex_solver.rs:19-23
    }
}

impl GpuPrimalSimplexSolver {
    pub fn new(max_iter: Option<u64>) -> Self {
        Self { max_iter: max_iter.unwrap_or(u64::MAX) } // :26-30
    }

    /// primal_simplex_solver.rs:32-93 with the device loop behind `solve_with_initial`
    pub fn solve(&self, prob: Problem) -> EllPResult {
        let mut phase_1: PrimalPhase1 = match prob.into() {
            Some(phase_1) => phase_1,
            None => return Ok(SolverResult::Infeasible),
        };
        let mut phase_2: PrimalPhase2 = match self.solve_with_initial(&mut phase_1)? {
            SolutionStatus::Optimal => {
                let obj = phase_1.obj();
                assert!(obj > -EPS);
                if obj < EPS {
                    phase_1.into()
                } else {
                    return Ok(SolverResult::Infeasible);
                }
            }
            SolutionStatus::Infeasible => return Ok(SolverResult::Infeasible),
            SolutionStatus::Unbounded => panic!("primal phase 1 should never be unbounded"),
            SolutionStatus::MaxIter => return Ok(SolverResult::MaxIter { obj: f64::INFINITY }),
        };
        Ok(match self.solve_with_initial(&mut phase_2)? {
            SolutionStatus::Optimal => {
                let opt_pt = OptimalPoint::new(phase_2.point.into_pt());
                SolverResult::Optimal(Solution::new(phase_2.std_form, opt_pt))
            }
            SolutionStatus::Infeasible => panic!("primal phase 2 should never be infeasible"),
            SolutionStatus::Unbounded => SolverResult::Unbounded,
            SolutionStatus::MaxIter => SolverResult::MaxIter { obj: phase_2.obj() },
        })
    }

    /// replaces primal_simplex_solver.rs:95-236 (+ pivot() :238-435)
    pub fn solve_with_initial<P>(&self, prob: &mut P) -> Result<SolutionStatus, EllPError>
    where
        P: StandardizedProblem<FeasiblePoint = PrimalFeasiblePoint>,
    {
        let (std_form, pt) = prob.unpack();
        device_solve_with_initial(std_form, &mut *pt, None, self.max_iter)
    }
}

// ---- dual -----------------------------------------------------------------------------------------------------
pub struct GpuDualSimplexSolver {
    max_iter: u64,
}

impl std::default::Default for GpuDualSimplexSolver {
    fn default() -> Self {
        Self { max_iter: 1000 } // dual_simplex_solver.rs:20-24
    }
}

impl GpuDualSimplexSolver {
    pub fn new(max_iter: Option<u64>) -> Self {
        Self { max_iter: max_iter.unwrap_or(u64::MAX) } // :27-31
    }

    /// dual_simplex_solver.rs:33-108 with the device loop behind `solve_with_initial`
    pub fn solve(&self, prob: Problem) -> EllPResult {
        let mut phase_1: DualPhase1 = match prob.into() {
            Some(phase_1) => phase_1,
            None => return Ok(SolverResult::Infeasible),
        };
        let mut phase_2: DualPhase2 = match self.solve_with_initial(&mut phase_1)? {
            SolutionStatus::Optimal => {
                let obj = phase_1.obj();
                assert!(obj < EPS);
                if obj > -EPS {
                    phase_1.into()
                } else {
                    // :50-67 (quirk Q11): the DEFAULT primal solver decides between infeasible and unbounded
                    let primal_solver = GpuPrimalSimplexSolver::default();
                    let result = primal_solver.solve(phase_1.into_orig_prob())?;
                    assert!(matches!(
                        result,
                        SolverResult::Infeasible | SolverResult::Unbounded | SolverResult::MaxIter { .. }
                    ));
                    return Ok(result);
                }
            }
            SolutionStatus::Infeasible => panic!("dual phase 1 should never be infeasible"),
            SolutionStatus::Unbounded => panic!("dual phase 1 should never be unbounded"),
            SolutionStatus::MaxIter => return Ok(SolverResult::MaxIter { obj: f64::INFINITY }),
        };
        Ok(match self.solve_with_initial(&mut phase_2)? {
            SolutionStatus::Optimal => {
                let opt_pt = OptimalPoint::new(phase_2.point.into_pt());
                SolverResult::Optimal(Solution::new(phase_2.std_form, opt_pt))
            }
            SolutionStatus::Infeasible => SolverResult::Infeasible,
            SolutionStatus::Unbounded => panic!("dual phase 2 should never return unbounded"),
            SolutionStatus::MaxIter => SolverResult::MaxIter { obj: phase_2.obj() },
        })
    }

    /// replaces dual_simplex_solver.rs:110-335
    pub fn solve_with_initial<P>(&self, prob: &mut P) -> Result<SolutionStatus, EllPError>
    where
        P: StandardizedProblem<FeasiblePoint = DualFeasiblePoint>,
    {
        let (std_form, pt) = prob.unpack();
        let DualFeasiblePoint { y, d, point } = &mut *pt;
        device_solve_with_initial(std_form, point, Some((y, d)), self.max_iter)
    }
}
