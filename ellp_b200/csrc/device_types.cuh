// device_types.cuh -- state shared between the host driver and the sm_100a kernels.
#pragma once
#include <cstdint>
#include "../../include/ellp_b200.h"

namespace ellp {

constexpr double kEps = 0.0000000001;  // reference: src/util.rs:1

// status values stored in PivotState::status while a solve is resident on the device
constexpr int32_t kRunning = -1;
// marks a basis row that takes no part in the ratio test (|d_i| < EPS, primal :321); no ratio can have this value (a negative
// ratio, quirk Q3, is a legitimate candidate that ends in assert!(lambda >= 0.))
constexpr double kLamSkipped = -1.7976931348623157e308;  // 0..3 are the ELLP_* statuses
// device-detected panic!/assert! sites of the reference (PivotState::err)
enum DevErr : int32_t {
    kErrNone = 0,
    kErrNaNPricing = 1,        // primal_simplex_solver.rs:282 "NaN detected"
    kErrLambdaNegative = 2,    // primal_simplex_solver.rs:402 assert!(lambda >= 0.)
    kErrFlipFree = 3,          // primal_simplex_solver.rs:229 "pivot should have been unbounded"
    kErrNaNDualRatio = 4,      // dual_simplex_solver.rs:279 partial_cmp().unwrap()
    kErrSingular = 5,          // primal_simplex_solver.rs:176-178 "invalid B, A_B is not invertible"
};

// Index-level state of one iteration (reference: PivotResult/Pivot, primal_simplex_solver.rs:438-449).
// Lives in device memory; the host reads it back (64 B) every `check_every` iterations.
struct PivotState {
    int32_t status;     // kRunning or ELLP_OPTIMAL..ELLP_MAXITER
    int32_t err;        // DevErr
    uint64_t pivots;    // iterations that returned PivotResult::Pivot so far
    uint64_t max_iter;
    int32_t q_pos;      // entering: position in N
    int32_t q_var;      // entering: variable index
    int32_t q_side;     // entering: side it sat at (ELLP_NB_*)
    int32_t r_pos;      // leaving: position in B, -1 for a bound flip
    int32_t leave_var;
    int32_t new_side;   // side the leaving variable goes to
    int32_t do_update;  // 1 when the basis changed in this iteration (rank-1 update needed)
    int32_t phase_tag;
    double alpha_r;     // pivot element (B^-1 a_q)[r]
    double step;        // lambda / theta_primal
    double obj;         // running objective
    double delta;       // dual: x_r - violated bound
    double theta_d;     // dual step
    double rq;          // primal: reduced cost of the entering variable
    long long lmin_bits; // primal ratio test: min finite ratio of this iteration (bits of a non-negative double; atomicMin)
    int32_t do_step;    // 1 when x moves in this iteration (lambda > 0)
    int32_t step_pad;
    int64_t trace_len;
    int64_t trace_cap;
    // Gauss-Jordan refactorisation scratch
    int32_t gj_piv;
    int32_t gj_pad;
};

}  // namespace ellp
