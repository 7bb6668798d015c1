/*
 * ellp_oracle.h -- C interface of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a single-threaded CPU restatement of the
 * simplex hot path of kehlert/ellp (reference cited file:line in
 * ellp_oracle.cpp).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * (ellp_b200/, include/ellp_b200.h) never links, imports or calls anything
 * declared here.
 *
 * Parity status: the reference is Rust + un-vendored nalgebra (Cargo.toml:16,
 * no lockfile) and cannot be compiled in this image, so the oracle cannot be
 * diffed against the reference binary.  It is PINNED against every golden
 * value the reference's own tests hold for this path (tests/problems/mod.rs:
 * 130-674: 25 problems x 2 solvers + 3 netlib x 2 solvers; status, objective,
 * primal point) -- see tests/test_oracle_golden.py.  Pivot sequences and
 * intermediate values are asserted by nothing in the reference: for those the
 * oracle is "parity unpinned" and is only a self-consistent restatement.
 */
#ifndef ELLP_ORACLE_H
#define ELLP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Bound kinds: problem.rs:190-197 */
enum { ORC_FREE = 0, ORC_LOWER = 1, ORC_UPPER = 2, ORC_TWOSIDED = 3, ORC_FIXED = 4 };
/* Nonbasic side: standard_form.rs:205-210 */
enum { ORC_NB_LOWER = 0, ORC_NB_UPPER = 1, ORC_NB_FREE = 2 };
/* ConstraintOp: problem.rs:298-303 */
enum { ORC_LTE = 0, ORC_EQ = 1, ORC_GTE = 2 };
/* SolverResult / SolutionStatus: solver.rs:6-12, 27-33 */
enum { ORC_OPTIMAL = 0, ORC_INFEASIBLE = 1, ORC_UNBOUNDED = 2, ORC_MAXITER = 3 };
/* return codes */
enum { ORC_OK = 0, ORC_ERR_ELLP = -1 /* Err(EllPError) */, ORC_ERR_PANIC = -2 /* panic!/assert! */ };
/* tie-rule mode */
enum { ORC_MODE_EXACT = 0 /* sequential folds as written */, ORC_MODE_CANONICAL = 1 /* order-free */ };
enum { ORC_PRIMAL = 0, ORC_DUAL = 1 };

/* A Problem as built by Problem::add_var / add_constraint (problem.rs:19-106).
 * Constraints are CSR over variable IDs (ids default to the position). */
typedef struct {
    int32_t nvars, ncons;
    const double*  obj;      /* nvars */
    const uint8_t* kind;     /* nvars */
    const double*  lb;       /* nvars (Lower/TwoSided lb, Fixed value) */
    const double*  ub;       /* nvars (Upper/TwoSided ub, Fixed value) */
    const int64_t* var_id;   /* nvars or NULL => id = position */
    const int32_t* row_ptr;  /* ncons+1 */
    const int64_t* col_id;   /* nnz */
    const double*  coef;     /* nnz */
    const uint8_t* op;       /* ncons */
    const double*  rhs;      /* ncons */
} orc_problem;

/* one record per loop iteration that produced a pivot */
typedef struct {
    int32_t phase;     /* 0 primal ph1, 1 primal ph2, 2 dual ph1, 3 dual ph2 */
    int32_t iter;      /* 0-based within the phase */
    int32_t entering;  /* std-form variable index */
    int32_t leaving;   /* std-form variable index, or -1 for a bound flip */
    double  step;      /* primal: lambda; dual: theta_primal */
    double  obj;       /* objective BEFORE the pivot (primal c.x / dual running obj) */
} orc_trace_rec;

typedef struct {
    int32_t status;          /* ORC_OPTIMAL.. */
    double  obj;             /* Solution::obj() or MaxIter{obj} */
    double* x;               /* caller buffer, nvars: Solution::x() */
    uint64_t iters[4];       /* pivots per phase (index = trace phase id) */
    int32_t used_primal_fallback;
    orc_trace_rec* trace;    /* caller buffer or NULL */
    int64_t trace_cap;
    int64_t trace_len;       /* records that WOULD have been written */
    char    err[256];
} orc_result;

/* {Primal,Dual}SimplexSolver::solve (primal_simplex_solver.rs:32-93,
 * dual_simplex_solver.rs:33-108).  max_iter: value of the solver's field
 * (default() = 1000, new(None) = UINT64_MAX). */
int ellp_oracle_solve(const orc_problem* p, int solver, uint64_t max_iter, int mode, orc_result* out);

/* Standard form as the hot loop sees it (standard_form.rs:27-34). */
typedef struct {
    int32_t m, n;
    const double* A;       /* column-major, lda = m */
    const double* c;       /* n */
    const double* b;       /* m */
    const uint8_t* kind;   /* n */
    const double* lb;      /* n */
    const double* ub;      /* n */
} orc_std_form;

typedef struct {
    double*  x;        /* n, in/out */
    int32_t* B;        /* m, in/out: variable index per basis position */
    int32_t* N;        /* n-m, in/out: variable index per nonbasic position */
    uint8_t* N_side;   /* n-m, in/out */
    double*  y;        /* m, dual only, in/out */
    double*  d;        /* n, dual only, in/out */
    int32_t  nB, nN;   /* lengths actually supplied (for the "invalid B/N" errors) */
} orc_point;

/* solve_with_initial (primal_simplex_solver.rs:95-236 / dual :110-335). */
int ellp_oracle_primal_solve_with_initial(const orc_std_form* sf, orc_point* pt, uint64_t max_iter,
                                          int mode, orc_result* out);
int ellp_oracle_dual_solve_with_initial(const orc_std_form* sf, orc_point* pt, uint64_t max_iter,
                                        int mode, orc_result* out);

/* Option<StandardForm>::from(Problem) (standard_form.rs:78-191) and the phase
 * builders, exposed so the product's host layer can be compared stage by stage.
 * which: 0 = standard form, 1 = primal phase 1, 2 = dual phase 1.
 * Returns an opaque handle (NULL with *infeasible=1 when the reference returns None). */
typedef struct orc_stage orc_stage;
orc_stage* ellp_oracle_stage_new(const orc_problem* p, int which, int* infeasible, char* err256);
void ellp_oracle_stage_free(orc_stage*);
void ellp_oracle_stage_dims(const orc_stage*, int32_t* m, int32_t* n, int32_t* nx, int32_t* nB, int32_t* nN);
/* copies: A (m*n), c (len_c = stage dependent, returned), b (m), kind/lb/ub (len_bounds), x (nx), B, N, N_side, y(m), d(n) */
void ellp_oracle_stage_copy(const orc_stage*, double* A, double* c, int32_t* len_c, double* b, uint8_t* kind,
                            double* lb, double* ub, int32_t* len_bounds, double* x, int32_t* B, int32_t* N,
                            uint8_t* N_side, double* y, double* d);

/* ---- kernel-level restatements (bit-exact targets for the CUDA kernels) ---- */

/* Rank-1 row reduction on E (R x C, column-major, ld): with p_j = rho[j] / alpha[r]
 *   E[r,j] = p_j ;  E[i,j] = fma(-alpha[i], p_j, E[i,j])  (i != r)
 * rho is the OLD pivot row (E[r,:] before the update). */
void ellp_oracle_rank1_update(double* E, int64_t R, int64_t C, int64_t ld, const double* alpha,
                              const double* rho, int64_t r);

/* y_j = dot(M[:,j], v) accumulated in the fixed order the CUDA GEMV-T kernel uses
 * (see ellp_b200/csrc/kernels.cuh: 64 interleaved partial sums, then the butterfly). */
void ellp_oracle_gemv_t(const double* M, int64_t R, int64_t C, int64_t ld, const double* v, double* y);

const char* ellp_oracle_version(void);

#ifdef __cplusplus
}
#endif
#endif
