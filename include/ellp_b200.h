/*
 * ellp_b200.h -- C ABI of the B200-native simplex pivoting engine.
 *
 * Drop-in boundary for the hot path of kehlert/ellp (paths relative to the reference repo):
 *   PrimalSimplexSolver::solve_with_initial   src/solvers/primal/primal_simplex_solver.rs:95-236 (+ pivot :238-435)
 *   DualSimplexSolver::solve_with_initial     src/solvers/dual/dual_simplex_solver.rs:110-335
 * plus the two-phase drivers around them (primal :32-93, dual :33-108) and kernel-level entry
 * points used by the parity tests and the roofline bench.
 *
 * Conventions
 *  - plain C, plain pointers and sizes; every pointer is caller-owned and only borrowed for the
 *    duration of the call; matrices are column-major with lda = m exactly like
 *    nalgebra::DMatrix<f64>::as_slice() (standard_form.rs:27-34), so a Rust caller passes its
 *    buffers without copying.
 *  - one opaque ctx per (host thread, CUDA device); a ctx is not thread-safe, different ctxs are
 *    independent.
 *  - no aborts: every panic!/assert!/Err(EllPError) of the reference becomes a negative return
 *    code and a message retrievable with ellp_b200_last_error(); the messages of the reference's
 *    Err values are preserved verbatim (primal :125-129,:135-139,:176-178; dual :154-158,:164-168).
 *  - there is NO CPU fallback: without a CUDA device ellp_b200_create() fails with ELLP_E_CUDA.
 */
#ifndef ELLP_B200_H
#define ELLP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- encodings (values cross the ABI) ------------------------------------------------------ */
/* Bound                     src/problem.rs:190-197 */
enum { ELLP_FREE = 0, ELLP_LOWER = 1, ELLP_UPPER = 2, ELLP_TWOSIDED = 3, ELLP_FIXED = 4 };
/* NonbasicBound             src/standard_form.rs:205-210 */
enum { ELLP_NB_LOWER = 0, ELLP_NB_UPPER = 1, ELLP_NB_FREE = 2 };
/* ConstraintOp              src/problem.rs:298-303 */
enum { ELLP_LTE = 0, ELLP_EQ = 1, ELLP_GTE = 2 };
/* SolutionStatus / SolverResult   src/solver.rs:6-12,27-33 */
enum { ELLP_OPTIMAL = 0, ELLP_INFEASIBLE = 1, ELLP_UNBOUNDED = 2, ELLP_MAXITER = 3 };
/* return codes */
enum {
    ELLP_OK = 0,
    ELLP_E_ELLP = -1,    /* the reference would return Err(EllPError(msg)) */
    ELLP_E_PANIC = -2,   /* the reference would panic!/assert! (message preserved) */
    ELLP_E_CUDA = -3,    /* CUDA runtime / no device / out of memory */
    ELLP_E_ARG = -4      /* malformed arguments */
};
enum { ELLP_PRIMAL = 0, ELLP_DUAL = 1 };
/* tie rules: 0 reproduces the reference's sequential folds (primal :271-286, :379-399) exactly,
 * 1 is the order-free form of SURVEY appendix A.1/A.2 (needed when columns are sharded). */
enum { ELLP_TIES_REFERENCE = 0, ELLP_TIES_CANONICAL = 1 };
/* dual leaving row: the reference's first-infeasible rule (dual :200-236) or dual steepest edge with EXACT weights
 * w_i = ||e_i^T B^-1||^2, recomputed every pivot as a by-product of the rank-1 update's write-back (no extra pass over
 * B^-1).  Dual entering column: the reference's first-minimum ratio (dual :263-279) or Harris' two-pass test with a
 * 1e-9 tolerance (largest |alpha| among the near-minimal ratios).  The reference has neither (README.md:114 lists steepest
 * edge as TODO): parity for these rules is status / objective only. */
enum { ELLP_PRICE_REFERENCE = 0, ELLP_PRICE_STEEPEST_EDGE = 1, ELLP_PRICE_DEVEX = 2 /* tableau engines: Devex reference weights, see ellp_opts::pricing */ };
enum { ELLP_RATIO_REFERENCE = 0, ELLP_RATIO_HARRIS = 1 };
/* engines: explicit basis inverse (revised simplex) or full tableau (column-shardable) */
enum { ELLP_ENGINE_AUTO = 0, ELLP_ENGINE_REVISED = 1, ELLP_ENGINE_TABLEAU = 2 };

typedef struct ellp_b200_ctx ellp_b200_ctx;

/* ---- context ------------------------------------------------------------------------------- */
int  ellp_b200_create(int device, ellp_b200_ctx** out);
void ellp_b200_destroy(ellp_b200_ctx* ctx);
const char* ellp_b200_last_error(const ellp_b200_ctx* ctx);
const char* ellp_b200_version(void);
/* number of kernels this library launched on ctx since creation / since the last reset */
uint64_t ellp_b200_launch_count(const ellp_b200_ctx* ctx);
void     ellp_b200_reset_launch_count(ellp_b200_ctx* ctx);

/* ---- the hot-path boundary: solve_with_initial ---------------------------------------------- */
/* StandardForm{c, A, b, bounds}   src/standard_form.rs:27-34 */
typedef struct {
    int32_t m, n;          /* rows(), cols() */
    const double* A;       /* m x n, column-major, lda = m */
    const double* c;       /* n */
    const double* b;       /* m */
    const uint8_t* kind;   /* n, ELLP_FREE.. */
    const double* lb;      /* n: Lower/TwoSided lower bound, Fixed value */
    const double* ub;      /* n: Upper/TwoSided upper bound */
} ellp_std_form;

/* Point{x, N, B} (+ y, d of DualFeasiblePoint)   src/standard_form.rs:20-25, dual_problem.rs:11-16 */
typedef struct {
    double*  x;        /* n, in/out */
    int32_t* B;        /* nB, in/out: variable index per basis position */
    int32_t* N;        /* nN, in/out: variable index per nonbasic position */
    uint8_t* N_side;   /* nN, in/out: ELLP_NB_* */
    double*  y;        /* m, in/out, dual only (may be NULL for primal) */
    double*  d;        /* n, in/out, dual only (may be NULL for primal) */
    int32_t  nB, nN;   /* lengths supplied (checked like primal :124-140 / dual :153-169) */
} ellp_point;

typedef struct {
    int32_t phase, iter;       /* phase: 0 primal ph1, 1 primal ph2, 2 dual ph1, 3 dual ph2 */
    int32_t entering, leaving; /* variable indices; leaving = -1 for a bound flip */
    double  step;              /* primal lambda / dual theta_primal */
    double  obj;               /* running objective before the pivot */
} ellp_trace_rec;

typedef struct {
    uint64_t max_iter;        /* solver.max_iter: Default = 1000, new(None) = UINT64_MAX */
    int32_t  tie_rule;        /* ELLP_TIES_* */
    int32_t  engine;          /* ELLP_ENGINE_* */
    int32_t  refactor_every;  /* rebuild the basis inverse (revised engine) / the tableau T = B^-1 A_N from the resident A (tableau engines) every k
                                 pivots; 0 = library default: 100 for m <= 512 (25 for the dual and Devex-priced solves on the tableau), above that
                                 only when |A x - b|_inf drifts (see ellp_b200_set_tuning "residual_every") */
    int32_t  check_every;     /* host reads the 16-byte status every k iterations (0 = default) */
    int32_t  phase_tag;       /* value stored in trace records */
    int32_t  profile;         /* 1: record CUDA events around every rank-1 update launch */
    ellp_trace_rec* trace;    /* optional caller buffer */
    int64_t  trace_cap;
    int32_t  pricing;         /* ELLP_PRICE_*: REFERENCE = the reference's rules (primal: Dantzig with its tie fold, dual: first infeasible row);
                                 STEEPEST_EDGE = exact dual steepest edge (revised engine); DEVEX = Devex reference weights, primal entering
                                 column and dual leaving row, blocked tableau engines (block_k > 1) */
    int32_t  ratio;           /* ELLP_RATIO_*: HARRIS = Harris' two-pass ratio test with a 1e-9 tolerance: dual entering column (revised engine)
                                 and -- with bound flipping -- primal leaving row (revised engine, rank-1 / blocked tableau engines on one GPU;
                                 the blocked engine then runs kernel per phase instead of the fused cooperative kernel) */
    int32_t  block_k;         /* tableau engine: > 1 defers the row reduction and applies it as ONE rank-k update
                                 (T -= U V on the fp64 tensor pipe) every block_k pivots; 0/1 = rank-1 update per pivot.
                                 Read at upload / generate time (allocates 8*(m+n)*block_k bytes) and by ellp_b200_run
                                 (may be lowered per run).  Capped at 64.  Pivot decisions are unchanged. */
} ellp_opts;

typedef struct {
    int32_t  status;          /* ELLP_OPTIMAL.. */
    uint64_t iters;           /* pivots performed (basis changes + bound flips) */
    double   obj;             /* primal: c.x ; dual: dual_obj(y, d) */
    int64_t  trace_len;
    uint64_t launches;        /* kernels launched by this call */
    double   ms_device;       /* device time of the iteration loop (CUDA events on the ctx stream) */
    double   ms_rank1;        /* with profile=1: summed device time of the row-reduction launches (k_rank1, or k_blk_flush when block_k > 1) */
    uint64_t n_rank1;         /* with profile=1: number of row-reduction launches timed */
    uint64_t refactors;
} ellp_result;

void ellp_b200_default_opts(ellp_opts* o);

/* Replaces PrimalSimplexSolver::solve_with_initial (primal_simplex_solver.rs:95-236):
 * uploads (std_form, point), pivots on the device until Optimal / Unbounded / MaxIter, writes the
 * final x, B, N back.  pt->y/pt->d are ignored. */
int ellp_b200_primal_solve_with_initial(ellp_b200_ctx*, const ellp_std_form*, ellp_point*, const ellp_opts*, ellp_result*);
/* Replaces DualSimplexSolver::solve_with_initial (dual_simplex_solver.rs:110-335). */
int ellp_b200_dual_solve_with_initial(ellp_b200_ctx*, const ellp_std_form*, ellp_point*, const ellp_opts*, ellp_result*);

/* The same call split in three so a caller can keep the LP resident in HBM between calls
 * (bench.py times `run` for the device-resident number and the one-shot call for e2e). */
int ellp_b200_upload(ellp_b200_ctx*, const ellp_std_form*, const ellp_point*, int solver /*ELLP_PRIMAL|ELLP_DUAL*/, const ellp_opts*);
int ellp_b200_run(ellp_b200_ctx*, const ellp_opts*, ellp_result*);      /* continues from the resident point */
int ellp_b200_download(ellp_b200_ctx*, ellp_point*);

/* Builds the synthetic dense LP of the bench configs directly in HBM (never crosses PCIe) and leaves it resident:
 *   min -c.x  s.t.  A x + s = b, x, s >= 0;  A ~ U(0,1) m x n_struct, b_i ~ U(1,2)*n_struct/4, c_j ~ U(0.5,1.5)
 * standard form m x (n_struct + m), starting point = slack basis.  Values come from a counter-based generator
 * (splitmix64 of seed and the global element index).  o->engine selects the resident representation. */
int ellp_b200_generate_dense(ellp_b200_ctx*, int32_t m, int32_t n_struct, uint64_t seed, const ellp_opts* o);
/* variant 0: as above (primal feasible slack basis); variant 1: min c.x s.t. A x - s = b: the
 * slack basis is dual feasible (x_B = -b, y = 0, d = c) and the resident solver is the dual one (revised engine, or -- with
 * ELLP_ENGINE_TABLEAU -- the blocked condensed tableau of dual_blocked.cuh, generated directly as T = B^-1 A_N = -A_N). */
int ellp_b200_generate_dense_ex(ellp_b200_ctx*, int32_t m, int32_t n_struct, uint64_t seed, int32_t variant, const ellp_opts* o);
/* copies the resident standard form to host buffers (any pointer may be NULL) */
int ellp_b200_download_std_form(ellp_b200_ctx*, double* A, double* c, double* b, uint8_t* kind, double* lb, double* ub);
/* tuning knobs: "rank1_cols_per_cta", "rank1_stream_min_mb", "refactor_mode", "flush_col_steps", "flush_waves", "flush_kernel" (rank-k row
 * reduction: 0 = auto, 1 = k_blk_flush, 3 = k_blk_flush3, 4 = k_blk_flush4, 5 / 6 = k_blk_flush5<2 / 4>, 7 / 8 = k_blk_flush6<1 / 2>, 9 =
 * k_blk_flush4r<3>; all bit-identical), "flush4_min_k" (auto: pending pairs from which the 128-column kernels take over, default 24), "flush_ld" (tile
 * access mode of versions 4 .. 9, -1 = auto), "flush_stages" (ring depth, 0 = auto), "refactor_panel" (panel width of the blocked LU), "coop_pivots",
 * "coop_ctas_per_sm" (CTAs per SM of the cooperative pivot kernels, 0 = occupancy query),
 * "coop_threads", "peer_exchange", "owner_ratio" (sharded primal: 1 = only the owner of the entering column runs the ratio test), "fast_upload"
 * (0 = always upload the whole A so that the tableau stays rebuildable), "residual_every" / "residual_tol_1e12" (residual-triggered rebuild of a
 * rebuildable tableau: pivots between checks of |A x - b|_inf, tolerance in units of 1e-12 relative to 1 + |b|_inf), "batch_pipeline" (0 = upload,
 * run, download of ellp_b200_primal_solve_batch one after the other), "cuda_graphs", "small_path", "phase_timing" */
int ellp_b200_set_tuning(ellp_b200_ctx*, const char* key, int value);
/* profiling aid: after set_tuning("phase_timing", P) the fused pivot kernel logs 10 clock64() stamps per pivot (block 0,
 * thread 0; phase boundaries, see peer.cuh) for the next P pivots; this copies the first `pivots` records (10 int64 each). */
int ellp_b200_phase_log(ellp_b200_ctx*, int64_t* out, int32_t pivots);
/* version of the rank-k row reduction (K3b) that the last flush launched: the "flush_kernel" numbering above (0 = none yet) */
int ellp_b200_last_flush_kernel(ellp_b200_ctx*);

/* ---- K6: batches of independent small LPs (BASELINE.json configs[3]) -------------------------------------------
 * Every LP of the batch has the same standard-form shape m x n and no Free variable; LP k sits at offset k*m*n (A,
 * column-major, lda = m), k*n (c, kind, lb, ub), k*m (b).  One CTA per LP performs PrimalPhase1::from + phase 1 +
 * verdict + PrimalPhase2::from + phase 2 (primal_problem.rs:95-141,234-291, primal_simplex_solver.rs:32-93) with the
 * tableau resident in shared memory: one kernel launch for the whole batch.  Needs (m+1)*(n+m)*8 bytes <= ~210 KB.
 * Multi-GPU: shard the batch over processes / GPUs; no collective is involved. */
typedef struct {
    int32_t nlp, m, n;
    const double* A; const double* c; const double* b; const uint8_t* kind; const double* lb; const double* ub;
} ellp_batch;
typedef struct {
    int32_t* status;      /* nlp: ELLP_OPTIMAL.. ; -1 = not solved (see err) */
    double*  obj;         /* nlp: Solution::obj() (c.x over all standard-form columns) */
    double*  x;           /* nlp * (n + m): standard-form point incl. the m artificial columns (may be NULL) */
    int32_t* iters;       /* 2 * nlp: pivots of phase 1 and phase 2 */
    int32_t* err;         /* nlp: 0, a device panic id, -100 = Free variable present (use ellp_b200_solve), 5 = singular */
    ellp_trace_rec* trace;/* nlp * trace_cap or NULL */
    int32_t  trace_cap;
    int32_t* trace_len;   /* nlp or NULL */
    double   ms_device;
    uint64_t launches;
    uint64_t pivots;      /* sum of iters */
} ellp_batch_result;
int ellp_b200_primal_solve_batch(ellp_b200_ctx*, const ellp_batch*, const ellp_opts*, ellp_batch_result*);  /* host buffers */
/* device-resident variants (bench): synthetic batch of configs[3] built in HBM, LP ids first_lp .. first_lp+nlp-1 */
int ellp_b200_batch_generate(ellp_b200_ctx*, int32_t nlp, int32_t m, int32_t n_struct, uint64_t seed, int64_t first_lp, int32_t trace_cap);
int ellp_b200_batch_upload(ellp_b200_ctx*, const ellp_batch*, int32_t trace_cap);
int ellp_b200_batch_run(ellp_b200_ctx*, const ellp_opts*, ellp_batch_result*);      /* fills ms_device, launches */
int ellp_b200_batch_download(ellp_b200_ctx*, ellp_batch_result*);                   /* fills the non-NULL arrays, pivots */
int ellp_b200_batch_download_lp(ellp_b200_ctx*, int32_t k, double* A, double* c, double* b);
int ellp_b200_batch_download_all(ellp_b200_ctx*, double* A, double* c, double* b);   /* whole resident batch */

/* ---- column-sharded tableau: one process per GPU (BASELINE.json configs[4]) ----------------------------------
 * Two engines.  The default for block_k > 1 is the PEER-MEMORY engine described further down (exchange fused into the pivot
 * kernel, no NCCL call per pivot).  The first implementation, kept for block_k <= 1 / tuning key "peer_exchange" = 0:
 * rank g stores global columns [g*n/G, (g+1)*n/G) of the tableau and of the reduced-cost row; x, the basis list and
 * the bounds are replicated.  Per pivot: local pricing, two tiny NCCL all-gathers (arg-select with the order-free
 * tie rule), an NCCL all-reduce that broadcasts the pivot column from its owner, a replicated ratio test, and the
 * local rank-1 update.  Requires n % G == 0, m % 4 == 0 and an identity (slack) starting basis.
 * nccl_path: full path of libnccl.so.2 (may be NULL when the process already loaded it, e.g. through torch). */
int ellp_b200_comm_unique_id(const char* nccl_path, void* out128);  /* rank 0; ship the 128 bytes to the other ranks */
int ellp_b200_comm_init(ellp_b200_ctx*, const char* nccl_path, const void* id128, int rank, int nranks);
int ellp_b200_sharded_generate_dense(ellp_b200_ctx*, int32_t m, int32_t n_struct, uint64_t seed, const ellp_opts* o);
/* sf describes the GLOBAL standard form except that sf->A points at THIS RANK's column block (m x n/G, lda = m);
 * pt is the global starting point.  Afterwards ellp_b200_run / ellp_b200_download work as in the single-GPU case
 * (download rebuilds N in ascending variable order). */
int ellp_b200_sharded_upload(ellp_b200_ctx*, const ellp_std_form* sf, const ellp_point* pt, const ellp_opts* o);
/* Peer-memory engine (o->block_k > 1; ellp_b200/csrc/peer.cuh): the CONDENSED tableau (only the n - m nonbasic columns are
 * stored) is split by nonbasic POSITION, rank g holding positions [g (n-m)/G, (g+1) (n-m)/G) of pt->N; the arg-reduce and
 * the pivot-column broadcast are stores into the other ranks' memory (cudaIpc-mapped, NVLink) issued by the same
 * persistent kernel that pivots -- no NCCL call per pivot.  ellp_b200_sharded_generate_dense selects this engine when
 * o->block_k > 1 (tuning key "peer_exchange" = 0 keeps the NCCL path).  For host data:
 * sf describes the GLOBAL standard form except that sf->A holds THIS RANK's nonbasic columns in pt->N order
 * (m x (n-m)/G, lda = m); the starting basis must be the identity (slack basis; the basis columns are never passed).
 * Requires (n - m) % G == 0, m % 4 == 0, G <= 8.  ellp_b200_run / ellp_b200_download work as in the single-GPU case and
 * return N in position order. */
int ellp_b200_sharded_upload_nonbasic(ellp_b200_ctx*, const ellp_std_form* sf, const ellp_point* pt, const ellp_opts* o);
/* The same for either solver (ELLP_PRIMAL | ELLP_DUAL: the dual loop of dual_simplex_solver.rs:188-334 on the sharded tableau,
 * dual_blocked.cuh; pt->y / pt->d required) and for a DIAGONAL starting basis: basis_diag[i] = A[i, B[i]] (NULL = identity; a
 * Gte row's slack gives -1), the rows of the tableau are divided by it on the device. */
int ellp_b200_sharded_upload_nonbasic_ex(ellp_b200_ctx*, const ellp_std_form* sf, const ellp_point* pt, int solver, const double* basis_diag,
                                         const ellp_opts* o);
/* variant as in ellp_b200_generate_dense_ex (1 = dual-feasible slack basis; peer engine only) */
int ellp_b200_sharded_generate_dense_ex(ellp_b200_ctx*, int32_t m, int32_t n_struct, uint64_t seed, int32_t variant, const ellp_opts* o);

/* ---- two-phase drivers: {Primal,Dual}SimplexSolver::solve ------------------------------------ */
/* A Problem as built by Problem::add_var / add_constraint (src/problem.rs:19-106); constraints in
 * CSR over variable ids. */
typedef struct {
    int32_t nvars, ncons;
    const double*  obj;      /* nvars */
    const uint8_t* kind;     /* nvars */
    const double*  lb;       /* nvars */
    const double*  ub;       /* nvars */
    const int64_t* var_id;   /* nvars, or NULL => id = position */
    const int32_t* row_ptr;  /* ncons + 1 */
    const int64_t* col_id;   /* nnz: variable ids */
    const double*  coef;     /* nnz */
    const uint8_t* op;       /* ncons: ELLP_LTE.. */
    const double*  rhs;      /* ncons */
} ellp_problem_desc;

typedef struct {
    int32_t  status;          /* SolverResult: ELLP_OPTIMAL.. */
    double   obj;             /* Solution::obj() (solver.rs:47-49) or MaxIter{obj} */
    double*  x;               /* caller buffer, nvars: Solution::x() (solver.rs:51-53) */
    uint64_t iters[4];        /* pivots per phase, index = trace phase id */
    int32_t  used_primal_fallback;   /* dual_simplex_solver.rs:50-67 */
    int64_t  trace_len;
    uint64_t launches;
    double   ms_device;
} ellp_solution;

/* PrimalSimplexSolver::solve (primal :32-93) / DualSimplexSolver::solve (dual :33-108): standard
 * form + phase construction on the host (standard_form.rs:78-191, primal_problem.rs:80-291,
 * dual_problem.rs:89-404), both phases pivoted on the device. */
int ellp_b200_solve(ellp_b200_ctx*, const ellp_problem_desc*, int solver, const ellp_opts*, ellp_solution*);

/* parse_mps (src/parse_mps.rs:23-66) with deterministic FILE order; returns a handle whose
 * description can be passed to ellp_b200_solve. */
typedef struct ellp_b200_model ellp_b200_model;
int  ellp_b200_parse_mps(const char* text, ellp_b200_model** out, char* err256);
void ellp_b200_model_free(ellp_b200_model*);
void ellp_b200_model_desc(const ellp_b200_model*, ellp_problem_desc* out);

/* Host-side stages, exposed for stage-by-stage comparison with the oracle.
 * which: 0 = Option<StandardForm>::from(Problem), 1 = PrimalPhase1, 2 = DualPhase1. */
typedef struct ellp_b200_stage ellp_b200_stage;
int  ellp_b200_stage_new(const ellp_problem_desc*, int which, ellp_b200_stage** out, int* infeasible, char* err256);
void ellp_b200_stage_free(ellp_b200_stage*);
void ellp_b200_stage_dims(const ellp_b200_stage*, int32_t* m, int32_t* n, int32_t* nx, int32_t* nB, int32_t* nN,
                          int32_t* len_c, int32_t* len_bounds);
void ellp_b200_stage_copy(const ellp_b200_stage*, double* A, double* c, double* b, uint8_t* kind, double* lb,
                          double* ub, double* x, int32_t* B, int32_t* N, uint8_t* N_side, double* y, double* d);

/* ---- kernel-level entry points (unit parity + roofline) -------------------------------------- */
/* device buffers owned by the ctx */
int ellp_b200_dev_alloc(ellp_b200_ctx*, uint64_t bytes, void** dptr);
int ellp_b200_dev_free(ellp_b200_ctx*, void* dptr);
int ellp_b200_h2d(ellp_b200_ctx*, void* dst, const void* src, uint64_t bytes);
int ellp_b200_d2h(ellp_b200_ctx*, void* dst, const void* src, uint64_t bytes);
int ellp_b200_sync(ellp_b200_ctx*);
/* fills a device array with U(lo,hi) doubles from a counter-based generator (splitmix64 of
 * seed + element index): lets the large synthetic LPs be built in HBM without crossing PCIe */
int ellp_b200_dev_fill_uniform(ellp_b200_ctx*, double* dptr, uint64_t count, uint64_t seed, uint64_t offset, double lo, double hi);

/* K3: rank-1 row reduction of E (R x C, column-major, ld >= R, DEVICE pointers):
 *   p_j = E[r,j] / alpha[r];  E[r,j] = p_j;  E[i,j] = fma(-alpha[i], p_j, E[i,j])  (i != r)
 * replaces the per-iteration `A_B.clone().lu()` of the reference (primal :173, dual :241).
 * reps launches are timed with CUDA events on the ctx stream; *ms_avg receives the mean. */
int ellp_b200_rank1_update_dev(ellp_b200_ctx*, double* E, int64_t R, int64_t C, int64_t ld, const double* alpha,
                               int64_t r, int32_t reps, float* ms_avg);
/* same on host buffers (copies in, one update, copies out) */
int ellp_b200_rank1_update(ellp_b200_ctx*, double* E, int64_t R, int64_t C, int64_t ld, const double* alpha, int64_t r);

/* K3b: rank-k row reduction E -= U V (the deferred form of k consecutive K3 updates, see ellp_opts::block_k), fp64
 * tensor pipe (DMMA m8n8k4).  DEVICE pointers: E is R x C column-major with even ld >= R, U is R x k column-major with
 * the SAME leading dimension ld, V is k x C row-major with row stride ldv >= C; 1 <= k <= 64. */
int ellp_b200_rankk_update_dev(ellp_b200_ctx*, double* E, int64_t R, int64_t C, int64_t ld, const double* U, const double* V,
                               int64_t ldv, int32_t k, int32_t reps, float* ms_avg);
/* same on host buffers: U is R x k column-major with leading dimension R, V is k x C row-major (row stride C) */
int ellp_b200_rankk_update(ellp_b200_ctx*, double* E, int64_t R, int64_t C, int64_t ld, const double* U, const double* V, int32_t k);

/* K1/K5: y[j] = dot(M[:, cols[j]], v) for j < ncols (cols == NULL => identity); host buffers */
int ellp_b200_gemv_t(ellp_b200_ctx*, const double* M, int64_t R, int64_t C, int64_t ld, const int32_t* cols,
                     int64_t ncols, const double* v, double* y);
/* K5: y = M v (FTRAN form, M column-major R x C); host buffers */
int ellp_b200_gemv_n(ellp_b200_ctx*, const double* M, int64_t R, int64_t C, int64_t ld, const double* v, double* y);
/* K4: explicit inverse of a dense m x m matrix (column-major, host buffers); returns ELLP_E_ELLP
 * "invalid B, A_B is not invertible" when a pivot is below EPS (primal :175-179). */
int ellp_b200_invert(ellp_b200_ctx*, const double* Bmat, int64_t m, double* Binv);
/* times K4 on a random dense m x m basis built in HBM; mode 1 = Gauss-Jordan, 2 = blocked LU + DMMA */
int ellp_b200_refactor_bench(ellp_b200_ctx*, int32_t m, uint64_t seed, int32_t mode, int32_t reps, float* ms_avg);

#ifdef __cplusplus
}
#endif
#endif /* ELLP_B200_H */
