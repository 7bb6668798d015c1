// batch.cuh -- K6: whole primal solves of independent small LPs, one CTA per LP, tableau resident in shared memory.
//
// BASELINE.json configs[3] (batch of 65536 independent 64x128 LPs) and the latency path for netlib-sized problems
// (one launch instead of ~7 launches per pivot).  Per LP the kernel performs what ellp's
//   PrimalPhase1::from(std_form)      src/solvers/primal/primal_problem.rs:95-141,234-253 (branch without free variables)
//   solve_with_initial (phase 1)      src/solvers/primal/primal_simplex_solver.rs:95-236 (+ pivot :238-435)
//   verdict + PrimalPhase2::from      primal_simplex_solver.rs:40-65, primal_problem.rs:263-291
//   solve_with_initial (phase 2)      primal_simplex_solver.rs:69-92
// do, i.e. PrimalSimplexSolver::solve minus the host-side standard form.
//
// Layout: the condensed tableau (nonbasic columns of B^-1 [A | artificials]) lives in shared memory, column-major with an
// ODD leading dimension (m+1 if m is even) so that both the column sweeps of the rank-1 update and the strided pivot-row
// gather are (almost) bank-conflict free.  The tie folds are the reference's sequential folds, evaluated by warp 0 with ballots exactly as in kernels.cuh.
//
// Per pivot (round 2): three block barriers -- [ratio partials] -> decision -> [x step, scaled pivot row, entering column zeroed] ->
// rank-1 sweep + reduced costs + keys / partials of the NEXT pricing -> [loop].  Decisions are replicated in every warp (REDUX on
// order-preserving keys over the 16 per-warp partials); bookkeeping runs on single threads of two otherwise idle warps.
#pragma once
#include "kernels.cuh"

namespace ellp {

constexpr int kBatchThreads = 512;

struct BatchArgs {
    int32_t nlp, m, n0, nc, ld;  // nc = columns held in shared memory (n0 + m with artificials), ld = odd leading dimension
    int32_t mode;                // 0 = two-phase from the standard form, 1 = solve_with_initial from a supplied point
    int32_t tie_rule;
    int32_t trace_cap;           // records per LP
    uint64_t max_iter;
    // per-LP strided inputs (LP k at offset k * stride)
    const double* A;             // m x n0, column-major, lda = m
    const double* c;             // n0
    const double* b;             // m
    const uint8_t* kind;         // n0
    const double* lb;            // n0
    const double* ub;            // n0
    // point: mode 1 in/out, mode 0 out
    double* x;                   // nc per LP
    int32_t* B;                  // m per LP
    int32_t* N;                  // nc - m per LP
    uint8_t* Ns;                 // nc - m per LP
    // results
    int32_t* status;             // SolverResult / SolutionStatus
    double* obj;                 // c . x (phase-2 costs)
    int32_t* iters;              // 2 per LP (phase 1, phase 2); mode 1 uses slot 1
    int32_t* err;                // DevErr
    ellp_trace_rec* trace;       // trace_cap per LP or nullptr
    int32_t* trace_len;          // per LP
};

// Condensed tableau: only the n0 NONBASIC columns of B^-1 [A | artificials] are stored (column j <-> nonbasic position j,
// variable Nv[j]); basic columns are implicit unit vectors.  A pivot replaces column q_pos by the column of the leaving
// variable (-alpha_i / alpha_r, 1 / alpha_r at row r), which is exactly what the full-tableau rank-1 update would leave
// in the leaving variable's column.  64 x 192 fp64 = 98 KB => two CTAs per SM.
struct BatchSmem {
    double* T;      // ld x n0
    double* dn;     // n0: reduced cost per nonbasic position
    double* prow;   // n0: scaled pivot row
    double* key;    // n0: Dantzig keys of the next pricing (written while the sweep still reads prow; read by the fold on ties)
    double* x;      // nc
    double* lo;     // nc
    double* hi;     // nc
    double* dcol;   // m
    double* lam;    // m
    int32_t* Bv;    // m
    int32_t* Nv;    // n0
    uint8_t* Ns;    // n0
    uint8_t* kind;  // nc
};

__host__ __device__ inline size_t batch_smem_bytes(int m, int n0, int ld) {
    const size_t nc = (size_t)n0 + m;
    size_t d = (size_t)ld * n0 + 3 * (size_t)n0 + 3 * nc + 2 * (size_t)m;  // doubles
    size_t bytes = d * 8 + 4 * ((size_t)m + n0) + (size_t)n0 + nc;
    return (bytes + 15) / 16 * 16;
}

__device__ inline BatchSmem batch_carve(unsigned char* base, int m, int n0, int ld) {
    const int nc = n0 + m;
    BatchSmem s;
    double* d = reinterpret_cast<double*>(base);
    s.T = d; d += (size_t)ld * n0;
    s.dn = d; d += n0;
    s.prow = d; d += n0;
    s.key = d; d += n0;
    s.x = d; d += nc;
    s.lo = d; d += nc;
    s.hi = d; d += nc;
    s.dcol = d; d += m;
    s.lam = d; d += m;
    int32_t* i = reinterpret_cast<int32_t*>(d);
    s.Bv = i; i += m;
    s.Nv = i; i += n0;
    uint8_t* u = reinterpret_cast<uint8_t*>(i);
    s.Ns = u; u += n0;
    s.kind = u;
    return s;
}

// cost of variable v in the current phase: phase 1 = 1 on the artificial columns (primal_problem.rs:137-141),
// phase 2 = the model's costs, 0 on the artificials (primal_problem.rs:269-276)
__device__ __forceinline__ double batch_cost(const double* __restrict__ c, int n0, int phase, int v) {
    if (phase == 0) return (v >= n0) ? 1. : 0.;
    return (v < n0) ? __ldg(c + v) : 0.;
}

// reduced costs per nonbasic position: dn_j = c_{N_j} - c_B^T T[:, j]
__device__ void batch_reduced_costs(const BatchSmem& s, const double* __restrict__ c, int m, int n0, int ld, int phase) {
    for (int i = threadIdx.x; i < m; i += kBatchThreads) s.dcol[i] = batch_cost(c, n0, phase, s.Bv[i]);
    __syncthreads();
    for (int j = threadIdx.x; j < n0; j += kBatchThreads) {
        double acc = 0.;
        const double* col = s.T + (size_t)j * ld;
        for (int i = 0; i < m; ++i) acc = fma(s.dcol[i], col[i], acc);
        s.dn[j] = batch_cost(c, n0, phase, s.Nv[j]) - acc;
    }
    __syncthreads();
}

// warp-wide max / min of NON-NEGATIVE doubles with two REDUX instructions each: for x >= +0.0 the IEEE bit pattern orders
// like an unsigned integer
__device__ __forceinline__ double warp_max_nonneg(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}
// warp-wide min of ANY doubles (no NaN): order-preserving 64-bit keys (sign bit flipped for x >= 0, all bits for x < 0).  The
// ratios of primal :327-367 are >= 0 except for the reference's TwoSided quirk ((lb - x_i) / d_i with d_i < 0 and x_i < lb),
// whose negative value must win the minimum so that assert!(lambda >= 0.) (:402) is reported as kErrLambdaNegative.
__device__ __forceinline__ double warp_min_any(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long k = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    const unsigned long long km = ((unsigned long long)mh << 32) | ml;
    return __longlong_as_double((long long)((km >> 63) ? (km & 0x7fffffffffffffffull) : ~km));
}

// Cross-warp combines: 16 lanes load one per-warp partial each and the warp reduces them with REDUX.  (Round 1 had every thread fold
// all 16 partials serially: fp64 max / min have no native instruction, each step is DSETP + 2 FSEL, and those replicated folds were
// 35 % of all instructions of the kernel -- ncu source page, profiles/r02_k6_batch_ncu_summary.txt.)

// One solve_with_initial on the condensed shared-memory tableau (primal :160-235).  Returns the SolutionStatus; all
// threads of the CTA call it and receive the same value.  dn must hold the reduced costs of the current phase on entry.
// MCT / N0CT > 0: rows / nonbasic columns known at compile time (the 64 x 192 shape of BASELINE.json configs[3]): the rank-1 sweep
// unrolls completely with immediate shared-memory offsets (4 instructions per element instead of ~10).
template <int MCT, int N0CT>
__device__ int batch_run_phase(const BatchSmem& s, int m_rt, int n0_rt, int ld_rt, uint64_t max_iter, int tie_rule, int phase_tag,
                               ellp_trace_rec* trace, int trace_cap, int* trace_len, double* obj_running, int* iters_out,
                               int* err_out) {
    __shared__ int sh_i[8];      // 0 q_pos, 2 side, 3 nb, 4 status (kRunning while pivoting), 5 err
    __shared__ double sh_d[4];   // 1 lambda
    // per-warp (best, second best, position) partials: separate arrays for pricing and for the ratio test, because no block barrier
    // separates a fast warp's ratio partial from a slow warp's read of the pricing partials any more
    __shared__ double shp1[kBatchThreads / 32], shp2[kBatchThreads / 32], shr1[kBatchThreads / 32], shr2[kBatchThreads / 32];
    __shared__ int shpi[kBatchThreads / 32], shri[kBatchThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    const int m = MCT > 0 ? MCT : m_rt, n0 = N0CT > 0 ? N0CT : n0_rt, ld = MCT > 0 ? ((MCT % 2 == 0) ? MCT + 1 : MCT) : ld_rt;
    // thread -> (row, column group) map of the rank-1 update: no division inside the sweep
    const int groups = (m <= kBatchThreads) ? kBatchThreads / m : 0;
    const int my_row = (groups > 0) ? tid % m : 0;
    const int my_group = (groups > 0) ? tid / m : 0;
    uint64_t pivots = 0;
    if (tid == 0) { sh_i[4] = kRunning; sh_i[5] = 0; }
    // Pricing keys (primal :253-270) of this thread's positions and the per-warp partials (best, second best, position of the best).
    // One reduction pass: the maximum is isolated -- nF == 1 && nBand == 0 of a two-pass formulation -- exactly when the SECOND best
    // fails the band test kmax - k < 2 EPS, because that test is monotone in k.  Runs once before the loop and then at the END of
    // every iteration, fused with the reduced-cost update (no separate pass, no barrier of its own); q_pos_prev / side_prev give the
    // new bound side of the position that was just pivoted or flipped (its Ns entry is being written by another thread).
    auto price = [&](bool update, int q_prev, int side_prev, double rq_prev, int nb_prev) {
        double a1 = 0., a2 = 0.;  // keys are > 0; +0.0 marks "no candidate" so that bit patterns order like unsigned integers
        int i1 = 0;
        for (int j = tid; j < n0; j += kBatchThreads) {
            double r = s.dn[j];
            if (update) {  // reduced costs after the pivot, and row r of the tableau = the scaled pivot row
                const double p = s.prow[j];
                r = fma(-rq_prev, p, (j == q_prev) ? 0. : r);
                s.dn[j] = r;
                s.T[(size_t)j * ld + nb_prev] = p;
            }
            const int side = (j == q_prev) ? side_prev : s.Ns[j];
            double k = -1.0;
            if (!(fabs(r) < kEps)) {
                if (r > 0. && side == ELLP_NB_UPPER) k = r;
                else if (!(r > 0.) && side == ELLP_NB_LOWER) k = -r;
                else if (side == ELLP_NB_FREE) k = fabs(r);
            }
            s.key[j] = k;  // read by the fold on ties
            if (k > a1) { a2 = a1; a1 = k; i1 = j; }
            else if (k > a2) a2 = k;  // also k == a1: the second best then equals the best ("tie")
        }
        const double w1 = warp_max_nonneg(a1);
        const bool hold = (a1 == w1);
        const unsigned hm = __ballot_sync(full, hold);
        double w2 = warp_max_nonneg(hold ? a2 : a1);
        if (__popc(hm) > 1) w2 = w1;
        const int wi = __shfl_sync(full, i1, __ffs(hm) - 1);
        if (lane == 0) { shp1[warp] = w1; shp2[warp] = w2; shpi[warp] = wi; }
    };
    price(false, -1, 0, 0., 0);
    __syncthreads();
    for (;;) {
        if (pivots >= max_iter) { if (tid == 0) sh_i[4] = ELLP_MAXITER; __syncthreads(); break; }  // :163-166
        // ---- pricing (primal :253-292): combine the per-warp partials left by price() -> CTA maximum -> is it isolated? (exact
        // shortcut); only ties / near-ties run the reference's sequential max_by fold on warp 0.
        const bool on16 = lane < kBatchThreads / 32;
        const double pb1 = on16 ? shp1[lane] : 0., pb2 = on16 ? shp2[lane] : 0.;
        const int pbi = on16 ? shpi[lane] : 0;
        const double kmax = warp_max_nonneg(pb1);
        if (kmax == 0.) {  // no candidate: optimal (:289-292); uniform across the CTA
            if (tid == 0) sh_i[4] = ELLP_OPTIMAL;
            __syncthreads();
            break;
        }
        int q_pos, q_var;
        {
            const bool hold = on16 && (pb1 == kmax);
            const unsigned hm = __ballot_sync(full, hold);
            double k2 = warp_max_nonneg(hold ? pb2 : pb1);
            if (__popc(hm) > 1) k2 = kmax;
            const int idx1 = __shfl_sync(full, pbi, __ffs(hm) - 1);
            if (tie_rule == ELLP_TIES_REFERENCE && !(kmax - k2 < 2. * kEps)) {
                q_pos = idx1;
            } else {
                if (warp == 0) {
                    int bp = -1, bv = -1;
                    if (tie_rule == ELLP_TIES_REFERENCE) {  // sequential max_by fold (primal :271-286)
                        bool have = false;
                        double bk = 0.;
                        for (int c0 = 0; c0 < n0; c0 += 32) {
                            const int j = c0 + lane;
                            const double k = (j < n0) ? s.key[j] : -1.0;
                            const int v = (j < n0) ? s.Nv[j] : 0;
                            const bool cand = (k != -1.0);
                            unsigned rem = __ballot_sync(full, cand);
                            while (rem) {
                                bool eff = false;
                                if (cand && ((rem >> lane) & 1u)) {
                                    if (!have) eff = true;
                                    else if (fabs(bk - k) >= kEps) eff = (k > bk);
                                    else eff = (v > bv);
                                }
                                const unsigned msk = __ballot_sync(full, eff);
                                if (!msk) break;
                                const int f = __ffs(msk) - 1;
                                bk = __shfl_sync(full, k, f);
                                bv = __shfl_sync(full, v, f);
                                bp = c0 + f;
                                have = true;
                                rem &= (f == 31) ? 0u : (full << (f + 1));
                            }
                        }
                    } else {  // order-free rule: largest variable index within EPS of the maximum
                        for (int j = lane; j < n0; j += 32) {
                            const double k = s.key[j];
                            if (k != -1.0 && (kmax - k < kEps) && s.Nv[j] > bv) { bv = s.Nv[j]; bp = j; }
                        }
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) {
                            const int ov = __shfl_xor_sync(full, bv, off), op = __shfl_xor_sync(full, bp, off);
                            if (ov > bv) { bv = ov; bp = op; }
                        }
                    }
                    if (lane == 0) sh_i[0] = bp;
                }
                __syncthreads();
                q_pos = sh_i[0];
            }
            q_var = s.Nv[q_pos];
        }
        const int side_q = s.Ns[q_pos];
        const bool at_lower = (side_q == ELLP_NB_LOWER);
        const double rq = s.dn[q_pos];
        // ---- pivot column, direction, ratios (primal :295-367): one row per thread
        double r1 = CUDART_INF, r2 = CUDART_INF;  // smallest / second smallest ratio of this thread's rows, position of the smallest
        int ri = 0;
        for (int i = tid; i < m; i += kBatchThreads) {
            const double a = s.T[(size_t)q_pos * ld + i];
            s.dcol[i] = a;
            const double d_i = at_lower ? -a : a;
            double lam = kLamSkipped;
            if (!(fabs(d_i) < kEps)) {
                const int var = s.Bv[i];
                lam = primal_ratio(s.kind[var], s.lo[var], s.hi[var], s.x[var], d_i);
                if (lam == 0.) lam = 0.;  // canonical +0.0
                if (lam < r1) { r2 = r1; r1 = lam; ri = i; }
                else if (lam < r2) r2 = lam;  // also lam == r1 ("tie"); +inf ratios never enter, as in the two-pass version
            }
            s.lam[i] = lam;
        }
        {
            const double w1 = warp_min_any(r1);
            const bool hold = (r1 == w1);
            const unsigned hm = __ballot_sync(full, hold);
            double w2 = warp_min_any(hold ? r2 : r1);
            if (__popc(hm) > 1) w2 = w1;
            const int wi = __shfl_sync(full, ri, __ffs(hm) - 1);
            if (lane == 0) { shr1[warp] = w1; shr2[warp] = w2; shri[warp] = wi; }
        }
        __syncthreads();
        const double rb1 = on16 ? shr1[lane] : CUDART_INF, rb2 = on16 ? shr2[lane] : CUDART_INF;
        const int rbi = on16 ? shri[lane] : 0;
        const double lmin = warp_min_any(rb1);
        double lambda;
        {
            const int kq = s.kind[q_var];  // :305-311
            lambda = (kq == ELLP_TWOSIDED) ? (s.hi[q_var] - s.lo[q_var]) : (kq == ELLP_FIXED ? 0. : CUDART_INF);
        }
        int nb = -1;
        if (tie_rule == ELLP_TIES_REFERENCE) {
            const double L = fmin(lmin, lambda);
            bool fast = !(L < CUDART_INF);  // nothing finite: lambda stays +inf
            if (!fast) {
                // second smallest of the multiset {ratios} + {lambda of the entering variable}: the minimum is isolated (the fold's result
                // is known without running it) iff that second value is at least 2 EPS above L (nF + f0 == 1 && nBand + band0 == 0)
                const bool hold = on16 && (rb1 == lmin);
                const unsigned hm = __ballot_sync(full, hold);
                double l2 = warp_min_any(hold ? rb2 : rb1);
                if (__popc(hm) > 1) l2 = lmin;
                const int idx1 = __shfl_sync(full, rbi, __ffs(hm) - 1);
                const bool q_is_min = lambda < lmin;
                const double second = q_is_min ? lmin : ((lambda == lmin) ? lmin : fmin(l2, lambda));
                if (!(second < L + 2. * kEps)) {
                    fast = true;
                    if (!q_is_min) { nb = idx1; lambda = lmin; }
                }
            }
            if (!fast) {  // sequential scan with its (lambda, new_basic, new_basic_index) state (primal :379-399), warp 0
                if (warp == 0) {
                    bool have_nbi = false;
                    int nbi = 0;
                    for (int c0 = 0; c0 < m; c0 += 32) {
                        const int i = c0 + lane;
                        const double l = (i < m) ? s.lam[i] : kLamSkipped;
                        const int v = (i < m) ? s.Bv[i] : 0;
                        const bool cand = (l != kLamSkipped) && (l < CUDART_INF);
                        unsigned rem = __ballot_sync(full, cand);
                        while (rem) {
                            int eff = 0;
                            if (cand && ((rem >> lane) & 1u)) {
                                if (l < lambda - kEps) eff = 1;
                                else if (fabs(l - lambda) < kEps && (!have_nbi || v < nbi)) eff = 2;
                            }
                            const unsigned msk = __ballot_sync(full, eff != 0);
                            if (!msk) break;
                            const int f = __ffs(msk) - 1;
                            const int kind_f = __shfl_sync(full, eff, f);
                            lambda = __shfl_sync(full, l, f);
                            nb = c0 + f;
                            if (kind_f == 2) { have_nbi = true; nbi = __shfl_sync(full, v, f); }
                            rem &= (f == 31) ? 0u : (full << (f + 1));
                        }
                    }
                    if (lane == 0) { sh_i[3] = nb; sh_d[1] = lambda; }
                }
                __syncthreads();
                nb = sh_i[3];
                lambda = sh_d[1];
            }
        } else if (lmin < lambda + kEps && lmin < CUDART_INF) {
            // order-free rule: smallest variable index within EPS of the minimum (every thread scans: m is small)
            int bestv = 0x7fffffff, bestp = -1;
            for (int i = 0; i < m; ++i) {
                const double l = s.lam[i];
                if (l != kLamSkipped && (l - lmin < kEps) && s.Bv[i] < bestv) { bestv = s.Bv[i]; bestp = i; }
            }
            nb = bestp;
            lambda = s.lam[nb];
        }
        // ---- step decision (primal :402-406, :229); identical on every thread
        int stop = kRunning, stop_err = 0;
        if (!(lambda >= 0.)) { stop = ELLP_UNBOUNDED; stop_err = kErrLambdaNegative; }
        else if (isinf(lambda)) stop = ELLP_UNBOUNDED;
        else if (nb < 0 && side_q == ELLP_NB_FREE) { stop = ELLP_UNBOUNDED; stop_err = kErrFlipFree; }
        if (stop != kRunning) {
            if (tid == 0) { sh_i[4] = stop; if (stop_err) sh_i[5] = stop_err; }
            __syncthreads();
            break;
        }
        // ---- step (primal :408-417) and scaled pivot row (1 / alpha_r in the slot of the entering position)
        const int leave_var = (nb >= 0) ? s.Bv[nb] : -1;  // read before the barrier: the swap below runs concurrently with the trace record
        if (lambda > 0.)
            for (int i = tid; i < m; i += kBatchThreads) {
                const double a = s.dcol[i];
                const double d_i = at_lower ? -a : a;
                const int var = s.Bv[i];
                s.x[var] = s.x[var] + lambda * d_i;
            }
        int new_side_q;  // bound side of position q_pos after this iteration (:217-231), known to every thread
        if (nb >= 0) {
            const double alpha_r = s.dcol[nb];
            const double d_nb = at_lower ? -alpha_r : alpha_r;
            new_side_q = (d_nb > 0.) ? ELLP_NB_UPPER : ELLP_NB_LOWER;
            for (int j = tid; j < n0; j += kBatchThreads) s.prow[j] = ((j == q_pos) ? 1. : s.T[(size_t)j * ld + nb]) / alpha_r;
            // the entering position's column restarts from the (implicit) unit column e_r of the leaving variable: zeros outside row r
            // (its old content lives in dcol; the row gather above skips it), so that the sweep below needs no per-element select
            for (int i = tid; i < m; i += kBatchThreads) s.T[(size_t)q_pos * ld + i] = 0.;
        } else {
            new_side_q = (side_q == ELLP_NB_LOWER) ? ELLP_NB_UPPER : ELLP_NB_LOWER;
        }
        __syncthreads();
        // bookkeeping by single threads of two warps that carry no reduced-cost work (warps 0 .. n0/32 do): entering value, trace
        // record and running objective on one, the index swap / bound flip (:208-231) on the other
        if (tid == kBatchThreads - 64) {
            if (lambda > 0.) s.x[q_var] = at_lower ? s.x[q_var] + lambda : s.x[q_var] - lambda;
            if (trace && *trace_len < trace_cap) {
                ellp_trace_rec rec;
                rec.phase = phase_tag;
                rec.iter = (int32_t)pivots;
                rec.entering = q_var;
                rec.leaving = leave_var;
                rec.step = lambda;
                rec.obj = *obj_running;
                trace[*trace_len] = rec;
            }
            *trace_len += 1;
            *obj_running = *obj_running + rq * (at_lower ? lambda : -lambda);
        }
        if (tid == kBatchThreads - 32) {
            if (nb >= 0) { s.Bv[nb] = q_var; s.Nv[q_pos] = leave_var; }
            s.Ns[q_pos] = (uint8_t)new_side_q;
        }
        // ---- rank-1 update of the condensed tableau (rows other than r; row r = the scaled pivot row is written by price())
        if (nb >= 0) {
            if (groups > 0) {
                if (my_group < groups && my_row != nb) {
                    const double na = -s.dcol[my_row];
                    constexpr int G = MCT > 0 ? kBatchThreads / MCT : 1, LD = (MCT % 2 == 0) ? MCT + 1 : MCT, IT = (N0CT + G - 1) / G;
                    if (MCT > 0 && N0CT > 0) {
                        // compile-time shape: 24 x {LDS, LDS, DFMA, STS} with immediate offsets.  (Contiguous column blocks per thread with
                        // 16-byte loads of the pivot-row pairs were measured 8 % SLOWER: 65.0 vs 70.3 M pivots/s.)
                        double* __restrict__ t = s.T + (size_t)my_group * LD + my_row;
                        const double* __restrict__ pp = s.prow + my_group;
#pragma unroll
                        for (int u = 0; u < IT; ++u)
                            if (N0CT % G == 0 || my_group + u * G < N0CT) t[u * G * LD] = fma(na, pp[u * G], t[u * G * LD]);
                    } else {
                        double* __restrict__ t = s.T + (size_t)my_group * ld + my_row;
                        const double* __restrict__ pp = s.prow + my_group;
                        const int tstep = groups * ld;
#pragma unroll 4
                        for (int j = my_group; j < n0; j += groups) { *t = fma(na, *pp, *t); t += tstep; pp += groups; }
                    }
                }
            } else {
                const int total = m * n0;
                for (int e = tid; e < total; e += kBatchThreads) {
                    const int j = e / m, i = e - j * m;
                    if (i == nb) continue;
                    double* t = s.T + (size_t)j * ld + i;
                    *t = fma(-s.dcol[i], s.prow[j], *t);
                }
            }
        }
        // reduced costs, row r, and the keys / partials of the NEXT pricing
        price(nb >= 0, q_pos, new_side_q, rq, nb >= 0 ? nb : 0);
        pivots += 1;
        __syncthreads();
    }
    __syncthreads();
    const int status = sh_i[4];
    if (tid == 0) { *iters_out = (int)pivots; if (sh_i[5]) *err_out = sh_i[5]; }
    __syncthreads();
    return status;
}

template <int MCT, int N0CT>
__global__ void __launch_bounds__(kBatchThreads, 2) k_batch_primal(BatchArgs a) {
    extern __shared__ __align__(16) unsigned char batch_smem[];
    __shared__ int sh_tl;
    __shared__ double sh_obj;
    __shared__ int sh_it[2], sh_err, sh_flag;
    const int tid = threadIdx.x;
    const int m = a.m, n0 = a.n0, nc = a.nc, ld = a.ld;
    const BatchSmem s = batch_carve(batch_smem, m, n0, ld);
    for (int lp = blockIdx.x; lp < a.nlp; lp += gridDim.x) {
        const double* A = a.A + (size_t)lp * m * n0;
        const double* c = a.c + (size_t)lp * n0;
        const double* b = a.b + (size_t)lp * m;
        const uint8_t* kind = a.kind + (size_t)lp * n0;
        const double* lb = a.lb + (size_t)lp * n0;
        const double* ub = a.ub + (size_t)lp * n0;
        ellp_trace_rec* trace = a.trace ? a.trace + (size_t)lp * a.trace_cap : nullptr;
        if (tid == 0) { sh_tl = 0; sh_obj = 0.; sh_it[0] = sh_it[1] = 0; sh_err = 0; sh_flag = 0; }
        // ---- load the standard form into shared memory
        for (int e = tid; e < m * n0; e += kBatchThreads) { const int j = e / m, i = e - j * m; s.T[(size_t)j * ld + i] = A[e]; }
        for (int j = tid; j < n0; j += kBatchThreads) { s.kind[j] = kind[j]; s.lo[j] = lb[j]; s.hi[j] = ub[j]; }
        __syncthreads();
        // ---- PrimalPhase1::from (primal_problem.rs:95-141, :234-253): every variable nonbasic at a bound, one artificial per
        // row with column signum(b~_i) e_i and value |b~_i|, costs 0 / 1.  Free variables are not handled here (flagged).
        for (int j = tid; j < n0; j += kBatchThreads) {
            const int k = s.kind[j];
            if (k == ELLP_FREE) sh_flag = 1;
            s.x[j] = (k == ELLP_UPPER) ? s.hi[j] : s.lo[j];
            s.Nv[j] = j;
            s.Ns[j] = (k == ELLP_UPPER) ? ELLP_NB_UPPER : ELLP_NB_LOWER;
        }
        __syncthreads();
        if (sh_flag) {
            if (tid == 0) { a.status[lp] = -1; a.err[lp] = -100; a.iters[2 * lp] = a.iters[2 * lp + 1] = 0; }
            __syncthreads();
            continue;
        }
        for (int i = tid; i < m; i += kBatchThreads) {  // b~ = b - A v  (:236)
            double bt = b[i];
            for (int j = 0; j < n0; ++j) bt -= s.T[(size_t)j * ld + i] * s.x[j];
            const int col = n0 + i;
            s.dcol[i] = signbit(bt) ? -1. : 1.;  // f64::signum (+0 -> 1)
            s.x[col] = fabs(bt);
            s.Bv[i] = col;
            s.kind[col] = ELLP_LOWER;
            s.lo[col] = 0.;
            s.hi[col] = 0.;
        }
        __syncthreads();
        // tableau of the artificial basis B = diag(sg): row i of A scaled by sg_i (B^-1 = diag(sg))
        for (int e = tid; e < m * n0; e += kBatchThreads) {
            const int j = e / m, i = e - j * m;
            if (s.dcol[i] < 0.) s.T[(size_t)j * ld + i] = -s.T[(size_t)j * ld + i];
        }
        __syncthreads();
        batch_reduced_costs(s, c, m, n0, ld, 0);
        if (tid == 0) { double o = 0.; for (int i = 0; i < m; ++i) o += s.x[n0 + i]; sh_obj = o; }
        __syncthreads();
        int status = batch_run_phase<MCT, N0CT>(s, m, n0, ld, a.max_iter, a.tie_rule, 0, trace, a.trace_cap, &sh_tl, &sh_obj, &sh_it[0], &sh_err);
        int result = -1;
        if (status == ELLP_OPTIMAL) {  // primal_simplex_solver.rs:41-52
            if (tid == 0) { double o = 0.; for (int j = n0; j < nc; ++j) o += s.x[j]; sh_obj = o; }
            __syncthreads();
            const double o = sh_obj;
            if (!(o > -kEps)) { if (tid == 0) sh_err = -102; result = ELLP_INFEASIBLE; }  // assert!(obj > -EPS)
            else if (!(o < kEps)) result = ELLP_INFEASIBLE;
        } else if (status == ELLP_INFEASIBLE) result = ELLP_INFEASIBLE;
        else if (status == ELLP_UNBOUNDED) { if (tid == 0 && !sh_err) sh_err = -101; result = ELLP_UNBOUNDED; }  // "phase 1 should never be unbounded"
        else result = ELLP_MAXITER;
        if (result >= 0) {
            if (tid == 0) { a.status[lp] = result; a.obj[lp] = (result == ELLP_MAXITER) ? CUDART_INF : 0.; }
        } else {
            // ---- PrimalPhase2::from (primal_problem.rs:263-291): artificials become Fixed(0) with cost 0, costs restored
            __syncthreads();
            for (int j = n0 + tid; j < nc; j += kBatchThreads) { s.kind[j] = ELLP_FIXED; s.lo[j] = 0.; s.hi[j] = 0.; }
            __syncthreads();
            batch_reduced_costs(s, c, m, n0, ld, 1);
            if (tid == 0) { double o = 0.; for (int j = 0; j < n0; ++j) o += c[j] * s.x[j]; sh_obj = o; }
            __syncthreads();
            status = batch_run_phase<MCT, N0CT>(s, m, n0, ld, a.max_iter, a.tie_rule, 1, trace, a.trace_cap, &sh_tl, &sh_obj, &sh_it[1], &sh_err);
            if (tid == 0) {
                double o = 0.;
                for (int j = 0; j < n0; ++j) o += c[j] * s.x[j];  // Solution::obj() = c . x (artificial costs are 0)
                a.status[lp] = status;
                a.obj[lp] = o;
            }
        }
        __syncthreads();
        // ---- write the point back
        double* xo = a.x + (size_t)lp * nc;
        for (int j = tid; j < nc; j += kBatchThreads) xo[j] = s.x[j];
        for (int i = tid; i < m; i += kBatchThreads) a.B[(size_t)lp * m + i] = s.Bv[i];
        for (int j = tid; j < n0; j += kBatchThreads) { a.N[(size_t)lp * n0 + j] = s.Nv[j]; a.Ns[(size_t)lp * n0 + j] = s.Ns[j]; }
        if (tid == 0) {
            a.iters[2 * lp] = sh_it[0];
            a.iters[2 * lp + 1] = sh_it[1];
            a.err[lp] = sh_err;
            if (a.trace_len) a.trace_len[lp] = sh_tl;
        }
        __syncthreads();
    }
}

// synthetic batch of configs[3]: LP k: min -c.x, A x <= b, x >= 0 with A ~ U(0,1) m x ns, b ~ U(1,2)*ns/4, c ~ U(0.5,1.5).
// Written directly as ellp's standard form: columns [0,ns) structural, slack of row i at column n0-1-i
// (standard_form.rs:115,129-136), every bound Lower(0), rows in their original order.
__global__ void k_gen_batch(double* __restrict__ A, double* __restrict__ c, double* __restrict__ b, uint8_t* __restrict__ kind,
                            double* __restrict__ lb, double* __restrict__ ub, int nlp, int m, int ns, uint64_t seed, int64_t lp0) {
    const int n0 = ns + m;
    const int64_t per = (int64_t)m * n0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < per * nlp; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = e / per, r = e - k * per;
        const int j = (int)(r / m), i = (int)(r - (int64_t)j * m);
        const uint64_t lpid = (uint64_t)(lp0 + k);
        double v;
        if (j < ns) {
            const uint64_t h = splitmix64((seed + 3 * lpid) * 0x2545f4914f6cdd1dull + (uint64_t)(j * m + i));
            v = (double)(h >> 11) * (1.0 / 9007199254740992.0);
        } else {
            v = (j == n0 - 1 - i) ? 1. : 0.;
        }
        A[e] = v;
    }
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)n0 * nlp; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = e / n0;
        const int j = (int)(e - k * n0);
        const uint64_t lpid = (uint64_t)(lp0 + k);
        double cj = 0.;
        if (j < ns) {
            const uint64_t h = splitmix64((seed + 3 * lpid + 1) * 0x2545f4914f6cdd1dull + (uint64_t)j);
            cj = -(0.5 + (double)(h >> 11) * (1.0 / 9007199254740992.0));
        }
        c[e] = cj;
        kind[e] = ELLP_LOWER;
        lb[e] = 0.;
        ub[e] = 0.;
    }
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)m * nlp; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = e / m;
        const int i = (int)(e - k * m);
        const uint64_t lpid = (uint64_t)(lp0 + k);
        const uint64_t h = splitmix64((seed + 3 * lpid + 2) * 0x2545f4914f6cdd1dull + (uint64_t)i);
        b[e] = (1.0 + (double)(h >> 11) * (1.0 / 9007199254740992.0)) * ((double)ns * 0.25);
    }
}

}  // namespace ellp
