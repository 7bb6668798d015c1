// small.cuh -- latency path of the revised engine for netlib-sized LPs (m up to ~100): ONE CTA runs what the general path
// spreads over 8-9 dependent launches per pivot (and 3 launches per column of a refactorisation).
//
//   k_gj_small     Gauss-Jordan inverse of the basis on [A_B | I] held in shared memory: the whole refactorisation
//                  (k_gj_init + m x {k_gj_pivot, k_gj_swap_gather, k_rank1}) in one launch, same pivot rule (first max |a| at or
//                  below the diagonal = the reference LU's U_kk, primal_simplex_solver.rs:173-179) and the same arithmetic
//   k_dual_small   up to `iters` iterations of DualSimplexSolver::solve_with_initial (dual_simplex_solver.rs:188-334) with the
//                  reference's rules: first infeasible basis position leaves (:200-236), rho = e_r^T B^-1 (:248-253),
//                  alpha = A_N^T rho (:255), first minimum ratio enters (:257-289), alpha_q = B^-1 a_q (:294), the updates of
//                  :296-316, the index swap :322-333 and the rank-1 update of B^-1 that replaces the next `lu()` (:241).
//                  Formulas and summation orders are those of the multi-kernel path (k_dual_leaving, k_gather_row, k_gemv_t,
//                  k_select_dual, k_ftran_partial, k_dual_update_vec / _tail, k_rank1), phase by phase, separated by
//                  __syncthreads instead of kernel boundaries; all operands live in the DevLP arrays (L1/L2 resident).
// AFIRO's dual solve: 360 launches -> 6.
#pragma once
#include "kernels.cuh"

namespace ellp {

constexpr int kSmallThreads = 1024;
constexpr int kSmallMaxM = 128;           // k_dual_small: rows (B^-1 is m x m in global memory, one element pass per pivot)
constexpr int kGjSmallSmemMax = 200 * 1024;
inline size_t gj_small_smem_bytes(int64_t ld, int m) { return sizeof(double) * (size_t)ld * 2 * (size_t)m; }

__global__ void __launch_bounds__(kSmallThreads) k_gj_small(DevLP lp, PivotState* st) {
    extern __shared__ __align__(16) double sG[];  // ld x 2m, column-major: [A_B | I]
    __shared__ double s_v[32];
    __shared__ int s_i[32];
    __shared__ double s_dcol[kSmallMaxM + 8];
    __shared__ double s_prow[2 * kSmallMaxM + 8];
    __shared__ int s_p;
    __shared__ double s_alpha;
    const int tid = threadIdx.x;
    const int m = lp.m;
    const int ld = (int)lp.ld, nc = 2 * m;
    const unsigned full = 0xffffffffu;
    if (st->err) return;
    for (int e = tid; e < ld * nc; e += kSmallThreads) {  // k_gj_init
        const int j = e / ld, i = e - j * ld;
        double v;
        if (j < m) v = (i < m) ? lp.A[(int64_t)lp.Bv[j] * lp.ld + i] : 0.;
        else v = (i == j - m) ? 1. : 0.;
        sG[e] = v;
    }
    __syncthreads();
    for (int k = 0; k < m; ++k) {
        // k_gj_pivot: first max |a| of column k at rows >= k
        const double* col = sG + (size_t)k * ld;
        double bv = -1.;
        int bi = 0x7fffffff;
        for (int i = k + tid; i < m; i += kSmallThreads) {
            const double v = fabs(col[i]);
            if (v > bv) { bv = v; bi = i; }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const double ov = __shfl_xor_sync(full, bv, off);
            const int oi = __shfl_xor_sync(full, bi, off);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { s_v[tid >> 5] = bv; s_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kSmallThreads / 32; ++w)
                if (s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi)) { bv = s_v[w]; bi = s_i[w]; }
            if (!(bv >= kEps)) { st->err = kErrSingular; s_p = -1; }  // |U_kk| < EPS: "invalid B, A_B is not invertible"
            else { s_p = bi; s_alpha = col[bi]; }
        }
        __syncthreads();
        const int p = s_p;
        if (p < 0) return;
        const double alpha = s_alpha;
        // pivot column after the row swap k <-> p, and (k_gj_swap_gather) the swap + scaled pivot row of the columns [k, 2m)
        for (int i = tid; i < ld; i += kSmallThreads) {
            const int src = (i == k) ? p : ((i == p) ? k : i);
            s_dcol[i] = (i < m) ? col[src] : 0.;
        }
        __syncthreads();
        for (int j = k + tid; j < nc; j += kSmallThreads) {
            double* c = sG + (size_t)j * ld;
            const double a = c[k], b = c[p];
            if (p != k) { c[k] = b; c[p] = a; }
            s_prow[j - k] = b / alpha;
        }
        __syncthreads();
        // k_rank1 on the columns [k, 2m), pivot row k: E[k,j] = p_j ; E[i,j] = fma(-alpha_i, p_j, E[i,j])
        const int cols = nc - k;
        for (int e = tid; e < cols * m; e += kSmallThreads) {
            const int jj = e / m, i = e - jj * m;
            double* c = sG + (size_t)(k + jj) * ld + i;
            const double pj = s_prow[jj];
            *c = (i == k) ? pj : fma(-s_dcol[i], pj, *c);
        }
        __syncthreads();
    }
    for (int e = tid; e < ld * nc; e += kSmallThreads) lp.G[e] = sG[e];
}

__global__ void __launch_bounds__(kSmallThreads) k_dual_small(DevLP lp, int kc, int KS, int iters, PivotState* st) {
    __shared__ double s_t[32];
    __shared__ int s_p[32];
    __shared__ int s_int[8];     // 0 r_pos, 1 leave_var, 2 new_side, 3 q_pos, 4 q_var, 5 status seen by the block, 6 nan flag
    __shared__ double s_dbl[4];  // 0 delta, 1 theta_d
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    const int m = lp.m, nN = lp.nN;
    const int64_t ld = lp.ld;
    const int len2 = (int)(ld >> 1);
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        if (tid == 0) { s_int[5] = st->status; s_int[6] = 0; }
        __syncthreads();
        if (s_int[5] != kRunning) break;
        // ---- leaving row: first basis position whose variable violates a bound by more than EPS (dual :200-236)
        int best = 0x7fffffff;
        for (int i = tid; i < m; i += kSmallThreads) {
            const int var = lp.Bv[i];
            const double x_i = lp.x[var];
            const int kind = lp.kind[var];
            bool viol = false;
            if (kind == ELLP_LOWER) viol = x_i < lp.lb[var] - kEps;
            else if (kind == ELLP_UPPER) viol = x_i > lp.ub[var] + kEps;
            else if (kind == ELLP_TWOSIDED) viol = (x_i > lp.ub[var] + kEps) || (x_i < lp.lb[var] - kEps);
            if (viol) { best = i; break; }
        }
        best = __reduce_min_sync(full, best);
        if (lane == 0) s_p[warp] = best;
        __syncthreads();
        if (tid == 0) {
            st->do_update = 0;
            for (int w = 1; w < kSmallThreads / 32; ++w) best = min(best, s_p[w]);
            if (best == 0x7fffffff) {
                st->status = ELLP_OPTIMAL;  // :243-246
                s_int[5] = ELLP_OPTIMAL;
            } else {
                const int var = lp.Bv[best];
                const double x_i = lp.x[var];
                const int kind = lp.kind[var];
                double delta;
                int side;
                if (kind == ELLP_LOWER) { delta = x_i - lp.lb[var]; side = ELLP_NB_LOWER; }
                else if (kind == ELLP_UPPER) { delta = x_i - lp.ub[var]; side = ELLP_NB_UPPER; }
                else if (x_i > lp.ub[var] + kEps) { delta = x_i - lp.ub[var]; side = ELLP_NB_UPPER; }
                else { delta = x_i - lp.lb[var]; side = ELLP_NB_LOWER; }
                s_int[0] = best; s_int[1] = var; s_int[2] = side;
                s_dbl[0] = delta;
                st->r_pos = best; st->leave_var = var; st->delta = delta; st->new_side = side;
            }
        }
        __syncthreads();
        if (s_int[5] != kRunning) break;
        const int r = s_int[0];
        const double delta = s_dbl[0];
        // ---- rho = row r of B^-1 (:248-253)
        for (int t = tid; t < m; t += kSmallThreads) lp.rho[t] = lp.Binv[(int64_t)t * ld + r];
        __syncthreads();
        // ---- alpha = A_N^T rho (:255): one warp per nonbasic position, the dot product of k_gemv_t
        for (int j = warp; j < nN; j += kSmallThreads / 32) {
            const double dot = warp_col_dot(lp.A + (int64_t)lp.Nv[j] * ld, lp.rho, len2, lane);
            if (lane == 0) lp.rN[j] = dot;
        }
        __syncthreads();
        // ---- entering: first minimum of d_j / alpha~_j over the eligible nonbasics (:257-289)
        const bool neg = delta < 0.;
        double bt = 0.;
        int bp = 0x7fffffff;
        for (int j = tid; j < nN; j += kSmallThreads) {
            double a = lp.rN[j];
            if (neg) a = -a;
            const int side = lp.Ns[j];
            const bool keep = (side == ELLP_NB_LOWER) ? (a > kEps) : (side == ELLP_NB_UPPER ? (a < -kEps) : true);
            if (keep) {
                const double t = lp.d[lp.Nv[j]] / a;
                if (t != t) s_int[6] = 1;
                if (bp == 0x7fffffff || t < bt) { bt = t; bp = j; }
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const double ot = __shfl_xor_sync(full, bt, off);
            const int op = __shfl_xor_sync(full, bp, off);
            if (op != 0x7fffffff && (bp == 0x7fffffff || ot < bt || (ot == bt && op < bp))) { bt = ot; bp = op; }
        }
        if (lane == 0) { s_t[warp] = bt; s_p[warp] = bp; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kSmallThreads / 32; ++w) {
                const double ot = s_t[w];
                const int op = s_p[w];
                if (op != 0x7fffffff && (bp == 0x7fffffff || ot < bt || (ot == bt && op < bp))) { bt = ot; bp = op; }
            }
            if (s_int[6]) { st->err = kErrNaNDualRatio; st->status = ELLP_INFEASIBLE; s_int[5] = ELLP_INFEASIBLE; }
            else if (bp == 0x7fffffff) { st->status = ELLP_INFEASIBLE; s_int[5] = ELLP_INFEASIBLE; }  // :281-284 dual unbounded
            else {
                s_int[3] = bp;
                s_int[4] = lp.Nv[bp];
                s_dbl[1] = neg ? -bt : bt;  // :286-289
                st->q_pos = bp; st->q_var = s_int[4]; st->q_side = lp.Ns[bp]; st->theta_d = s_dbl[1];
            }
        }
        __syncthreads();
        if (s_int[5] != kRunning) break;
        const int q_pos = s_int[3], q_var = s_int[4];
        const double theta_d = s_dbl[1];
        // ---- alpha_q = B^-1 a_q (:294): split-K partial sums in the order of k_ftran_partial + sum_partials
        {
            const double* aq = lp.A + (int64_t)q_var * ld;
            for (int i = tid; i < (int)ld; i += kSmallThreads) {
                double a = 0.;
                for (int ks = 0; ks < KS; ++ks) {
                    const int k0 = ks * kc, kn = min(kc, m - k0);
                    double pa = 0.;
                    for (int k = 0; k < kn; ++k) pa = fma(lp.Binv[(int64_t)(k0 + k) * ld + i], aq[k0 + k], pa);
                    a += pa;
                }
                lp.dcol[i] = a;
            }
        }
        __syncthreads();
        // ---- updates (:296-316)
        const double alpha_r = lp.dcol[r];
        const double theta_p = delta / alpha_r;  // :306
        for (int t = tid; t < m; t += kSmallThreads) {
            const double a = lp.dcol[t];
            const double rho = lp.rho[t];
            lp.y[t] = lp.y[t] + theta_d * rho;    // :304
            const int var = lp.Bv[t];
            lp.x[var] = lp.x[var] - theta_p * a;  // :310-312
            lp.prow[t] = rho / alpha_r;
        }
        for (int t = tid; t < nN; t += kSmallThreads) {  // :298-300
            const int var = lp.Nv[t];
            lp.d[var] = lp.d[var] - theta_d * lp.rN[t];
        }
        __syncthreads();
        if (tid == 0) {  // :296, :302, :314-333 (k_dual_update_tail)
            const int leave_var = s_int[1];
            lp.d[leave_var] = -theta_d;
            lp.d[q_var] = 0.;
            lp.x[q_var] = lp.x[q_var] + theta_p;
            const int64_t t = st->trace_len;
            if (lp.trace && t < st->trace_cap) {
                ellp_trace_rec rec;
                rec.phase = st->phase_tag;
                rec.iter = (int32_t)st->pivots;
                rec.entering = q_var;
                rec.leaving = leave_var;
                rec.step = theta_p;
                rec.obj = st->obj;
                lp.trace[t] = rec;
            }
            st->trace_len = t + 1;
            st->obj = st->obj + theta_d * delta;  // :316
            lp.Bv[r] = q_var;                     // :322-323
            lp.Nv[q_pos] = leave_var;
            lp.Ns[q_pos] = (uint8_t)s_int[2];
            lp.cB[r] = lp.c[q_var];
            st->alpha_r = alpha_r;
            st->step = theta_p;
            st->do_update = 1;
            st->pivots += 1;
            if (st->pivots >= st->max_iter) st->status = ELLP_MAXITER;  // :191-194 at the next loop head
        }
        // ---- B^-1 <- rank-1 row reduction with pivot row r (k_rank1; replaces the next iteration's lu(), :241)
        for (int e = tid; e < m * m; e += kSmallThreads) {
            const int j = e / m, i = e - j * m;
            double* c = lp.Binv + (int64_t)j * ld + i;
            const double pj = lp.prow[j];
            *c = (i == r) ? pj : fma(-lp.dcol[i], pj, *c);
        }
    }
}

}  // namespace ellp
