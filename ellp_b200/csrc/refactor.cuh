// refactor.cuh -- K4: periodic refactorisation of the basis inverse as a blocked partial-pivot LU whose trailing updates
// run on the fp64 tensor pipe (DMMA, mma.sync.m8n8k4.f64 -- there is no tcgen05 / wgmma kind for f64).
//
// Replaces the per-iteration `A_B.clone().lu()` + solves of the reference (primal_simplex_solver.rs:173-187,295;
// dual_simplex_solver.rs:241-253,294) every `refactor_every` pivots instead of every pivot.  The pivot rule is the
// reference LU's (first max |a| at or below the diagonal), so the pivots are the U_kk that the reference tests against
// EPS ("invalid B, A_B is not invertible", primal :175-179).
//
// Works on the augmented matrix G = [W | X] (ld x 2m, column-major): W starts as A_B, X as I.
//   forward  (per panel of kPanel columns): panel LU with row pivoting (k_lu_panel_coop: the panel's rows spread over up to
//            148 CTAs in shared memory, one grid barrier per column) -> row swaps on the rest of G -> U12 = L11^-1 G12
//            -> G22 -= L21 * U12 over ALL remaining columns of G (W's trailing block and the whole of X)
//            => W = L\U, X = L^-1 P
//   backward (panels in reverse): X[k,:] = U11^-1 X[k,:] ; X[0:k,:] -= U[0:k,k] * X[k,:]   => X = U^-1 L^-1 P = A_B^-1
// Both rank-kPanel updates run on the fp64 tensor pipe through the SAME kernel as the tableau's deferred row reduction
// (k_blk_flush3, blocked.cuh): the triangular solves also write their block row to a row-major scratch (V), so the update
// is E -= U V with U = the panel's column block (column-major, same ld) -- register-prefetched tiles, bulk-copy ring.
// Flops: 2/3 m^3 (LU) + m^3 (forward on X) + m^3 (backward) -- the dense contraction of the hot path.
#pragma once
#include <cooperative_groups.h>
#include "kernels.cuh"

namespace ellp {

constexpr int kPanel = 64;

// ---- panel factorisation: columns [k0, k0+nb) of G, rows [k0, m); one CTA -----------------------------------------
__global__ void __launch_bounds__(1024) k_lu_panel(double* __restrict__ G, int64_t ld, int m, int k0, int nb, int32_t* __restrict__ piv,
                                                   PivotState* st) {
    if (st->err) return;
    __shared__ double s_v[32];
    __shared__ int s_i[32];
    __shared__ double s_row[kPanel];  // pivot row of the panel (columns j+1..nb-1)
    __shared__ int s_p;
    __shared__ double s_inv;
    const int tid = threadIdx.x;
    const unsigned full = 0xffffffffu;
    for (int j = 0; j < nb; ++j) {
        const int kj = k0 + j;
        double* colj = G + (int64_t)kj * ld;
        // first max |a| at or below the diagonal
        double bv = -1.;
        int bi = 0x7fffffff;
        for (int i = kj + tid; i < m; i += blockDim.x) {
            const double v = fabs(colj[i]);
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
        if (tid < 32) {
            bv = s_v[tid];
            bi = s_i[tid];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const double ov = __shfl_xor_sync(full, bv, off);
                const int oi = __shfl_xor_sync(full, bi, off);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (tid == 0) {
                s_p = (bv >= kEps) ? bi : -1;  // |U_kk| < EPS => not invertible
                if (s_p >= 0) { piv[kj] = bi; s_inv = 1.0 / colj[bi]; }
                else st->err = kErrSingular;
            }
        }
        __syncthreads();
        const int p = s_p;
        if (p < 0) return;
        // swap rows kj <-> p inside the panel, keep the pivot row of the remaining columns in shared memory
        if (tid < nb) {
            double* c = G + (int64_t)(k0 + tid) * ld;
            const double a = c[kj], b = c[p];
            if (p != kj) { c[kj] = b; c[p] = a; }
            s_row[tid] = b;
        }
        __syncthreads();
        // multipliers (reciprocal-scaled like the reference LU) and rank-1 update of the remaining panel columns
        const double inv = s_inv;
        const int rem = nb - j - 1;
        for (int i = kj + 1 + tid; i < m; i += blockDim.x) {
            const double l = colj[i] * inv;
            colj[i] = l;
            for (int c = 0; c < rem; ++c) {
                double* e = G + (int64_t)(kj + 1 + c) * ld + i;
                *e = fma(-l, s_row[j + 1 + c], *e);
            }
        }
        __syncthreads();
    }
}

// apply the panel's row swaps to every column outside the panel: [0, k0) and [k0+nb, ncols)
__global__ void k_lu_swap_rows(double* __restrict__ G, int64_t ld, int ncols, int k0, int nb, const int32_t* __restrict__ piv,
                               const PivotState* st) {
    if (st->err) return;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols - nb) return;
    if (j >= k0) j += nb;
    double* c = G + (int64_t)j * ld;
    for (int t = 0; t < nb; ++t) {
        const int r = k0 + t, p = piv[r];
        if (p != r) { const double a = c[r]; c[r] = c[p]; c[p] = a; }
    }
}

// U12 = L11^-1 G[k0:k0+nb, c0:ncols): unit lower triangular solve, one thread per column, L11 staged in shared memory
__global__ void __launch_bounds__(128) k_lu_trsm_lower(double* __restrict__ G, int64_t ld, int ncols, int k0, int nb, int c0,
                                                       const PivotState* st, double* __restrict__ vout, int64_t ldv) {
    if (st->err) return;
    __shared__ double sL[kPanel * kPanel];
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) { const int c = e / nb, r = e - c * nb; sL[c * kPanel + r] = G[(int64_t)(k0 + c) * ld + k0 + r]; }
    __syncthreads();
    const int j = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    double* col = G + (int64_t)j * ld + k0;
    double x[kPanel];
#pragma unroll
    for (int r = 0; r < kPanel; ++r) x[r] = (r < nb) ? col[r] : 0.;
#pragma unroll
    for (int c = 0; c < kPanel; ++c) {
        if (c < nb) {
            const double xc = x[c];
#pragma unroll
            for (int r = c + 1; r < kPanel; ++r) if (r < nb) x[r] = fma(-sL[c * kPanel + r], xc, x[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < kPanel; ++r)
        if (r < nb) {
            col[r] = x[r];
            if (vout) vout[(int64_t)r * ldv + (j - c0)] = x[r];  // row-major copy of the block row: V operand of the rank-nb update
        }
}

// X[k0:k0+nb, :] = U11^-1 X[k0:k0+nb, :] for the columns [c0, ncols) of G: upper triangular back substitution
__global__ void __launch_bounds__(128) k_lu_trsm_upper(double* __restrict__ G, int64_t ld, int ncols, int k0, int nb, int c0,
                                                       const PivotState* st, double* __restrict__ vout, int64_t ldv) {
    if (st->err) return;
    __shared__ double sU[kPanel * kPanel];
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) { const int c = e / nb, r = e - c * nb; sU[c * kPanel + r] = G[(int64_t)(k0 + c) * ld + k0 + r]; }
    __syncthreads();
    const int j = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    double* col = G + (int64_t)j * ld + k0;
    double x[kPanel];
#pragma unroll
    for (int r = 0; r < kPanel; ++r) x[r] = (r < nb) ? col[r] : 0.;
#pragma unroll
    for (int c = kPanel - 1; c >= 0; --c) {
        if (c < nb) {
            const double xc = x[c] / sU[c * kPanel + c];
            x[c] = xc;
#pragma unroll
            for (int r = 0; r < c; ++r) x[r] = fma(-sU[c * kPanel + r], xc, x[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < kPanel; ++r)
        if (r < nb) {
            col[r] = x[r];
            if (vout) vout[(int64_t)r * ldv + (j - c0)] = x[r];  // row-major copy of the block row: V operand of the rank-nb update
        }
}

// ---- DMMA GEMM: C (M x N) -= A (M x K) * B (K x N), all column-major with the same leading dimension, K <= kPanel ----
// CTA tile 64 x 64, 8 warps as 4 (M) x 2 (N), warp tile 16 x 32 = 2 x 4 mma tiles of m8n8k4 (fp64 tensor pipe).
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ---- cooperative panel factorisation: columns [k0, k0+nb) of G, rows [k0, m) spread over the grid --------------------
// CTA b keeps rows [k0 + b*rpc, k0 + (b+1)*rpc) of the panel in shared memory (row-major, stride kPanel + 1).  Per column
// j: local first-max |a| -> every CTA publishes (value, row, that row's nb entries); the owner of the diagonal row
// publishes it too -> ONE grid barrier -> every CTA derives the same pivot, swaps inside its slice, scales its
// multipliers and applies the rank-1 update to its rows.  Publication buffers are double-buffered by the parity of j
// (a CTA cannot get two columns ahead of a reader).  Same arithmetic as k_lu_panel: multipliers a * (1/pivot), fma update,
// first maximum wins.
struct LuPanelPub {
    double val;    // best |a| of the CTA (-1: none)
    int32_t row;   // its global row
    int32_t pad;
    double entries[kPanel];  // that row of the panel
};
constexpr int kLuPanelSmemMax = 200 * 1024;  // shared memory a CTA of k_lu_panel_coop may use for its slice of the panel
inline size_t lu_panel_pub_doubles(int nctas) { return (size_t)2 * ((size_t)nctas + 1) * (sizeof(LuPanelPub) / sizeof(double)); }

__global__ void __launch_bounds__(256) k_lu_panel_coop(double* __restrict__ G, int64_t ld, int m, int k0, int nb, int rpc,
                                                       int32_t* __restrict__ piv, LuPanelPub* __restrict__ pub, PivotState* st) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double sP[];  // rpc x (kPanel + 1)
    __shared__ double s_v[8];
    __shared__ int s_i[8];
    __shared__ double s_row[kPanel];
    __shared__ int s_p;
    constexpr int SP = kPanel + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int nctas = gridDim.x;
    const int r0 = k0 + blockIdx.x * rpc;               // first global row of this CTA
    const int nr = max(0, min(rpc, m - r0));            // rows held
    const unsigned full = 0xffffffffu;
    if (__ldcg(&st->err)) return;                        // uniform: err was set before the launch
    for (int e = tid; e < nr * nb; e += blockDim.x) {
        const int c = e / nr, i = e - c * nr;
        sP[i * SP + c] = G[(int64_t)(k0 + c) * ld + r0 + i];
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        const int kj = k0 + j;
        LuPanelPub* slot = pub + (size_t)(j & 1) * (nctas + 1);
        // local first max |a| over the rows >= kj
        double bv = -1.;
        int bi = 0x7fffffff;
        for (int i = tid; i < nr; i += blockDim.x) {
            if (r0 + i >= kj) {
                const double v = fabs(sP[i * SP + j]);
                if (v > bv) { bv = v; bi = r0 + i; }
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const double ov = __shfl_xor_sync(full, bv, off);
            const int oi = __shfl_xor_sync(full, bi, off);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < nwarps; ++w)
                if (s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi)) { bv = s_v[w]; bi = s_i[w]; }
            s_p = (bv >= 0.) ? bi : -1;
            slot[blockIdx.x].val = bv;
            slot[blockIdx.x].row = bi;
        }
        __syncthreads();
        {
            const int cand = s_p;
            if (cand >= 0 && tid < nb) slot[blockIdx.x].entries[tid] = sP[(cand - r0) * SP + tid];
            if (kj >= r0 && kj < r0 + nr && tid < nb) slot[nctas].entries[tid] = sP[(kj - r0) * SP + tid];  // the diagonal row
        }
        grid.sync();
        // global winner: largest |a|, smallest row on ties (= first maximum in row order)
        double gv = -1.;
        int gi = 0x7fffffff, gb = -1;
        for (int b = tid; b < nctas; b += blockDim.x) {
            const double v = __ldcg(&slot[b].val);
            const int r = __ldcg(&slot[b].row);
            if (v > gv || (v == gv && r < gi)) { gv = v; gi = r; gb = b; }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const double ov = __shfl_xor_sync(full, gv, off);
            const int oi = __shfl_xor_sync(full, gi, off), ob = __shfl_xor_sync(full, gb, off);
            if (ov > gv || (ov == gv && oi < gi)) { gv = ov; gi = oi; gb = ob; }
        }
        __syncthreads();
        if (lane == 0) { s_v[warp] = gv; s_i[warp] = (gb << 0); }
        __shared__ int s_r[8];
        if (lane == 0) s_r[warp] = gi;
        __syncthreads();
        if (tid == 0) {
            gv = s_v[0]; gi = s_r[0]; gb = s_i[0];
            for (int w = 1; w < nwarps; ++w)
                if (s_v[w] > gv || (s_v[w] == gv && s_r[w] < gi)) { gv = s_v[w]; gi = s_r[w]; gb = s_i[w]; }
            s_p = (gv >= kEps) ? gb : -1;  // |U_kk| < EPS => not invertible (primal :175-179)
            s_i[0] = gi;
        }
        __syncthreads();
        const int wb = s_p, p = s_i[0];
        if (wb < 0) {  // every CTA takes the same exit
            if (blockIdx.x == 0 && tid == 0) st->err = kErrSingular;
            break;
        }
        if (tid < nb) s_row[tid] = __ldcg(&slot[wb].entries[tid]);
        if (blockIdx.x == 0 && tid == 0) piv[kj] = p;
        __syncthreads();
        // swap rows kj <-> p inside the panel
        if (p != kj && tid < nb) {
            if (p >= r0 && p < r0 + nr) sP[(p - r0) * SP + tid] = __ldcg(&slot[nctas].entries[tid]);
            if (kj >= r0 && kj < r0 + nr) sP[(kj - r0) * SP + tid] = s_row[tid];
        }
        __syncthreads();
        // multipliers (reciprocal-scaled like the reference LU) and rank-1 update of the remaining panel columns
        const double inv = 1.0 / s_row[j];
        for (int i = warp; i < nr; i += nwarps) {
            if (r0 + i > kj) {
                const double l = sP[i * SP + j] * inv;
                __syncwarp();
                for (int c = j + 1 + lane; c < nb; c += 32) sP[i * SP + c] = fma(-l, s_row[c], sP[i * SP + c]);
                if (lane == 0) sP[i * SP + j] = l;
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < nr * nb; e += blockDim.x) {
        const int c = e / nr, i = e - c * nr;
        G[(int64_t)(k0 + c) * ld + r0 + i] = sP[i * SP + c];
    }
}

}  // namespace ellp
