// refactor.cuh -- K4: periodic refactorisation of the basis inverse as a blocked partial-pivot LU whose trailing updates
// run on the fp64 tensor pipe (DMMA, mma.sync.m8n8k4.f64 -- there is no tcgen05 / wgmma kind for f64).
//
// Replaces the per-iteration `A_B.clone().lu()` + solves of the reference (primal_simplex_solver.rs:173-187,295;
// dual_simplex_solver.rs:241-253,294) every `refactor_every` pivots instead of every pivot.  The pivot rule is the
// reference LU's (first max |a| at or below the diagonal), so the pivots are the U_kk that the reference tests against
// EPS ("invalid B, A_B is not invertible", primal :175-179).
//
// Works on the augmented matrix G = [W | X] (ld x 2m, column-major): W starts as A_B, X as I.
//   forward  (per panel of kPanel columns): panel LU with row pivoting -> row swaps on the rest of G -> U12 = L11^-1 G12
//            -> G22 -= L21 * U12 (DMMA) over ALL remaining columns of G (W's trailing block and the whole of X)
//            => W = L\U, X = L^-1 P
//   backward (panels in reverse): X[k,:] = U11^-1 X[k,:] ; X[0:k,:] -= U[0:k,k] * X[k,:] (DMMA)   => X = U^-1 L^-1 P = A_B^-1
// Flops: 2/3 m^3 (LU) + m^3 (forward on X) + m^3 (backward) -- the dense contraction of the hot path.
#pragma once
#include "kernels.cuh"

namespace ellp {

constexpr int kPanel = 32;

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
                                                       const PivotState* st) {
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
    for (int r = 0; r < kPanel; ++r) if (r < nb) col[r] = x[r];
}

// X[k0:k0+nb, :] = U11^-1 X[k0:k0+nb, :] for the columns [c0, ncols) of G: upper triangular back substitution
__global__ void __launch_bounds__(128) k_lu_trsm_upper(double* __restrict__ G, int64_t ld, int ncols, int k0, int nb, int c0,
                                                       const PivotState* st) {
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
    for (int r = 0; r < kPanel; ++r) if (r < nb) col[r] = x[r];
}

// ---- DMMA GEMM: C (M x N) -= A (M x K) * B (K x N), all column-major with the same leading dimension, K <= kPanel ----
// CTA tile 64 x 64, 8 warps as 4 (M) x 2 (N), warp tile 16 x 32 = 2 x 4 mma tiles of m8n8k4 (fp64 tensor pipe).
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int kGemmTile = 64;
constexpr int kGemmLdA = 72;            // padded shared leading dimensions (2-way conflicts are the floor for 8-byte fragments)
constexpr int kGemmLdB = kPanel + 4;

__global__ void __launch_bounds__(256) k_dgemm_sub_dmma(double* __restrict__ C, const double* __restrict__ A, const double* __restrict__ B,
                                                        int64_t ld, int M, int N, int K, const PivotState* st) {
    if (st->err) return;
    __shared__ double sA[kPanel * kGemmLdA];     // sA[k][row]
    __shared__ double sB[kGemmTile * kGemmLdB];  // sB[col][k]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * kGemmTile, n0 = blockIdx.y * kGemmTile;
    // stage A (64 rows x K) and B (K x 64 cols); out-of-range entries are zero
    for (int e = tid; e < kGemmTile * kPanel; e += 256) {
        const int k = e / kGemmTile, r = e - k * kGemmTile;
        sA[k * kGemmLdA + r] = (k < K && m0 + r < M) ? A[(int64_t)k * ld + m0 + r] : 0.;
    }
    for (int e = tid; e < kGemmTile * kPanel; e += 256) {
        const int c = e / kPanel, k = e - c * kPanel;
        sB[c * kGemmLdB + k] = (k < K && n0 + c < N) ? B[(int64_t)(n0 + c) * ld + k] : 0.;
    }
    __syncthreads();
    const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
    const int fr = lane >> 2, fk = lane & 3;  // fragment row / k (A), k / col (B)
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.;
    const int ksteps = (K + 3) / 4;
    for (int ks = 0; ks < ksteps; ++ks) {
        double a[2], b[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = sA[(ks * 4 + fk) * kGemmLdA + wm + i * 8 + fr];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = sB[(wn + j * 8 + fr) * kGemmLdB + ks * 4 + fk];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    // C -= acc : lane holds rows (lane/4), columns 2*(lane%4), +1 of each 8 x 8 tile
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = m0 + wm + i * 8 + fr;
            const int c = n0 + wn + j * 8 + 2 * fk;
            if (r < M) {
                if (c < N) { double* p = C + (int64_t)c * ld + r; *p = *p - acc[i][j][0]; }
                if (c + 1 < N) { double* p = C + (int64_t)(c + 1) * ld + r; *p = *p - acc[i][j][1]; }
            }
        }
}

}  // namespace ellp
