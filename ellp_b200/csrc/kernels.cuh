// kernels.cuh -- hand-written sm_100a kernels of the simplex pivot loop (all fp64, device resident).
//
// Kernel <-> reference map (paths relative to the reference repo):
//   k_gemv_t<EPI_PLAIN>         BTRAN  u = A_B^-T c_B      primal_simplex_solver.rs:184-187
//                               pivot row alpha = A_N^T rho dual_simplex_solver.rs:255
//   k_gemv_t<EPI_PRIMAL_PRICE>  r = c_N - A_N^T u + keys    primal_simplex_solver.rs:189, :253-270
//   k_select_primal             Dantzig max_by fold         primal_simplex_solver.rs:271-292
//   k_ftran_partial             B^-1 a_q (split-K GEMV)     primal :295, dual :294
//   k_ratio_primal              ratio fold, step, apply     primal :296-434, :205-232
//   k_dual_leaving              first infeasible basic      dual :200-236
//   k_select_dual               min-ratio entering          dual :257-289
//   k_dual_update_vec/_tail     d, y, x, obj, swap          dual :296-333
//   k_gather_row / k_rank1      replace `A_B.clone().lu()`  primal :173, dual :241
//   k_gj_*                      basis refactorisation (Gauss-Jordan with the LU's pivot rule)
//
// Compiled with -fmad=false: products and sums round separately exactly where the reference's
// Rust does; fused multiply-adds appear only where written as fma().
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "device_types.cuh"

namespace ellp {

struct DevLP {
    int32_t m, n, nN;
    int64_t ld;  // leading dimension of A, Binv, G (m rounded up to a multiple of 4; padding rows are zero)
    const double* A;
    const double* c;
    const double* b;
    const double* lb;
    const double* ub;
    const uint8_t* kind;
    double* x;
    int32_t* Bv;
    int32_t* Nv;
    uint8_t* Ns;
    double* y;
    double* d;
    double* G;     // ld x 2m Gauss-Jordan workspace [A_B | I]
    double* Binv;  // = G + m*ld
    double* cB;    // ld (padding zero)
    double* u;     // ld
    double* rN;    // nN: reduced costs (primal) / pivot row alpha (dual)
    double* key;   // nN
    double* dcol;  // ld: B^-1 a_q (un-negated)
    double* rho;   // ld: row r of B^-1
    double* prow;  // 2m: scaled pivot row consumed by k_rank1
    double* part;  // KS x ld split-K partial sums
    double* lam;   // m
    int32_t* lu_piv;  // m: row pivots of the blocked LU refactorisation
    double* w;        // ld: dual steepest-edge weights ||e_i^T B^-1||^2
    double* npart;    // (m / kNormCols + 1) x ld: per-column-group partial row norms written by k_rank1<.., true>
    ellp_trace_rec* trace;
    // tableau engine (ELLP_ENGINE_TABLEAU): T = B^-1 A lives in the buffer of A (in place), dj = reduced costs
    double* T;     // ld x n, nullptr for the revised engine
    double* dj;    // n
    // column-sharded tableau: this rank stores global columns [col_lo, col_lo + n); x / kind / lb / ub / c are
    // replicated with n_glob entries, Bv holds GLOBAL variable indices, colstat replaces the N list
    int32_t n_glob;
    int32_t col_lo;
    uint8_t* colstat;  // n local entries: ELLP_NB_* or kColBasic; nullptr when not sharded
    // blocked (deferred rank-k) tableau engine, see blocked.cuh: T_current = T - U V over the pending slots
    // single-GPU tableau engine: CONDENSED tableau -- T stores only the nN nonbasic columns, column p belongs to the
    // variable at nonbasic position p (Nv[p]); a pivot overwrites the entering column with the leaving variable's
    // column (the reference swaps A_B / A_N columns the same way, primal :214-217).  Basic columns are unit vectors
    // and are never stored or updated: half the bytes and flops of the full m x n tableau when n = 2m.
    int32_t condensed; // 1: T is ld x nN indexed by position; 0: T is ld x n indexed by (local) variable (sharded engine)
    int32_t nT;        // columns stored in T (nN when condensed, n otherwise)
    int32_t pos_lo;    // peer-sharded condensed tableau (peer.cuh): T / dj / V / key / rN hold the nonbasic POSITIONS
                       // [pos_lo, pos_lo + nT) of the replicated N list; 0 on a single GPU (nT == nN)
    double* U;         // ld x kBlkMax, column j = pivot column of pending pivot j minus e_r (nullptr: rank-1 engine)
    double* V;         // kBlkMax x ldv, row j = scaled pivot row of pending pivot j
    int64_t ldv;
    double* coop;      // 6 * 1024 doubles: per-block partials of the cooperative pivot kernel (blocked.cuh)
    double* xchg;      // small exchange buffers: [0] local max key | [8..8+3) candidate | [16..16+G) gathered max | [32..32+3G) gathered candidates
    // dual simplex on the tableau (dual_blocked.cuh): the basis the tableau was built from, needed to rebuild y from d
    int32_t* Bv0;      // m: variable that was basic in row i when T was built
    double* bscale;    // ld: diagonal of that basis when it was diagonal (generated LPs / identity start); scratch for the y right-hand side
    double* xpart;     // 2 * kXChunks * ld: partial sums of the x_B recomputation at a tableau rebuild
    double* wN;        // nT: primal Devex reference weights of the stored nonbasic positions (ELLP_PRICE_DEVEX)
    double* dpos;      // nN: reduced costs of all nonbasic positions (download: all-gather of dj on the peer engine)
};

constexpr uint8_t kColBasic = 3;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Sum of the KS split-K partials of row `row` in FIXED order (ks ascending: bitwise deterministic), the loads issued 16 at a
// time (the additions are a dependent chain anyway; the exposed latency was one L2 round trip per 4 partials).
__device__ __forceinline__ double sum_partials(const double* __restrict__ part, int64_t ld, int KS, int64_t row) {
    double a = 0.;
    int ks = 0;
    for (; ks + 16 <= KS; ks += 16) {
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = part[(int64_t)(ks + q) * ld + row];
#pragma unroll
        for (int q = 0; q < 16; ++q) a += v[q];
    }
    for (; ks < KS; ++ks) a += part[(int64_t)ks * ld + row];
    return a;
}

__device__ __forceinline__ double2 ld_f64x2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double2 ld_f64x2_stream(const double* p) { return __ldcs(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void st_f64x2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
__device__ __forceinline__ void st_f64x2_stream(double* p, double2 v) { __stcs(reinterpret_cast<double2*>(p), v); }

// ------------------------------------------------------------------------------------------------
// K1/K5 (transpose form): one warp per column, 16-byte coalesced loads, 4 loads in flight per lane,
// warp-shuffle reduction.  len2 = ld/2 double2 elements per column (padding rows are zero in M and v).
// ------------------------------------------------------------------------------------------------
enum { EPI_PLAIN = 0, EPI_PRIMAL_PRICE = 1, EPI_REDCOST = 2 };

__device__ __forceinline__ double warp_col_dot(const double* __restrict__ col, const double* __restrict__ v, int len2, int lane) {
    const double2* c2 = reinterpret_cast<const double2*>(col);
    const double2* v2 = reinterpret_cast<const double2*>(v);
    double s0 = 0., s1 = 0., s2 = 0., s3 = 0.;
    int k = lane;
    for (; k + 96 < len2; k += 128) {
        const double2 a0 = __ldcs(c2 + k), a1 = __ldcs(c2 + k + 32), a2 = __ldcs(c2 + k + 64), a3 = __ldcs(c2 + k + 96);
        const double2 b0 = __ldg(v2 + k), b1 = __ldg(v2 + k + 32), b2 = __ldg(v2 + k + 64), b3 = __ldg(v2 + k + 96);
        s0 = fma(a0.x, b0.x, s0); s0 = fma(a0.y, b0.y, s0);
        s1 = fma(a1.x, b1.x, s1); s1 = fma(a1.y, b1.y, s1);
        s2 = fma(a2.x, b2.x, s2); s2 = fma(a2.y, b2.y, s2);
        s3 = fma(a3.x, b3.x, s3); s3 = fma(a3.y, b3.y, s3);
    }
    for (; k < len2; k += 32) {
        const double2 a0 = __ldcs(c2 + k);
        const double2 b0 = __ldg(v2 + k);
        s0 = fma(a0.x, b0.x, s0); s0 = fma(a0.y, b0.y, s0);
    }
    return warp_sum((s0 + s1) + (s2 + s3));
}

template <int EPI>
__global__ void __launch_bounds__(256) k_gemv_t(const double* __restrict__ M, int64_t ld, const int32_t* __restrict__ cols,
                                                int ncols, const double* __restrict__ v, double* __restrict__ out,
                                                const double* __restrict__ c, const uint8_t* __restrict__ Ns,
                                                double* __restrict__ key, PivotState* st, int clear_update) {
    if (st) {
        if (clear_update && blockIdx.x == 0 && threadIdx.x == 0) { st->do_update = 0; st->do_step = 0; }
        if (st->status != kRunning) return;
    }
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int len2 = (int)(ld >> 1);
    for (int j = blockIdx.x * warps_per_block + (threadIdx.x >> 5); j < ncols; j += gridDim.x * warps_per_block) {
        const int col = cols ? cols[j] : j;
        const double dot = warp_col_dot(M + (int64_t)col * ld, v, len2, lane);
        if (lane == 0) {
            if (EPI == EPI_PLAIN) {
                out[j] = dot;
            } else if (EPI == EPI_REDCOST) {
                out[j] = c[col] - dot;  // d_j = c_j - c_B^T (B^-1 a_j): reduced-cost row of a fresh tableau
            } else {
                // r_j = c_j - a_j^T u (primal :189) and the Dantzig key (primal :258-269); -1 marks "not a candidate"
                const double r = c[col] - dot;
                const int side = Ns[j];
                double k = -1.0;
                if (!(fabs(r) < kEps)) {
                    if (r > 0. && side == ELLP_NB_UPPER) k = r;
                    else if (!(r > 0.) && side == ELLP_NB_LOWER) k = -r;
                    else if (side == ELLP_NB_FREE) k = fabs(r);
                }
                out[j] = r;
                key[j] = k;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Dantzig selection.  tie_rule 0 reproduces Iterator::max_by over N in position order with the
// reference's EPS-tolerant comparator (primal :271-286): the accumulator survives only when it
// compares Greater.  One warp walks the keys 32 at a time; inside a chunk the lanes that would
// change the accumulator are found with a ballot and applied in order, which is exactly the
// sequential fold.  tie_rule 1 is the order-free form (max key, then largest variable index
// within EPS of it).
// ------------------------------------------------------------------------------------------------
constexpr int kScanTile = 4096;                          // (value, tag) pairs per shared-memory tile
constexpr int kScanSmemBytes = 2 * kScanTile * (8 + 4);  // double-buffered
constexpr int kScanThreads = 1024;

// Warps 1..31 stream (value, tag) tiles from global into shared memory while warp 0 folds the previous tile.
__device__ __forceinline__ void scan_load_tile(const double* __restrict__ val, const int32_t* __restrict__ tag, int n, int tile,
                                               double* sval, int* stag, int first_thread, double pad = -1.0) {
    const int base = tile * kScanTile;
    for (int t = threadIdx.x - first_thread; t < kScanTile; t += (int)blockDim.x - first_thread) {
        const int i = base + t;
        sval[t] = (i < n) ? __ldcg(val + i) : pad;
        stag[t] = (i < n) ? __ldcg(tag + i) : 0;
    }
}

// Body shared by k_select_primal (one CTA of kScanThreads threads) and block 0 of the cooperative pivot kernels (any
// block size that is a multiple of 32 and at least 64)
// (blocked.cuh); needs kScanSmemBytes of dynamic shared memory at scan_smem.
__device__ __forceinline__ void select_primal_body(const double* key, const double* rN, const int32_t* Nv, const uint8_t* Ns,
                                                   int nN, int tie_rule, PivotState* st, unsigned char* scan_smem) {
    double* sval[2] = {reinterpret_cast<double*>(scan_smem), reinterpret_cast<double*>(scan_smem) + kScanTile};
    int* stag[2] = {reinterpret_cast<int*>(scan_smem + 2 * kScanTile * 8), reinterpret_cast<int*>(scan_smem + 2 * kScanTile * 8) + kScanTile};
    __shared__ double s_red[32];
    __shared__ int s_redi[32], s_redp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    bool have = false;
    double bk = 0.;
    int bv = 0, bp = -1;
    if (tid == 0) st->lmin_bits = 0x7ff0000000000000ll;  // +inf: reset for this iteration's k_ratio_prep
    bool fast = false;
    if (tie_rule == ELLP_TIES_REFERENCE) {
        // Fast path (exact): let K* = max key.  If exactly one key lies within EPS of K* and none lies in the next EPS band,
        // that element wins the fold whatever the order: it beats every other key strictly, and no other key can displace
        // it or tie with it.  Ties / near-ties fall through to the sequential fold below.
        double kmax = -1.0;
        for (int j = tid; j < nN; j += (int)blockDim.x) kmax = fmax(kmax, __ldcg(key + j));
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) kmax = fmax(kmax, __shfl_xor_sync(full, kmax, off));
        if (lane == 0) s_red[warp] = kmax;
        __syncthreads();
        kmax = s_red[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) kmax = fmax(kmax, s_red[w]);
        __syncthreads();
        int nF = 0, nBand = 0, idxF = 0x7fffffff;
        if (kmax != -1.0) {
            for (int j = tid; j < nN; j += (int)blockDim.x) {
                const double k = __ldcg(key + j);
                if (k == -1.0) continue;
                if (kmax - k < kEps) { ++nF; idxF = min(idxF, j); }
                else if (kmax - k < 2. * kEps) ++nBand;
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            nF += __shfl_xor_sync(full, nF, off);
            nBand += __shfl_xor_sync(full, nBand, off);
            idxF = min(idxF, __shfl_xor_sync(full, idxF, off));
        }
        if (lane == 0) { s_redi[warp] = nF; s_redp[warp] = nBand; s_red[warp] = (double)idxF; }
        __syncthreads();
        nF = 0; nBand = 0; idxF = 0x7fffffff;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { nF += s_redi[w]; nBand += s_redp[w]; idxF = min(idxF, (int)s_red[w]); }
        __syncthreads();
        if (kmax == -1.0) { fast = true; bp = -1; }
        else if (nF == 1 && nBand == 0) { fast = true; bp = idxF; bv = Nv[idxF]; }
    }
    if (tie_rule == ELLP_TIES_REFERENCE && !fast) {
        const int ntiles = (nN + kScanTile - 1) / kScanTile;
        scan_load_tile(key, Nv, nN, 0, sval[0], stag[0], 0);
        __syncthreads();
        for (int t = 0; t < ntiles; ++t) {
            if (warp == 0) {
                const double* sk = sval[t & 1];
                const int* sv = stag[t & 1];
                for (int c = 0; c < kScanTile; c += 32) {
                    const double k = sk[c + lane];
                    const int v = sv[c + lane];
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
                        bp = t * kScanTile + c + f;
                        have = true;
                        rem &= (f == 31) ? 0u : (full << (f + 1));
                    }
                }
            } else if (t + 1 < ntiles) {
                scan_load_tile(key, Nv, nN, t + 1, sval[(t + 1) & 1], stag[(t + 1) & 1], 32);
            }
            __syncthreads();
        }
    } else if (tie_rule != ELLP_TIES_REFERENCE) {
        // order-free rule: max key, then the largest variable index within EPS of it
        double kmax = -1.0;
        for (int j = tid; j < nN; j += (int)blockDim.x) kmax = fmax(kmax, __ldcg(key + j));
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) kmax = fmax(kmax, __shfl_xor_sync(full, kmax, off));
        if (lane == 0) s_red[warp] = kmax;
        __syncthreads();
        kmax = s_red[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) kmax = fmax(kmax, s_red[w]);
        int bestv = -1, bestp = -1;
        if (kmax != -1.0) {
            for (int j = tid; j < nN; j += (int)blockDim.x) {
                const double k = __ldcg(key + j);
                if (k != -1.0 && (kmax - k < kEps)) {
                    const int v = __ldcg(Nv + j);
                    if (v > bestv) { bestv = v; bestp = j; }
                }
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const int ov = __shfl_xor_sync(full, bestv, off), op = __shfl_xor_sync(full, bestp, off);
            if (ov > bestv) { bestv = ov; bestp = op; }
        }
        if (lane == 0) { s_redi[warp] = bestv; s_redp[warp] = bestp; }
        __syncthreads();
        if (tid == 0) {
            bv = -1;
            bp = -1;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
                if (s_redp[w] >= 0 && s_redi[w] > bv) { bv = s_redi[w]; bp = s_redp[w]; }
        }
    }
    if (tid == 0) {
        if (bp < 0) {
            st->status = ELLP_OPTIMAL;  // primal :289-292
        } else {
            st->q_pos = bp;
            st->q_var = bv;
            st->q_side = Ns[bp];
            st->rq = rN[bp];
        }
    }
}

__global__ void __launch_bounds__(kScanThreads) k_select_primal(const double* __restrict__ key, const double* __restrict__ rN,
                                                                const int32_t* __restrict__ Nv, const uint8_t* __restrict__ Ns,
                                                                int nN, int tie_rule, PivotState* st) {
    if (st->status != kRunning) return;
    extern __shared__ __align__(16) unsigned char scan_smem[];
    select_primal_body(key, rN, Nv, Ns, nN, tie_rule, st, scan_smem);
}

// ------------------------------------------------------------------------------------------------
// K5 FTRAN: alpha = B^-1 a_q with B^-1 column-major: each thread owns two rows (one 16-byte load per
// column), a CTA owns 256 rows x kc columns; partial sums go to part[ks][row] and are added in a
// fixed order by the consumer (k_ratio_primal / k_dual_update) => bitwise deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int kFtranMaxKc = 512;

__global__ void __launch_bounds__(128) k_ftran_partial(const double* __restrict__ Binv, int64_t ld, int m,
                                                       const double* __restrict__ A, const PivotState* st,
                                                       double* __restrict__ part, int kc) {
    if (st->status != kRunning) return;
    __shared__ double vs[kFtranMaxKc];
    const int k0 = blockIdx.y * kc;
    const int kn = min(kc, m - k0);
    const double* acol = A + (int64_t)st->q_var * ld;
    for (int t = threadIdx.x; t < kn; t += blockDim.x) vs[t] = acol[k0 + t];
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * 256 + 2 * threadIdx.x;
    if (row >= ld) return;
    double ax = 0., ay = 0.;
    const double* p = Binv + row + (int64_t)k0 * ld;
    int k = 0;
    for (; k + 4 <= kn; k += 4) {
        const double2 e0 = ld_f64x2(p + (int64_t)(k + 0) * ld), e1 = ld_f64x2(p + (int64_t)(k + 1) * ld);
        const double2 e2 = ld_f64x2(p + (int64_t)(k + 2) * ld), e3 = ld_f64x2(p + (int64_t)(k + 3) * ld);
        ax = fma(e0.x, vs[k + 0], ax); ay = fma(e0.y, vs[k + 0], ay);
        ax = fma(e1.x, vs[k + 1], ax); ay = fma(e1.y, vs[k + 1], ay);
        ax = fma(e2.x, vs[k + 2], ax); ay = fma(e2.y, vs[k + 2], ay);
        ax = fma(e3.x, vs[k + 3], ax); ay = fma(e3.y, vs[k + 3], ay);
    }
    for (; k < kn; ++k) {
        const double2 e0 = ld_f64x2(p + (int64_t)k * ld);
        ax = fma(e0.x, vs[k], ax); ay = fma(e0.y, vs[k], ay);
    }
    st_f64x2(part + (int64_t)blockIdx.y * ld + row, make_double2(ax, ay));
}

// generic y = M v on host-supplied device data (kernel-level entry point ellp_b200_gemv_n)
__global__ void __launch_bounds__(128) k_gemv_n_partial(const double* __restrict__ M, int64_t ld, int R, int C,
                                                        const double* __restrict__ v, double* __restrict__ part, int kc) {
    __shared__ double vs[kFtranMaxKc];
    const int k0 = blockIdx.y * kc;
    const int kn = min(kc, C - k0);
    for (int t = threadIdx.x; t < kn; t += blockDim.x) vs[t] = v[k0 + t];
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * 256 + 2 * threadIdx.x;
    if (row >= ld) return;
    double ax = 0., ay = 0.;
    const double* p = M + row + (int64_t)k0 * ld;
    for (int k = 0; k < kn; ++k) {
        const double2 e0 = ld_f64x2(p + (int64_t)k * ld);
        ax = fma(e0.x, vs[k], ax); ay = fma(e0.y, vs[k], ay);
    }
    st_f64x2(part + (int64_t)blockIdx.y * ld + row, make_double2(ax, ay));
}

__global__ void k_sum_partials(const double* __restrict__ part, int64_t ld, int R, int KS, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    double a = 0.;
    for (int ks = 0; ks < KS; ++ks) a += part[(int64_t)ks * ld + i];
    y[i] = a;
}

// ------------------------------------------------------------------------------------------------
// K2 primal: bounded ratio test (primal :305-400), step (:402-434) and pivot application (:205-232)
// in one single-CTA kernel.  tie_rule 0 reproduces the sequential scan with its (lambda, new_basic,
// new_basic_index) state exactly -- including the stale-index behaviour of :379-399.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double primal_ratio(int kind, double lb, double ub, double x_i, double d_i) {
    switch (kind) {
        case ELLP_FREE: return CUDART_INF;
        case ELLP_LOWER:
            if (d_i > 0.) return CUDART_INF;
            return (x_i > lb) ? (lb - x_i) / d_i : 0.;
        case ELLP_UPPER:
            if (d_i > 0.) return (x_i < ub) ? (ub - x_i) / d_i : 0.;
            return CUDART_INF;
        case ELLP_TWOSIDED:
            if (d_i > 0.) return (x_i < ub) ? (ub - x_i) / d_i : 0.;
            return (x_i < lb) ? (lb - x_i) / d_i : 0.;  // primal :359 (sic)
        default: return 0.;                              // Fixed
    }
}

// K2a: pivot column, direction and the per-row ratios (primal :295-367), one row per thread over the whole grid.
// KS > 0: revised engine, alpha = sum of split-K partials (fixed order); KS == 0: tableau engine, alpha = T[:, q]
// (copied out because k_rank1 overwrites that column); KS < 0: alpha already sits in dcol (column-sharded tableau).
// cnt > 0 (blocked tableau engine, KS == 0): the stored tableau is stale by cnt pending pivots; their rank-1
// corrections are applied to the entering column in pivot order, a_i = fma(-U[i,j], V[j,q], a_i) -- the same
// sequence of roundings the rank-1 engine performs on that element.
__global__ void __launch_bounds__(256) k_ratio_prep(DevLP lp, int KS, int cnt, PivotState* st) {
    if (st->status != kRunning) return;
    __shared__ double s_vq[64];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int q_var = st->q_var;
    const bool at_lower = (st->q_side == ELLP_NB_LOWER);
    const int qc = lp.condensed ? st->q_pos : (q_var - lp.col_lo);  // stored column of the entering variable
    if (cnt > 0) {
        if (threadIdx.x < cnt) s_vq[threadIdx.x] = lp.V[(int64_t)threadIdx.x * lp.ldv + qc];
        __syncthreads();
    }
    double lam = kLamSkipped;  // skipped (|d_i| < EPS, :321)
    if (i < lp.m) {
        double a = 0.;
        if (KS > 0) a = sum_partials(lp.part, lp.ld, KS, i);
        else if (KS == 0) {
            a = lp.T[(int64_t)qc * lp.ld + i];
            for (int j = 0; j < cnt; ++j) a = fma(-lp.U[(int64_t)j * lp.ld + i], s_vq[j], a);
        }
        else a = lp.dcol[i];
        lp.dcol[i] = a;
        const double d_i = at_lower ? -a : a;  // :296-300
        if (!(fabs(d_i) < kEps)) {
            const int var = lp.Bv[i];
            lam = primal_ratio(lp.kind[var], lp.lb[var], lp.ub[var], lp.x[var], d_i);
        }
        lp.lam[i] = lam;
    }
    // block minimum of the finite ratios -> one atomicMin per block (non-negative doubles order like their bit patterns).  A
    // negative ratio (quirk Q3) enters as 0: the pick kernel then sees it below L + EPS and either takes it (=> assert!(lambda >=
    // 0.), primal :402) or falls back to the exact sequential fold -- it must never look like "no finite ratio".
    double v = (lam != kLamSkipped && lam < CUDART_INF) ? fmax(lam, 0.) : CUDART_INF;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) v = fmin(v, s[w]);
        if (v < CUDART_INF) atomicMin(&st->lmin_bits, __double_as_longlong(v));
    }
}

// K2b: the ratio fold (primal :305-400), the step decision (:402-434) and the pivot bookkeeping (:205-232), single CTA.
// tie_rule 0 reproduces the sequential scan with its (lambda, new_basic, new_basic_index) state exactly -- including the
// stale-index behaviour of :379-399 -- with an exact shortcut when the minimum is isolated.
__device__ __forceinline__ void ratio_commit(const DevLP& lp, PivotState* st, int nb, double lambda, bool at_lower, int q_var);

__device__ __forceinline__ void ratio_pick_body(const DevLP& lp, int tie_rule, PivotState* st, unsigned char* scan_smem) {
    __shared__ double s_lambda;
    __shared__ int s_nb;
    const int tid = threadIdx.x;
    const int m = lp.m;
    const int q_var = st->q_var;
    const bool at_lower = (st->q_side == ELLP_NB_LOWER);
    __syncthreads();
    {
        double* sval[2] = {reinterpret_cast<double*>(scan_smem), reinterpret_cast<double*>(scan_smem) + kScanTile};
        int* stag[2] = {reinterpret_cast<int*>(scan_smem + 2 * kScanTile * 8), reinterpret_cast<int*>(scan_smem + 2 * kScanTile * 8) + kScanTile};
        __shared__ double s_red[32];
        __shared__ int s_redi[32], s_redp[32];
        const unsigned full = 0xffffffffu;
        const int lane = tid & 31, warp = tid >> 5;
        double lambda;
        {
            const int kq = lp.kind[q_var];  // :305-311
            lambda = (kq == ELLP_TWOSIDED) ? (lp.ub[q_var] - lp.lb[q_var]) : (kq == ELLP_FIXED ? 0. : CUDART_INF);
        }
        int nb = -1;
        bool fast = false;
        if (tie_rule == ELLP_TIES_REFERENCE) {
            // Fast path (exact): L = min(lambda0, min_i lambda_i).  If exactly one of {lambda0, lambda_i} lies below L + EPS and
            // none lies in [L + EPS, L + 2 EPS), the fold ends on that element whatever happened before it: when it is reached
            // the running lambda exceeds it by more than EPS (strict branch), and afterwards nothing is within EPS of it.
            const double L = fmin(__longlong_as_double(st->lmin_bits), lambda);
            int nF = 0, nBand = 0, idxF = 0x7fffffff;
            if (L < CUDART_INF) {
                for (int i = tid; i < m; i += blockDim.x) {
                    const double l = lp.lam[i];
                    if (l == kLamSkipped) continue;
                    if (l < L + kEps) { ++nF; idxF = min(idxF, i); }
                    else if (l < L + 2. * kEps) ++nBand;
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                nF += __shfl_xor_sync(full, nF, off);
                nBand += __shfl_xor_sync(full, nBand, off);
                idxF = min(idxF, __shfl_xor_sync(full, idxF, off));
            }
            if (lane == 0) { s_redi[warp] = nF; s_redp[warp] = nBand; s_red[warp] = (double)idxF; }
            __syncthreads();
            nF = 0; nBand = 0; idxF = 0x7fffffff;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { nF += s_redi[w]; nBand += s_redp[w]; idxF = min(idxF, (int)s_red[w]); }
            __syncthreads();
            const int f0 = (lambda < L + kEps) ? 1 : 0;
            const int band0 = (!f0 && lambda < L + 2. * kEps) ? 1 : 0;
            if (!(L < CUDART_INF)) { fast = true; }                               // nothing finite: lambda stays +inf (unbounded)
            else if (nF + f0 == 1 && nBand + band0 == 0) {
                fast = true;
                if (!f0) { nb = idxF; lambda = lp.lam[idxF]; }                    // else: the entering variable's own bound flip
            }
            if (fast && tid == 0) { s_lambda = lambda; s_nb = nb; }
        }
        if (tie_rule == ELLP_TIES_REFERENCE && !fast) {
            bool have_nbi = false;
            int nbi = 0;
            const int ntiles = (m + kScanTile - 1) / kScanTile;
            scan_load_tile(lp.lam, lp.Bv, m, 0, sval[0], stag[0], 0, kLamSkipped);
            __syncthreads();
            for (int t = 0; t < ntiles; ++t) {
                if (warp == 0) {
                    const double* sl = sval[t & 1];
                    const int* sv = stag[t & 1];
                    for (int c = 0; c < kScanTile; c += 32) {
                        const double l = sl[c + lane];
                        const int v = sv[c + lane];
                        // lambda_i = +inf never changes the state (:379, :387), so it is not a candidate here
                        const bool cand = (l != kLamSkipped) && (l < CUDART_INF);
                        unsigned rem = __ballot_sync(full, cand);
                        while (rem) {
                            int eff = 0;  // 1 strict (:379), 2 tie accepted (:387-399)
                            if (cand && ((rem >> lane) & 1u)) {
                                if (l < lambda - kEps) eff = 1;
                                else if (fabs(l - lambda) < kEps && (!have_nbi || v < nbi)) eff = 2;
                            }
                            const unsigned msk = __ballot_sync(full, eff != 0);
                            if (!msk) break;
                            const int f = __ffs(msk) - 1;
                            const int kind_f = __shfl_sync(full, eff, f);
                            lambda = __shfl_sync(full, l, f);
                            nb = t * kScanTile + c + f;
                            if (kind_f == 2) { have_nbi = true; nbi = __shfl_sync(full, v, f); }
                            rem &= (f == 31) ? 0u : (full << (f + 1));
                        }
                    }
                } else if (t + 1 < ntiles) {
                    scan_load_tile(lp.lam, lp.Bv, m, t + 1, sval[(t + 1) & 1], stag[(t + 1) & 1], 32, kLamSkipped);
                }
                __syncthreads();
            }
            if (tid == 0) { s_lambda = lambda; s_nb = nb; }
        } else if (tie_rule != ELLP_TIES_REFERENCE) {
            double lmin = CUDART_INF;
            for (int i = tid; i < m; i += blockDim.x) {
                const double l = lp.lam[i];
                if (l != kLamSkipped && l < lmin) lmin = l;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) lmin = fmin(lmin, __shfl_xor_sync(full, lmin, off));
            if (lane == 0) s_red[warp] = lmin;
            __syncthreads();
            lmin = s_red[0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) lmin = fmin(lmin, s_red[w]);
            int bestv = 0x7fffffff, bestp = -1;
            const bool basic_wins = (lmin < lambda + kEps && lmin < CUDART_INF);
            if (basic_wins) {
                for (int i = tid; i < m; i += blockDim.x) {
                    const double l = lp.lam[i];
                    if (l != kLamSkipped && (l - lmin < kEps)) {
                        const int v = lp.Bv[i];
                        if (v < bestv) { bestv = v; bestp = i; }
                    }
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const int ov = __shfl_xor_sync(full, bestv, off), op = __shfl_xor_sync(full, bestp, off);
                if (ov < bestv) { bestv = ov; bestp = op; }
            }
            if (lane == 0) { s_redi[warp] = bestv; s_redp[warp] = bestp; }
            __syncthreads();
            if (tid == 0) {
                bestv = 0x7fffffff;
                bestp = -1;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
                    if (s_redp[w] >= 0 && s_redi[w] < bestv) { bestv = s_redi[w]; bestp = s_redp[w]; }
                if (basic_wins && bestp >= 0) { nb = bestp; lambda = lp.lam[nb]; }
                s_lambda = lambda;
                s_nb = nb;
            }
        }
    }
    __syncthreads();
    const double lambda = s_lambda;
    const int nb = s_nb;
    if (tid == 0) ratio_commit(lp, st, nb, lambda, at_lower, q_var);
}

// The step decision (:402-434) and the pivot bookkeeping (:205-232) for the outcome (nb, lambda) of the ratio fold:
// nb = leaving basis position or -1 for a bound flip of the entering variable.  One thread.
__device__ __forceinline__ void ratio_commit(const DevLP& lp, PivotState* st, int nb, double lambda, bool at_lower, int q_var) {
    if (!(lambda >= 0.)) {  // :402 assert!(lambda >= 0.)
        st->err = kErrLambdaNegative; st->status = ELLP_UNBOUNDED; st->do_update = 0; st->do_step = 0;
        return;
    }
    if (isinf(lambda)) {  // :404-406
        st->status = ELLP_UNBOUNDED; st->do_update = 0; st->do_step = 0;
        return;
    }
    {
        st->do_step = (lambda > 0.) ? 1 : 0;  // :408-417 is carried out by k_step_gather (x moves along d by lambda)
        const int q_pos = st->q_pos;
        const int64_t t = st->trace_len;
        int leave_var = -1;
        if (nb >= 0) {  // :208-221
            const double a = lp.dcol[nb];
            const double d_nb = at_lower ? -a : a;
            const uint8_t side_new = (d_nb > 0.) ? ELLP_NB_UPPER : ELLP_NB_LOWER;
            leave_var = lp.Bv[nb];
            lp.Bv[nb] = q_var;
            if (lp.colstat) {  // sharded: the owners of the two columns record the swap
                if (q_var >= lp.col_lo && q_var < lp.col_lo + lp.n) lp.colstat[q_var - lp.col_lo] = kColBasic;
                if (leave_var >= lp.col_lo && leave_var < lp.col_lo + lp.n) lp.colstat[leave_var - lp.col_lo] = side_new;
            } else {
                lp.Nv[q_pos] = leave_var;
                lp.Ns[q_pos] = side_new;
            }
            lp.cB[nb] = lp.c[q_var];
            st->r_pos = nb;
            st->leave_var = leave_var;
            st->alpha_r = a;
            st->do_update = 1;
        } else {  // :223-231 bound flip
            const int side = st->q_side;
            uint8_t flipped = ELLP_NB_FREE;
            if (side == ELLP_NB_LOWER) flipped = ELLP_NB_UPPER;
            else if (side == ELLP_NB_UPPER) flipped = ELLP_NB_LOWER;
            else { st->err = kErrFlipFree; st->status = ELLP_UNBOUNDED; }
            if (flipped != ELLP_NB_FREE) {
                if (lp.colstat) { if (q_var >= lp.col_lo && q_var < lp.col_lo + lp.n) lp.colstat[q_var - lp.col_lo] = flipped; }
                else lp.Ns[q_pos] = flipped;
            }
            st->r_pos = -1;
            st->do_update = 0;
        }
        if (lp.trace && t < st->trace_cap) {
            ellp_trace_rec rec;
            rec.phase = st->phase_tag;
            rec.iter = (int32_t)st->pivots;
            rec.entering = q_var;
            rec.leaving = leave_var;
            rec.step = lambda;
            rec.obj = st->obj;
            lp.trace[t] = rec;
        }
        st->trace_len = t + 1;
        st->step = lambda;
        st->obj = st->obj + st->rq * (at_lower ? lambda : -lambda);
        st->pivots += 1;
        if (st->status == kRunning && st->pivots >= st->max_iter) st->status = ELLP_MAXITER;  // :163-166 at the next loop head
    }
}

// ELLP_RATIO_HARRIS on the primal side (no reference counterpart: ellp's ratio test is the textbook fold of :305-400): Harris'
// two-pass test with bound flipping.  Pass 1: theta_max = min over the blocking rows of (distance to the bound + tol) / |d_i|, and the
// entering variable's own range.  Pass 2: among the rows whose exact ratio is <= theta_max take the LARGEST |d_i| (ties: smallest
// position) -- a pivot element that is not needlessly small; the step is max(ratio, 0).  If the entering variable's own range is not
// larger than that step it flips bounds instead (no basis change).  Ends in the same ratio_commit as the reference rule.
__global__ void __launch_bounds__(1024) k_ratio_pick_harris(DevLP lp, double tol, PivotState* st) {
    if (st->status != kRunning) return;
    __shared__ double s_v[32];
    __shared__ int s_i[32];
    __shared__ double s_theta;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    const int m = lp.m;
    const int q_var = st->q_var;
    const bool at_lower = (st->q_side == ELLP_NB_LOWER);
    const int kq = lp.kind[q_var];
    const double lambda0 = (kq == ELLP_TWOSIDED) ? (lp.ub[q_var] - lp.lb[q_var]) : (kq == ELLP_FIXED ? 0. : CUDART_INF);
    // exact and relaxed ratio of row i (CUDART_INF when the row does not block)
    auto ratios = [&](int i, double* exact, double* relaxed, double* mag) {
        const double a = lp.dcol[i];
        const double d_i = at_lower ? -a : a;
        *exact = CUDART_INF; *relaxed = CUDART_INF; *mag = fabs(d_i);
        if (fabs(d_i) < kEps) return;
        const int var = lp.Bv[i];
        const int kind = lp.kind[var];
        const double x_i = lp.x[var];
        double slack = CUDART_INF;
        if (d_i < 0.) { if (kind == ELLP_LOWER || kind == ELLP_TWOSIDED || kind == ELLP_FIXED) slack = x_i - lp.lb[var]; }
        else { if (kind == ELLP_UPPER || kind == ELLP_TWOSIDED) slack = lp.ub[var] - x_i; else if (kind == ELLP_FIXED) slack = lp.lb[var] - x_i; }
        if (slack == CUDART_INF) return;
        *exact = fmax(slack, 0.) / fabs(d_i);
        *relaxed = (fmax(slack, 0.) + tol) / fabs(d_i);
    };
    double tmin = lambda0;
    for (int i = tid; i < m; i += blockDim.x) {
        double e, r, g;
        ratios(i, &e, &r, &g);
        tmin = fmin(tmin, r);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) tmin = fmin(tmin, __shfl_xor_sync(full, tmin, off));
    if (lane == 0) s_v[warp] = tmin;
    __syncthreads();
    if (tid == 0) { double t = s_v[0]; for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = fmin(t, s_v[w]); s_theta = t; }
    __syncthreads();
    const double theta = s_theta;
    double bg = -1.;
    int bi = 0x7fffffff;
    for (int i = tid; i < m; i += blockDim.x) {
        double e, r, g;
        ratios(i, &e, &r, &g);
        if (e <= theta && (g > bg)) { bg = g; bi = i; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double og = __shfl_xor_sync(full, bg, off);
        const int oi = __shfl_xor_sync(full, bi, off);
        if (og > bg || (og == bg && oi < bi)) { bg = og; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { s_v[warp] = bg; s_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        bg = -1.; bi = 0x7fffffff;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) if (s_v[w] > bg || (s_v[w] == bg && s_i[w] < bi)) { bg = s_v[w]; bi = s_i[w]; }
        int nb = -1;
        double lambda = lambda0;  // +inf when nothing blocks: ratio_commit reports Unbounded
        if (bi != 0x7fffffff) {
            double e, r, g;
            ratios(bi, &e, &r, &g);
            if (!(lambda0 <= e)) { nb = bi; lambda = e; }
        }
        ratio_commit(lp, st, nb, lambda, at_lower, q_var);
    }
}

__global__ void __launch_bounds__(1024) k_ratio_pick(DevLP lp, int tie_rule, PivotState* st) {
    if (st->status != kRunning) return;
    extern __shared__ __align__(16) unsigned char scan_smem[];
    ratio_pick_body(lp, tie_rule, st, scan_smem);
}

// ------------------------------------------------------------------------------------------------
// pivot row of E: out[j] = E[r, j] (mode 0, dual rho) or E[r, j] / alpha_r (mode 1, feeds k_rank1)
// ------------------------------------------------------------------------------------------------
__global__ void k_gather_row(const double* __restrict__ E, int64_t ld, int C, const PivotState* st, double* __restrict__ out,
                             int mode) {
    if (mode == 0) { if (st->status != kRunning) return; }
    else if (!st->do_update) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= C) return;
    const double e = E[(int64_t)j * ld + st->r_pos];
    out[j] = (mode == 0) ? e : e / st->alpha_r;
}

// K2c: the primal step x_B += lambda d, x_q +-= lambda (primal :408-417) fused with the scaled pivot-row gather that feeds
// k_rank1.  Runs after the bookkeeping of k_ratio_pick, so basis position r already holds the entering variable: the
// leaving variable's value is updated through PivotState::leave_var.
__global__ void k_step_gather(DevLP lp, const double* __restrict__ E, int C, PivotState* st) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (st->do_step) {
        const double lambda = st->step;
        const bool at_lower = (st->q_side == ELLP_NB_LOWER);
        if (t < lp.m) {
            const double a = lp.dcol[t];
            const double d_i = at_lower ? -a : a;
            const int var = (st->do_update && t == st->r_pos) ? st->leave_var : lp.Bv[t];
            lp.x[var] = lp.x[var] + lambda * d_i;
        }
        if (t == 0) {
            const int q = st->q_var;
            lp.x[q] = at_lower ? lp.x[q] + lambda : lp.x[q] - lambda;
        }
    }
    if (st->do_update && t < C) lp.prow[t] = E[(int64_t)t * lp.ld + st->r_pos] / st->alpha_r;
}

// K2c for the condensed tableau: as k_step_gather, plus the column bookkeeping of the condensed form.  The entering
// column (position q_pos) is handed over to the leaving variable, whose current column is e_r: the stored column is
// overwritten with e_r and its pivot-row entry is 1/alpha_r, so the generic update T[:,q] -= (d - e_r) p_q (k_rank1)
// leaves -d_i/alpha_r in rows i != r and 1/alpha_r in row r; the reduced cost of the leaving variable, 0 while basic,
// becomes -d_q/alpha_r the same way.
__global__ void k_step_gather_cond(DevLP lp, PivotState* st) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (st->do_step) {
        const double lambda = st->step;
        const bool at_lower = (st->q_side == ELLP_NB_LOWER);
        if (t < lp.m) {
            const double a = lp.dcol[t];
            const double d_i = at_lower ? -a : a;
            const int var = (st->do_update && t == st->r_pos) ? st->leave_var : lp.Bv[t];
            lp.x[var] = lp.x[var] + lambda * d_i;
        }
        if (t == 0) {
            const int q = st->q_var;
            lp.x[q] = at_lower ? lp.x[q] + lambda : lp.x[q] - lambda;
        }
    }
    if (!st->do_update) return;
    const int r = st->r_pos, qp = st->q_pos;
    if (t < lp.nT) {
        if (t == qp) { lp.prow[t] = 1.0 / st->alpha_r; lp.dj[t] = 0.; }
        else lp.prow[t] = lp.T[(int64_t)t * lp.ld + r] / st->alpha_r;
    }
    if (t < lp.ld) lp.T[(int64_t)qp * lp.ld + t] = (t == r) ? 1. : 0.;
}

// condensed tableau set-up: T[:, p] = F[:, Nv[p]] (F = full tableau B^-1 A built in the buffer of A)
__global__ void k_gather_cols(const double* __restrict__ F, int64_t ld, const int32_t* __restrict__ Nv, int nN, double* __restrict__ T) {
    const int p = blockIdx.y;
    if (p >= nN) return;
    const double* src = F + (int64_t)Nv[p] * ld;
    double* dst = T + (int64_t)p * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_redcost_pos(const double* __restrict__ c, const int32_t* __restrict__ Nv, int nN, double* __restrict__ dj) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nN) dj[p] = c[Nv[p]] - dj[p];  // d_p = c_p - c_B^T (B^-1 a_p); dj holds the dot product on entry
}

// ------------------------------------------------------------------------------------------------
// K3: rank-1 row reduction  E[r,j] = p_j ; E[i,j] = fma(-alpha_i, p_j, E[i,j])   (p_j = E_old[r,j]/alpha_r)
// HBM-bound read-modify-write stream: 16 B of traffic per element, 2 flops.
//   - a thread owns two consecutive rows (one 16-byte access per column), a CTA 2*blockDim rows;
//     its -alpha pair lives in registers for the whole CTA lifetime
//   - kColsInFlight columns are loaded before any is stored => kColsInFlight x 16 B in flight/thread
//   - STREAM=true uses evict-first loads/stores (tableau much larger than L2); false keeps the
//     default policy so a basis inverse that fits the 126 MB L2 stays resident between pivots
// ------------------------------------------------------------------------------------------------
constexpr int kRank1Threads = 256;
constexpr int kColsInFlight = 8;

constexpr int kNormCols = 128;  // columns per CTA when the row norms ride along (keeps the partial-norm scratch at m^2/128 doubles)

template <bool STREAM, bool NORMS>
__global__ void __launch_bounds__(kRank1Threads) k_rank1(double* __restrict__ E, int64_t ld, int R, int C,
                                                         const double* __restrict__ alpha, const double* __restrict__ prow,
                                                         const PivotState* st, int r_fixed, int cols_per_cta,
                                                         double* __restrict__ dj, double* __restrict__ npart) {
    int r = r_fixed;
    if (st) {
        if (!st->do_update) return;
        r = st->r_pos;
    }
    const int c0 = blockIdx.y * cols_per_cta;
    const int c1 = min(C, c0 + cols_per_cta);
    // tableau engine: the reduced-cost row is one more row of the same rank-1 update,
    // d_j -= d_q * p_j; the first row-block of every column group carries it (C extra doubles in total)
    if (dj != nullptr && blockIdx.x == 0) {
        const double nrq = -st->rq;
        for (int j = c0 + threadIdx.x; j < c1; j += kRank1Threads) dj[j] = fma(nrq, __ldg(prow + j), dj[j]);
    }
    const int64_t row = ((int64_t)blockIdx.x * kRank1Threads + threadIdx.x) * 2;
    if (row >= R) return;
    double2 na = ld_f64x2(alpha + row);
    na.x = -na.x;
    na.y = (row + 1 < R) ? -na.y : 0.;  // a padding row (ld > R) is never modified: fma(0, p, e) == e for finite p
    const bool r0 = (row == r), r1 = (row + 1 == r);
    double* base = E + row;
    double n0 = 0., n1 = 0.;  // NORMS: sum of squares of the UPDATED entries of this thread's two rows over the CTA's columns
    int j = c0;
    for (; j + kColsInFlight <= c1; j += kColsInFlight) {
        double2 e[kColsInFlight];
        double p[kColsInFlight];
#pragma unroll
        for (int u = 0; u < kColsInFlight; ++u)
            e[u] = STREAM ? ld_f64x2_stream(base + (int64_t)(j + u) * ld) : ld_f64x2(base + (int64_t)(j + u) * ld);
#pragma unroll
        for (int u = 0; u < kColsInFlight; ++u) p[u] = __ldg(prow + j + u);
#pragma unroll
        for (int u = 0; u < kColsInFlight; ++u) {
            e[u].x = r0 ? p[u] : fma(na.x, p[u], e[u].x);
            e[u].y = r1 ? p[u] : fma(na.y, p[u], e[u].y);
            if (NORMS) { n0 = fma(e[u].x, e[u].x, n0); n1 = fma(e[u].y, e[u].y, n1); }
            if (STREAM) st_f64x2_stream(base + (int64_t)(j + u) * ld, e[u]);
            else st_f64x2(base + (int64_t)(j + u) * ld, e[u]);
        }
    }
    for (; j < c1; ++j) {
        double2 e = ld_f64x2(base + (int64_t)j * ld);
        const double p = __ldg(prow + j);
        e.x = r0 ? p : fma(na.x, p, e.x);
        e.y = r1 ? p : fma(na.y, p, e.y);
        if (NORMS) { n0 = fma(e.x, e.x, n0); n1 = fma(e.y, e.y, n1); }
        st_f64x2(base + (int64_t)j * ld, e);
    }
    if (NORMS) st_f64x2(npart + (int64_t)blockIdx.y * ld + row, make_double2(n0, (row + 1 < R) ? n1 : 0.));
}

// scalar variant for matrices whose leading dimension is odd (kernel-level entry point only)
__global__ void k_rank1_scalar(double* __restrict__ E, int64_t ld, int R, int C, const double* __restrict__ alpha,
                               const double* __restrict__ prow, int r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const double na = -alpha[i];
    for (int j = blockIdx.y; j < C; j += gridDim.y) {
        const double p = prow[j];
        double* e = E + (int64_t)j * ld + i;
        *e = (i == r) ? p : fma(na, p, *e);
    }
}

// row norms of E without updating it (initial dual steepest-edge weights / after a refactorisation)
__global__ void __launch_bounds__(kRank1Threads) k_rownorms_partial(const double* __restrict__ E, int64_t ld, int R, int C,
                                                                    int cols_per_cta, double* __restrict__ npart) {
    const int64_t row = ((int64_t)blockIdx.x * kRank1Threads + threadIdx.x) * 2;
    if (row >= R) return;
    const int c0 = blockIdx.y * cols_per_cta, c1 = min(C, c0 + cols_per_cta);
    double n0 = 0., n1 = 0.;
    for (int j = c0; j < c1; ++j) {
        const double2 e = ld_f64x2(E + (int64_t)j * ld + row);
        n0 = fma(e.x, e.x, n0);
        n1 = fma(e.y, e.y, n1);
    }
    st_f64x2(npart + (int64_t)blockIdx.y * ld + row, make_double2(n0, (row + 1 < R) ? n1 : 0.));
}

// w_i = sum over column groups in a fixed order => the weights are bitwise reproducible
__global__ void k_sum_norms(const double* __restrict__ npart, int64_t ld, int m, int groups, double* __restrict__ w, const PivotState* st,
                            int need_update) {
    if (need_update && !st->do_update) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double a = 0.;
    for (int g = 0; g < groups; ++g) a += npart[(int64_t)g * ld + i];
    w[i] = a;
}

// Scratch of the multi-block selection kernels: per-block partial results + the ticket of the last-block pattern (the block
// whose arrival completes the grid finishes the selection, so no second launch and no single-CTA scan is needed).
constexpr int kSelMaxBlocks = 128;
struct SelScratch {
    unsigned int ticket;
    int flag;                      // NaN seen (dual ratio)
    double val[kSelMaxBlocks];
    int idx[kSelMaxBlocks];
};
__device__ __forceinline__ bool sel_arrive_last(SelScratch* sc, int* s_last) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&sc->ticket, 1u);
        *s_last = (t == gridDim.x - 1);
        if (*s_last) { sc->ticket = 0u; __threadfence(); }
    }
    __syncthreads();
    return *s_last != 0;
}

// ------------------------------------------------------------------------------------------------
// Dual simplex kernels
// ------------------------------------------------------------------------------------------------
// dual :200-236 -- first basis position (in B order) whose variable violates its bound by more than EPS
__global__ void __launch_bounds__(256) k_dual_leaving(DevLP lp, PivotState* st, SelScratch* sc) {
    if (blockIdx.x == 0 && threadIdx.x == 0) st->do_update = 0;
    if (st->status != kRunning) return;
    __shared__ int s_min[8];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    int best = 0x7fffffff;
    for (int i = blockIdx.x * blockDim.x + tid; i < lp.m; i += gridDim.x * blockDim.x) {
        const int var = lp.Bv[i];
        const double x_i = lp.x[var];
        const int kind = lp.kind[var];
        bool viol = false;
        if (kind == ELLP_LOWER) viol = x_i < lp.lb[var] - kEps;
        else if (kind == ELLP_UPPER) viol = x_i > lp.ub[var] + kEps;
        else if (kind == ELLP_TWOSIDED) viol = (x_i > lp.ub[var] + kEps) || (x_i < lp.lb[var] - kEps);
        if (viol) { best = i; break; }  // i increases monotonically per thread
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
    if ((tid & 31) == 0) s_min[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = min(best, s_min[w]);
        sc->idx[blockIdx.x] = best;
    }
    if (!sel_arrive_last(sc, &s_last)) return;
    if (tid == 0) {  // the first violated position over all blocks: exact min, order-free
        best = 0x7fffffff;
        for (int b = 0; b < (int)gridDim.x; ++b) best = min(best, __ldcg(&sc->idx[b]));
        if (best == 0x7fffffff) {
            st->status = ELLP_OPTIMAL;  // :243-246
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
            st->r_pos = best;
            st->leave_var = var;
            st->delta = delta;
            st->new_side = side;
        }
    }
}

// dual :257-289 -- entering = first minimum of d_j / alpha~_j over eligible nonbasics (exact compare,
// first in N-position order on equality) == lexicographic min of (ratio, position): order-free.
__global__ void __launch_bounds__(256) k_select_dual(DevLP lp, PivotState* st, SelScratch* sc) {
    if (st->status != kRunning) return;
    __shared__ double s_t[8];
    __shared__ int s_p[8];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const bool neg = st->delta < 0.;
    double bt = 0.;
    int bp = 0x7fffffff;
    bool nan_seen = false;
    for (int j = blockIdx.x * blockDim.x + tid; j < lp.nN; j += gridDim.x * blockDim.x) {
        double a = lp.rN[j];
        if (neg) a = -a;
        const int side = lp.Ns[j];
        const bool keep = (side == ELLP_NB_LOWER) ? (a > kEps) : (side == ELLP_NB_UPPER ? (a < -kEps) : true);
        if (keep) {
            const double t = lp.d[lp.Nv[j]] / a;
            if (t != t) nan_seen = true;
            if (bp == 0x7fffffff || t < bt) { bt = t; bp = j; }  // j increases per thread: keeps the first minimum
        }
    }
    if (nan_seen) atomicOr(&sc->flag, 1);
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double ot = __shfl_xor_sync(full, bt, off);
        const int op = __shfl_xor_sync(full, bp, off);
        if (op != 0x7fffffff && (bp == 0x7fffffff || ot < bt || (ot == bt && op < bp))) { bt = ot; bp = op; }
    }
    if ((tid & 31) == 0) { s_t[tid >> 5] = bt; s_p[tid >> 5] = bp; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            const double ot = s_t[w];
            const int op = s_p[w];
            if (op != 0x7fffffff && (bp == 0x7fffffff || ot < bt || (ot == bt && op < bp))) { bt = ot; bp = op; }
        }
        sc->val[blockIdx.x] = bt;
        sc->idx[blockIdx.x] = bp;
    }
    if (!sel_arrive_last(sc, &s_last)) return;
    if (tid == 0) {  // lexicographic min of (ratio, position) over the blocks: the first minimum in N order
        bt = 0.;
        bp = 0x7fffffff;
        for (int b = 0; b < (int)gridDim.x; ++b) {
            const double ot = __ldcg(&sc->val[b]);
            const int op = __ldcg(&sc->idx[b]);
            if (op != 0x7fffffff && (bp == 0x7fffffff || ot < bt || (ot == bt && op < bp))) { bt = ot; bp = op; }
        }
        const int nan_flag = atomicExch(&sc->flag, 0);
        if (nan_flag) { st->err = kErrNaNDualRatio; st->status = ELLP_INFEASIBLE; }
        else if (bp == 0x7fffffff) st->status = ELLP_INFEASIBLE;  // :281-284 dual unbounded
        else {
            st->q_pos = bp;
            st->q_var = lp.Nv[bp];
            st->q_side = lp.Ns[bp];
            st->theta_d = neg ? -bt : bt;  // :286-289
        }
    }
}

// dual steepest edge: leaving row = argmax infeasibility_i^2 / w_i (ties: smallest position); same outputs as k_dual_leaving
__global__ void __launch_bounds__(256) k_dual_leaving_dse(DevLP lp, PivotState* st, SelScratch* sc) {
    if (blockIdx.x == 0 && threadIdx.x == 0) st->do_update = 0;
    if (st->status != kRunning) return;
    __shared__ double s_v[8];
    __shared__ int s_i[8];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const unsigned full = 0xffffffffu;
    double bv = -1.;
    int bi = 0x7fffffff;
    for (int i = blockIdx.x * blockDim.x + tid; i < lp.m; i += gridDim.x * blockDim.x) {
        const int var = lp.Bv[i];
        const double x_i = lp.x[var];
        const int kind = lp.kind[var];
        double p = 0.;
        if ((kind == ELLP_LOWER || kind == ELLP_TWOSIDED) && x_i < lp.lb[var] - kEps) p = lp.lb[var] - x_i;
        if ((kind == ELLP_UPPER || kind == ELLP_TWOSIDED) && x_i > lp.ub[var] + kEps) p = x_i - lp.ub[var];
        if (p > 0.) {
            const double score = p * p / fmax(lp.w[i], 1e-300);
            if (score > bv) { bv = score; bi = i; }
        }
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
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi)) { bv = s_v[w]; bi = s_i[w]; }
        sc->val[blockIdx.x] = bv;
        sc->idx[blockIdx.x] = bi;
    }
    if (!sel_arrive_last(sc, &s_last)) return;
    if (tid == 0) {
        bv = -1.;
        bi = 0x7fffffff;
        for (int b = 0; b < (int)gridDim.x; ++b) {
            const double ov = __ldcg(&sc->val[b]);
            const int oi = __ldcg(&sc->idx[b]);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (bi == 0x7fffffff) {
            st->status = ELLP_OPTIMAL;
        } else {
            const int var = lp.Bv[bi];
            const double x_i = lp.x[var];
            const int kind = lp.kind[var];
            double delta;
            int side;
            if ((kind == ELLP_UPPER || kind == ELLP_TWOSIDED) && x_i > lp.ub[var] + kEps) { delta = x_i - lp.ub[var]; side = ELLP_NB_UPPER; }
            else { delta = x_i - lp.lb[var]; side = ELLP_NB_LOWER; }
            st->r_pos = bi;
            st->leave_var = var;
            st->delta = delta;
            st->new_side = side;
        }
    }
}

// Harris' two-pass dual ratio test: theta_max = min over eligible j of (d_j +- tol) / alpha~_j, then the largest |alpha~_j|
// among { j : d_j / alpha~_j <= theta_max } (ties: first position)
__global__ void __launch_bounds__(1024) k_select_dual_harris(DevLP lp, double tol, PivotState* st) {
    if (st->status != kRunning) return;
    __shared__ double s_t[32];
    __shared__ int s_p[32];
    const int tid = threadIdx.x;
    const unsigned full = 0xffffffffu;
    const bool neg = st->delta < 0.;
    double tmin = CUDART_INF;
    for (int j = tid; j < lp.nN; j += blockDim.x) {
        double a = lp.rN[j];
        if (neg) a = -a;
        const int side = lp.Ns[j];
        const bool keep = (side == ELLP_NB_LOWER) ? (a > kEps) : (side == ELLP_NB_UPPER ? (a < -kEps) : (fabs(a) > kEps));
        if (keep) {
            const double t = lp.d[lp.Nv[j]] / a + tol / fabs(a);
            if (t < tmin) tmin = t;
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) tmin = fmin(tmin, __shfl_xor_sync(full, tmin, off));
    if ((tid & 31) == 0) s_t[tid >> 5] = tmin;
    __syncthreads();
    tmin = s_t[0];
    for (int w = 1; w < 32; ++w) tmin = fmin(tmin, s_t[w]);
    __syncthreads();
    double ba = -1.;
    int bp = 0x7fffffff;
    for (int j = tid; j < lp.nN; j += blockDim.x) {
        double a = lp.rN[j];
        if (neg) a = -a;
        const int side = lp.Ns[j];
        const bool keep = (side == ELLP_NB_LOWER) ? (a > kEps) : (side == ELLP_NB_UPPER ? (a < -kEps) : (fabs(a) > kEps));
        if (keep && lp.d[lp.Nv[j]] / a <= tmin && fabs(a) > ba) { ba = fabs(a); bp = j; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double oa = __shfl_xor_sync(full, ba, off);
        const int op = __shfl_xor_sync(full, bp, off);
        if (oa > ba || (oa == ba && op < bp)) { ba = oa; bp = op; }
    }
    if ((tid & 31) == 0) { s_t[tid >> 5] = ba; s_p[tid >> 5] = bp; }
    __syncthreads();
    if (tid == 0) {
        ba = -1.;
        bp = 0x7fffffff;
        for (int w = 0; w < 32; ++w) if (s_t[w] > ba || (s_t[w] == ba && s_p[w] < bp)) { ba = s_t[w]; bp = s_p[w]; }
        if (bp == 0x7fffffff) st->status = ELLP_INFEASIBLE;  // dual unbounded
        else {
            double a = lp.rN[bp];
            if (neg) a = -a;
            const double t = lp.d[lp.Nv[bp]] / a;
            st->q_pos = bp;
            st->q_var = lp.Nv[bp];
            st->q_side = lp.Ns[bp];
            st->theta_d = neg ? -t : t;
        }
    }
}

// dual :294-316, element-wise part over the whole grid: alpha_q = B^-1 a_q (sum of the split-K partials, fixed order),
// d_N -= theta_d alpha, y += theta_d rho, x_B -= theta_p alpha_q, and the scaled pivot row for k_rank1.
// Every thread re-derives alpha_q[r] itself (KS cached loads), so no second launch is needed to broadcast it.
__global__ void __launch_bounds__(256) k_dual_update_vec(DevLP lp, int KS, const PivotState* st) {
    if (st->status != kRunning) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = lp.m;
    const int r_pos = st->r_pos;
    const double theta_d = st->theta_d, delta = st->delta;
    const double alpha_r = sum_partials(lp.part, lp.ld, KS, r_pos);
    const double theta_p = delta / alpha_r;  // :306
    if (t < m) {
        const double a = sum_partials(lp.part, lp.ld, KS, t);
        lp.dcol[t] = a;
        const double rho = lp.rho[t];
        lp.y[t] = lp.y[t] + theta_d * rho;       // :304
        const int var = lp.Bv[t];
        lp.x[var] = lp.x[var] - theta_p * a;     // :310-312
        lp.prow[t] = rho / alpha_r;
    }
    if (t < lp.nN) {                             // :298-300
        const int var = lp.Nv[t];
        lp.d[var] = lp.d[var] - theta_d * lp.rN[t];
    }
}

// dual :296, :302, :314-333: the scalar tail (entering / leaving entries, objective, index swap, trace)
__global__ void k_dual_update_tail(DevLP lp, int KS, PivotState* st) {
    if (st->status != kRunning) return;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int r_pos = st->r_pos, q_pos = st->q_pos, q_var = st->q_var, leave_var = st->leave_var;
    const double theta_d = st->theta_d, delta = st->delta;
    const double alpha_r = lp.dcol[r_pos];
    const double theta_p = delta / alpha_r;
    lp.d[leave_var] = -theta_d;  // :296 (the leaving variable is not in N, so the order vs :298 is immaterial)
    lp.d[q_var] = 0.;            // :302
    lp.x[q_var] = lp.x[q_var] + theta_p;  // :314
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
    lp.Bv[r_pos] = q_var;                 // :322-323
    lp.Nv[q_pos] = leave_var;
    lp.Ns[q_pos] = (uint8_t)st->new_side;
    lp.cB[r_pos] = lp.c[q_var];
    st->alpha_r = alpha_r;
    st->step = theta_p;
    st->do_update = 1;
    st->pivots += 1;
    if (st->pivots >= st->max_iter) st->status = ELLP_MAXITER;  // :191-194 at the next loop head
}

// ------------------------------------------------------------------------------------------------
// Refactorisation: Gauss-Jordan on G = [A_B | I] with the pivot rule of the reference's LU (first
// max |a| at or below the diagonal), so the pivots equal the U_kk the reference tests against EPS
// (primal :175-179).  Each elimination step is one k_rank1 launch on the ld x (2m - k) trailing block.
// ------------------------------------------------------------------------------------------------
__global__ void k_gj_init(DevLP lp) {
    const int j = blockIdx.x;  // column of G
    const int m = lp.m;
    for (int64_t i = threadIdx.x; i < lp.ld; i += blockDim.x) {
        double v;
        if (j < m) v = (i < m) ? lp.A[(int64_t)lp.Bv[j] * lp.ld + i] : 0.;
        else v = (i == j - m) ? 1. : 0.;
        lp.G[(int64_t)j * lp.ld + i] = v;
    }
}

// M: ld x ncols working matrix; pcol: column whose entries at rows >= k are searched for the pivot
// (k itself for [A_B | I]; the basis column Bv[k] when a tableau is built in place)
__global__ void __launch_bounds__(1024) k_gj_pivot(double* __restrict__ M, int64_t ld, int m, const int32_t* __restrict__ pcol_of,
                                                   int k, double* __restrict__ dcol, PivotState* st) {
    if (st->err) return;
    __shared__ double s_v[32];
    __shared__ int s_i[32];
    const int tid = threadIdx.x;
    const int pc = pcol_of ? pcol_of[k] : k;
    const double* col = M + (int64_t)pc * ld;
    double bv = -1.;
    int bi = 0x7fffffff;
    for (int i = k + tid; i < m; i += blockDim.x) {
        const double v = fabs(col[i]);
        if (v > bv) { bv = v; bi = i; }  // strict: first max per thread
    }
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double ov = __shfl_xor_sync(full, bv, off);
        const int oi = __shfl_xor_sync(full, bi, off);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { s_v[tid >> 5] = bv; s_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
        const int nw = blockDim.x >> 5;
        bv = (tid < nw) ? s_v[tid] : -1.;
        bi = (tid < nw) ? s_i[tid] : 0x7fffffff;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const double ov = __shfl_xor_sync(full, bv, off);
            const int oi = __shfl_xor_sync(full, bi, off);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (tid == 0) {
            s_i[0] = bi;
            st->gj_piv = bi;
            st->r_pos = k;
            st->do_update = 1;
            if (!(bv >= kEps)) { st->err = kErrSingular; st->do_update = 0; s_i[0] = -1; }  // |U_kk| < EPS
            else st->alpha_r = col[bi];
        }
    }
    __syncthreads();
    const int p = s_i[0];
    if (p < 0) return;
    // pivot column after the row swap k <-> p (padding rows stay zero)
    for (int64_t i = tid; i < ld; i += blockDim.x) {
        const int64_t src = (i == k) ? p : ((i == p) ? k : i);
        dcol[i] = (i < m) ? col[src] : 0.;
    }
}

// swaps rows k and p of columns [c_begin, c_end) and emits the scaled pivot row prow[j - c_begin] = M[k, j] / pivot
__global__ void k_gj_swap_gather(double* __restrict__ M, int64_t ld, int c_begin, int c_end, int k, double* __restrict__ prow,
                                 const PivotState* st) {
    if (st->err) return;
    const int j = c_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c_end) return;
    const int p = st->gj_piv;
    double* col = M + (int64_t)j * ld;
    const double a = col[k], b = col[p];
    if (p != k) { col[k] = b; col[p] = a; }
    prow[j - c_begin] = b / st->alpha_r;
}

// revised engine: is A_B diagonal (slack / artificial starting bases are)?  Then B^-1 = diag(1 / a_ii) needs no factorisation.
__global__ void k_check_diag_basis(const double* __restrict__ A, int64_t ld, int m, const int32_t* __restrict__ Bv, int* __restrict__ flags) {
    const int i = blockIdx.x;
    const double* col = A + (int64_t)Bv[i] * ld;
    bool offdiag = false;
    for (int k = threadIdx.x; k < m; k += blockDim.x)
        if (k != i && col[k] != 0.) offdiag = true;
    if (offdiag) flags[0] = 1;
    if (threadIdx.x == 0 && fabs(col[i]) < kEps) flags[1] = 1;  // |U_ii| < EPS (primal :175-179)
}

__global__ void k_diag_inverse(DevLP lp) {
    const int j = blockIdx.x;  // column of Binv
    const double d = lp.A[(int64_t)lp.Bv[j] * lp.ld + j];
    for (int64_t i = threadIdx.x; i < lp.ld; i += blockDim.x) lp.Binv[(int64_t)j * lp.ld + i] = (i == j) ? 1.0 / d : 0.;
}

// tableau engine: is column Bv[i] of T exactly e_i for every i (slack / identity starting basis)?
__global__ void k_check_identity_basis(const double* __restrict__ T, int64_t ld, int m, const int32_t* __restrict__ Bv,
                                       int* __restrict__ mismatch) {
    const int i = blockIdx.x;
    const double* col = T + (int64_t)Bv[i] * ld;
    bool bad = false;
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        const double want = (k == i) ? 1. : 0.;
        if (col[k] != want) bad = true;
    }
    if (bad) *mismatch = 1;
}

// tableau engine pricing: reduced costs are a maintained row, so the Dantzig keys are a pure O(n - m) pass
__global__ void k_price_tab(const double* __restrict__ dj, const int32_t* __restrict__ Nv, const uint8_t* __restrict__ Ns, int nN,
                            double* __restrict__ rN, double* __restrict__ key, PivotState* st, int condensed) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->do_update = 0; st->do_step = 0; }
    if (st->status != kRunning) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nN) return;
    const double r = condensed ? dj[j] : dj[Nv[j]];
    const int side = Ns[j];
    double k = -1.0;
    if (!(fabs(r) < kEps)) {  // primal :258-269
        if (r > 0. && side == ELLP_NB_UPPER) k = r;
        else if (!(r > 0.) && side == ELLP_NB_LOWER) k = -r;
        else if (side == ELLP_NB_FREE) k = fabs(r);
    }
    rN[j] = r;
    key[j] = k;
}

// ------------------------------------------------------------------------------------------------
// Column-sharded tableau (one rank per GPU).  Per pivot: local Dantzig keys -> all-gather of the local maxima ->
// local candidate under the order-free rule (largest variable index within EPS of the GLOBAL maximum) ->
// all-gather of the candidates -> every rank derives the same entering variable; its owner contributes the pivot
// column to an all-reduce (the others contribute zeros) -> replicated ratio test -> local rank-1 update.
// ------------------------------------------------------------------------------------------------
__global__ void k_price_shard(DevLP lp, PivotState* st) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->do_update = 0; st->do_step = 0; }
    if (st->status != kRunning) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= lp.n) return;
    const int side = lp.colstat[j];
    double k = -1.0;
    if (side != kColBasic) {
        const double r = lp.dj[j];
        if (!(fabs(r) < kEps)) {
            if (r > 0. && side == ELLP_NB_UPPER) k = r;
            else if (!(r > 0.) && side == ELLP_NB_LOWER) k = -r;
            else if (side == ELLP_NB_FREE) k = fabs(r);
        }
    }
    lp.key[j] = k;
}

__device__ __forceinline__ double block_max_1024(double v, double* s) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? s[threadIdx.x] : -1.0;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (threadIdx.x == 0) s[0] = v;
    }
    __syncthreads();
    return s[0];
}

__global__ void __launch_bounds__(1024) k_shard_local_max(DevLP lp, const PivotState* st) {
    __shared__ double s[32];
    if (st->status != kRunning) return;
    double v = -1.0;
    for (int j = threadIdx.x; j < lp.n; j += blockDim.x) v = fmax(v, lp.key[j]);
    v = block_max_1024(v, s);
    if (threadIdx.x == 0) lp.xchg[0] = v;
}

// xchg[16..16+G): gathered local maxima.  Candidate = largest local variable index with key > K* - EPS.
__global__ void __launch_bounds__(1024) k_shard_pick(DevLP lp, int G, PivotState* st) {
    __shared__ int s_i[32];
    if (st->status != kRunning) return;
    double kmax = -1.0;
    for (int g = 0; g < G; ++g) kmax = fmax(kmax, lp.xchg[16 + g]);
    if (kmax == -1.0) {  // no candidate on any rank: optimal (primal :289-292); identical decision on every rank
        __syncthreads();
        if (threadIdx.x == 0) st->status = ELLP_OPTIMAL;
        return;
    }
    int best = -1;
    for (int j = threadIdx.x; j < lp.n; j += blockDim.x) {
        const double k = lp.key[j];
        if (k != -1.0 && (kmax - k < kEps)) best = max(best, j);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, off));
    if ((threadIdx.x & 31) == 0) s_i[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        best = (threadIdx.x < (blockDim.x >> 5)) ? s_i[threadIdx.x] : -1;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, off));
        if (threadIdx.x == 0) {
            lp.xchg[8] = (best >= 0) ? (double)(lp.col_lo + best) : -1.0;
            lp.xchg[9] = (best >= 0) ? lp.dj[best] : 0.;
            lp.xchg[10] = (best >= 0) ? (double)lp.colstat[best] : 0.;
        }
    }
}

// xchg[32..32+3G): gathered candidates.  Every rank derives the same winner; the owner stages its pivot column.
// cnt > 0 (blocked engine): the owner applies the cnt pending rank-1 corrections to its stale column first.
__global__ void k_shard_stage_column(DevLP lp, int G, int cnt, PivotState* st, double* __restrict__ sendcol) {
    if (st->status != kRunning) return;
    double qv = -1.0, rq = 0., side = 0.;
    for (int g = 0; g < G; ++g) {
        const double v = lp.xchg[32 + 3 * g];
        if (v > qv) { qv = v; rq = lp.xchg[32 + 3 * g + 1]; side = lp.xchg[32 + 3 * g + 2]; }
    }
    const int q_var = (int)qv;
    const bool mine = (q_var >= lp.col_lo && q_var < lp.col_lo + lp.n);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < lp.ld) {
        double a = 0.;
        if (mine) {
            const int ql = q_var - lp.col_lo;
            a = lp.T[(int64_t)ql * lp.ld + i];
            for (int j = 0; j < cnt; ++j) a = fma(-lp.U[(int64_t)j * lp.ld + i], __ldg(lp.V + (int64_t)j * lp.ldv + ql), a);
        }
        sendcol[i] = a;
    }
    if (i == 0) {
        st->lmin_bits = 0x7ff0000000000000ll;
        st->q_var = q_var;
        st->q_pos = -1;
        st->q_side = (int)side;
        st->rq = rq;
    }
}

// every rank checks the basis columns it owns: column Bv[i] must be e_i
__global__ void k_check_identity_shard(DevLP lp, int* __restrict__ mismatch) {
    const int i = blockIdx.x;
    const int var = lp.Bv[i];
    if (var < lp.col_lo || var >= lp.col_lo + lp.n) return;
    const double* col = lp.T + (int64_t)(var - lp.col_lo) * lp.ld;
    bool bad = false;
    for (int k = threadIdx.x; k < lp.m; k += blockDim.x) {
        const double want = (k == i) ? 1. : 0.;
        if (col[k] != want) bad = true;
    }
    if (bad) *mismatch = 1;
}

__global__ void k_shard_init(DevLP lp, const uint8_t* __restrict__ side_of_var /* n_glob: ELLP_NB_* or kColBasic */) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < lp.n) lp.colstat[j] = side_of_var[lp.col_lo + j];
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__global__ void k_init_cB(DevLP lp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < lp.ld) lp.cB[i] = (i < lp.m) ? lp.c[lp.Bv[i]] : 0.;
}

__global__ void __launch_bounds__(1024) k_obj_dot(const double* __restrict__ c, const double* __restrict__ x, int n, PivotState* st) {
    __shared__ double s[32];
    double a = 0.;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a = fma(c[i], x[i], a);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        a = (threadIdx.x < (blockDim.x >> 5)) ? s[threadIdx.x] : 0.;
        a = warp_sum(a);
        if (threadIdx.x == 0) st->obj = a;
    }
}

// counter-based U(lo,hi): splitmix64(seed + golden*(offset+i)) -> 53-bit mantissa
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void k_fill_uniform(double* __restrict__ out, uint64_t count, uint64_t seed, uint64_t offset, double lo, double hi) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t h = splitmix64(seed * 0x2545f4914f6cdd1dull + offset + i);
        const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
        out[i] = lo + (hi - lo) * u;
    }
}


// ---- synthetic dense LP built directly in HBM (bench workloads) ----------------------------
//   min -c.x  s.t.  A x + s = b, x, s >= 0      A ~ U(0,1) m x ns,  b_i ~ U(1,2) * ns/4,  c_j ~ U(0.5,1.5)
// columns [0, ns) structural, [ns, ns+m) slack (identity); starting point = slack basis (primal feasible).
// col_lo/col_hi select the locally stored column range (column sharding); element values depend only on the
// GLOBAL (row, column) index so every sharding sees the same LP.
__global__ void k_gen_dense_cols(double* __restrict__ A, int64_t ld, int m, int64_t ns, int64_t col_lo, int64_t col_hi, uint64_t seed,
                                 double slack_sign, double struct_sign) {
    const int64_t ncol = col_hi - col_lo;
    const int64_t total = ncol * ld;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t jl = e / ld, i = e - jl * ld;
        const int64_t j = col_lo + jl;
        double v = 0.;
        if (i < m) {
            if (j < ns) {
                const uint64_t h = splitmix64(seed * 0x2545f4914f6cdd1dull + (uint64_t)(j * m + i));
                v = struct_sign * ((double)(h >> 11) * (1.0 / 9007199254740992.0));
            } else {
                v = (j - ns == i) ? slack_sign : 0.;
            }
        }
        A[e] = v;
    }
}

// variant 0: min -c.x, Ax + s = b (slack basis primal feasible); variant 1: min c.x, Ax - s = b (slack basis dual
// feasible: x_B = -b < 0, y = 0, d = c >= 0) -- SURVEY 8(d) configs 4/5 and 3
__global__ void k_gen_dense_vectors(DevLP lp, int64_t ns, uint64_t seed, int variant) {
    const int m = lp.m, n = lp.n_glob;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        double cj = 0., xj = 0.;
        if (j < ns) {
            const uint64_t h = splitmix64((seed + 1) * 0x2545f4914f6cdd1dull + (uint64_t)j);
            cj = 0.5 + (double)(h >> 11) * (1.0 / 9007199254740992.0);
            if (variant == 0) cj = -cj;
        } else {
            const uint64_t h = splitmix64((seed + 2) * 0x2545f4914f6cdd1dull + (uint64_t)(j - ns));
            xj = (1.0 + (double)(h >> 11) * (1.0 / 9007199254740992.0)) * ((double)ns * 0.25);
            const int i = (int)(j - ns);
            const_cast<double*>(lp.b)[i] = xj;
            if (variant == 1) { xj = -xj; lp.y[i] = 0.; }
            lp.Bv[i] = (int32_t)j;
        }
        const_cast<double*>(lp.c)[j] = cj;
        const_cast<double*>(lp.lb)[j] = 0.;
        const_cast<double*>(lp.ub)[j] = 0.;
        const_cast<uint8_t*>(lp.kind)[j] = ELLP_LOWER;
        lp.x[j] = xj;
        if (variant == 1) lp.d[j] = cj;
        if (j < ns && lp.colstat == nullptr) { lp.Nv[j] = (int32_t)j; lp.Ns[j] = ELLP_NB_LOWER; }
        if (lp.colstat != nullptr && j >= lp.col_lo && j < lp.col_lo + lp.n) lp.colstat[j - lp.col_lo] = (j < ns) ? (uint8_t)ELLP_NB_LOWER : kColBasic;
    }
    (void)m;
}

}  // namespace ellp
