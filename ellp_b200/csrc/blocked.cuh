// blocked.cuh -- deferred ("rank-k") row reduction of the tableau engine.
//
// The rank-1 engine (k_rank1, kernels.cuh) re-streams the whole m x n tableau through HBM for every pivot:
// 16*m*n bytes per pivot, which is what bounds pivots/s once K3 sits at the HBM roofline.  A simplex iteration,
// however, only needs ONE column (the entering one) and ONE row (the leaving one) of the current tableau.  The
// blocked engine therefore keeps the tableau stale and carries the last k pivots as a low-rank correction
//
//        T_current = T_stale - U V        U: ld x k  (column j = pivot column of pending pivot j, minus e_r)
//                                         V: k x ldv (row j    = scaled pivot row of pending pivot j)
//
// (the rank-1 update of pivot (r, q) is T <- T - (d - e_r) p with d = T[:, q], p = T[r, :] / d_r: rows i != r
// get T[i,:] - d_i p, row r gets T[r,:] - (d_r - 1) p = p).  Per pivot the engine touches
//   * one column  d = T_stale[:, q] - U V[:, q]              (k_ratio_prep, cnt > 0)      8*m*(1 + cnt) bytes, L2 resident
//   * one row     p = (T_stale[r, :] - U[r, :] V) / d_r      (k_blk_row)                  8*n*(1 + cnt) bytes + n sectors
// and every k pivots one launch applies T -= U V to the whole tableau on the fp64 tensor pipe (k_blk_flush: DMMA
// m8n8k4, operands staged in shared memory, T streamed HBM -> registers -> HBM exactly once).  HBM traffic per
// pivot drops from 16*m*n to 16*m*n / k; the flush stays HBM-bound while 2*k flop per 16 B fit under the DMMA rate.
//
// Decision parity: every number a decision is taken on (the entering column, the ratios, the reduced-cost row, x)
// is produced by the same sequence of fused multiply-adds as in the rank-1 engine (pending pivots applied in
// order); only the stored tableau differs in the last bits after a flush (tensor-pipe accumulation order), which
// is far below the reference's EPS = 1e-10 decision tolerance.  Reference lines replaced: the per-iteration
// `A_B.clone().lu()` + solves, primal_simplex_solver.rs:173-189,295.
#pragma once
#include "kernels.cuh"
#include "refactor.cuh"

namespace ellp {

constexpr int kBlkMax = 64;  // largest number of pending pivots (slots) supported by the kernels below

// K2c (blocked): the primal step x_B += lambda d, x_q +-= lambda (primal :408-417), the scaled pivot row of the CURRENT
// tableau, the reduced-cost row update d_j -= d_q p_j, and the new (U, V) slot.  A bound flip or a finished solve
// leaves an all-zero slot, so the host can schedule slots without knowing what the device decided.
__global__ void __launch_bounds__(256) k_blk_row(DevLP lp, int slot, PivotState* st) {
    __shared__ double su[kBlkMax];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = lp.nT, m = lp.m;
    if (st->do_step) {
        const double lambda = st->step;
        const bool at_lower = (st->q_side == ELLP_NB_LOWER);
        if (t < m) {
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
    double* Uslot = lp.U + (int64_t)slot * lp.ld;
    double* Vslot = lp.V + (int64_t)slot * lp.ldv;
    if (!st->do_update) {
        if (t < lp.ld) Uslot[t] = 0.;
        if (t < lp.ldv) Vslot[t] = 0.;
        return;
    }
    const int r = st->r_pos;
    const int qp = lp.condensed ? st->q_pos : -1;
    if (threadIdx.x < slot) su[threadIdx.x] = lp.U[(int64_t)threadIdx.x * lp.ld + r];
    __syncthreads();
    if (t < n) {
        if (t == qp) {
            // condensed tableau: this stored column is handed over to the leaving variable, whose current column is e_r
            // (see k_step_gather_cond): pivot-row entry 1/alpha_r, reduced cost 0 - d_q / alpha_r
            const double p = 1.0 / st->alpha_r;
            Vslot[t] = p;
            lp.dj[t] = fma(-st->rq, p, 0.);
        } else {
            double e = lp.T[t * lp.ld + r];
            for (int j = 0; j < slot; ++j) e = fma(-su[j], lp.V[(int64_t)j * lp.ldv + t], e);
            const double p = e / st->alpha_r;
            Vslot[t] = p;
            lp.dj[t] = fma(-st->rq, p, lp.dj[t]);
        }
    } else if (t < lp.ldv) {
        Vslot[t] = 0.;
    }
    if (t < lp.ld) {
        Uslot[t] = (t < m ? lp.dcol[t] : 0.) - (t == r ? 1. : 0.);
        if (qp >= 0) lp.T[(int64_t)qp * lp.ld + t] = (t == r) ? 1. : 0.;  // stale column := e_r ...
    }
    if (qp >= 0 && t < slot) lp.V[t * lp.ldv + qp] = 0.;                 // ... with no pending correction before this slot
}

// ------------------------------------------------------------------------------------------------
// K3b: T -= U V for the whole tableau, fp64 tensor pipe.
//   CTA = 128 rows x (kFlushColsPerCta columns, in steps of 64); 8 warps as 4 (rows) x 2 (columns); warp tile 32 x 32.
//   The mma is issued on the TRANSPOSED tile (mma rows = tableau columns, mma columns = tableau rows) so that the two
//   accumulator values a lane owns are two consecutive rows of one column: T moves HBM <-> registers as 16-byte
//   accesses, 64 contiguous bytes per column per quad, each sector touched once.
//   -U (128 x K) is staged once per CTA, V (K x 64) per column step (double-buffered); padded strides (== 4 mod 16 doubles) keep the 8-byte
//   fragment loads at the 2-wavefront floor.
// ------------------------------------------------------------------------------------------------
constexpr int kFlushRows = 128;
constexpr int kFlushCols = 64;
constexpr int kFlushSU = kFlushRows + 4;
constexpr int kFlushSV = kFlushCols + 4;

inline size_t blk_flush_smem_bytes(int K4) { return sizeof(double) * (size_t)K4 * (kFlushSU + (K4 <= 48 ? 2 : 1) * kFlushSV); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// One CTA walks its column steps; two CTAs share an SM so that one streams T (HBM <-> registers) while the other
// occupies the tensor pipe.  Within a CTA the V tile of step s+1 is fetched L2 -> shared memory by cp.async while the
// mma of step s runs (double buffer; single buffer above 48 slots, where two stages no longer fit twice per SM).
template <bool STREAM>
__global__ void __launch_bounds__(256, 2) k_blk_flush(double* __restrict__ T, int64_t ld, int R, int C, const double* __restrict__ U,
                                                      const double* __restrict__ V, int64_t ldv, int cnt, int col_steps) {
    extern __shared__ __align__(16) double blk_smem[];
    const int K4 = (cnt + 3) & ~3;
    const bool dbuf = K4 <= 48;
    double* sU = blk_smem;                         // sU[j][row] = -U[row0 + row, j]
    double* sV0 = blk_smem + K4 * kFlushSU;        // sV[j][col] = V[j, col0 + col]
    double* sV1 = dbuf ? sV0 + K4 * kFlushSV : sV0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * kFlushRows;
    const int wr = (warp & 3) * 32, wc = (warp >> 2) * 32;
    const int fq = lane >> 2, fk = lane & 3;
    const int ksteps = K4 >> 2;
    const int64_t step0 = (int64_t)blockIdx.y * col_steps;
    const int64_t steps_total = (C + kFlushCols - 1) / kFlushCols;
    const int nsteps = (int)max((int64_t)0, min((int64_t)col_steps, steps_total - step0));
    if (nsteps == 0) return;
    const bool v_aligned = ((ldv & 1) == 0) && ((reinterpret_cast<uintptr_t>(V) & 15) == 0);

    auto stage_v = [&](int s) {  // V tile of step s -> shared memory, 16-byte cp.async where the tile is full and aligned
        double* sV = (s & 1) ? sV1 : sV0;
        const int64_t col0 = (step0 + s) * kFlushCols;
        if (v_aligned && col0 + kFlushCols <= C) {
            for (int e = tid; e < K4 * (kFlushCols / 2); e += 256) {
                const int j = e >> 5, c2 = (e & 31) * 2;
                double* dst = sV + j * kFlushSV + c2;
                if (j < cnt) cp_async16(dst, V + (int64_t)j * ldv + col0 + c2);
                else { dst[0] = 0.; dst[1] = 0.; }
            }
        } else {
            for (int e = tid; e < K4 * kFlushCols; e += 256) {
                const int j = e >> 6, c = e & (kFlushCols - 1);
                sV[j * kFlushSV + c] = (j < cnt && col0 + c < C) ? V[(int64_t)j * ldv + col0 + c] : 0.;
            }
        }
        cp_async_commit();
    };

    stage_v(0);
    for (int e = tid; e < K4 * kFlushRows; e += 256) {
        const int j = e >> 7, i = e & (kFlushRows - 1);
        sU[j * kFlushSU + i] = (j < cnt && row0 + i < ld) ? -U[(int64_t)j * ld + row0 + i] : 0.;
    }
    for (int s = 0; s < nsteps; ++s) {
        const int64_t col0 = (step0 + s) * kFlushCols;
        // issue the loads of this step's 32 x 32 warp tile first: 16 x 16 B in flight per lane
        double2 acc[4][4];  // [column tile][row tile]: rows row0+wr+8*rt+2*fk,+1 ; column col0+wc+8*ct+fq
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int64_t c = col0 + wc + ct * 8 + fq;
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                const int64_t r = row0 + wr + rt * 8 + 2 * fk;
                if (c < C && r < ld) {
                    const double* p = T + c * ld + r;
                    acc[ct][rt] = STREAM ? ld_f64x2_stream(p) : ld_f64x2(p);
                } else {
                    acc[ct][rt] = make_double2(0., 0.);
                }
            }
        }
        if (!dbuf && s > 0) { __syncthreads(); stage_v(s); }  // single buffer: refill after every warp left step s-1
        cp_async_wait<0>();
        __syncthreads();  // V tile s (and, for s == 0, -U) visible to every warp; every warp left step s-1
        if (dbuf && s + 1 < nsteps) stage_v(s + 1);
        const double* sV = (s & 1) ? sV1 : sV0;
#pragma unroll 2
        for (int ks = 0; ks < ksteps; ++ks) {
            double a[4], b[4];
            const int j = ks * 4 + fk;
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) a[ct] = sV[j * kFlushSV + wc + ct * 8 + fq];
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) b[rt] = sU[j * kFlushSU + wr + rt * 8 + fq];
#pragma unroll
            for (int ct = 0; ct < 4; ++ct)
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) dmma_m8n8k4(acc[ct][rt].x, acc[ct][rt].y, a[ct], b[rt]);
        }
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int64_t c = col0 + wc + ct * 8 + fq;
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                const int64_t r = row0 + wr + rt * 8 + 2 * fk;
                if (c < C && r < ld) {
                    double* p = T + c * ld + r;
                    if (STREAM) st_f64x2_stream(p, acc[ct][rt]);
                    else st_f64x2(p, acc[ct][rt]);
                }
            }
        }
    }
    (void)R;
}

}  // namespace ellp
