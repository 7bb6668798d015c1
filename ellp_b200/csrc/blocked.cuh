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
#include <cooperative_groups.h>
#include "refactor.cuh"

namespace ellp {

constexpr int kBlkMax = 64;  // largest number of pending pivots (slots) supported by the kernels below

// K2c (blocked): the primal step x_B += lambda d, x_q +-= lambda (primal :408-417), the scaled pivot row of the CURRENT
// tableau, the reduced-cost row update d_j -= d_q p_j, and the new (U, V) slot.  A bound flip or a finished solve
// leaves an all-zero slot, so the host can schedule slots without knowing what the device decided.
// Grid-stride over [t0, ...) with stride `stride`; every thread of the block must call it (one __syncthreads).
__device__ __forceinline__ void blk_row_body(const DevLP& lp, int slot, const PivotState* st, int64_t t0, int64_t stride, double* su) {
    const int n = lp.nT, m = lp.m;
    const int do_step = __ldcg(&st->do_step), do_update = __ldcg(&st->do_update);
    const int r = __ldcg(&st->r_pos);
    const int64_t tmax = max(lp.ld, lp.ldv);
    if (do_step) {
        const double lambda = __ldcg(&st->step);
        const bool at_lower = (__ldcg(&st->q_side) == ELLP_NB_LOWER);
        const int leave_var = __ldcg(&st->leave_var);
        for (int64_t t = t0; t < m; t += stride) {
            const double a = __ldcg(lp.dcol + t);
            const double d_i = at_lower ? -a : a;
            const int var = (do_update && t == r) ? leave_var : __ldcg(lp.Bv + t);
            lp.x[var] = __ldcg(lp.x + var) + lambda * d_i;
        }
        if (t0 == 0) {
            const int q = __ldcg(&st->q_var);
            lp.x[q] = at_lower ? __ldcg(lp.x + q) + lambda : __ldcg(lp.x + q) - lambda;
        }
    }
    double* Uslot = lp.U + (int64_t)slot * lp.ld;
    double* Vslot = lp.V + (int64_t)slot * lp.ldv;
    if (!do_update) {
        for (int64_t t = t0; t < tmax; t += stride) {
            if (t < lp.ld) Uslot[t] = 0.;
            if (t < lp.ldv) Vslot[t] = 0.;
        }
        return;
    }
    int qp = -1;  // locally stored column that is handed over to the leaving variable (condensed tableau)
    if (lp.condensed) {
        const int q = __ldcg(&st->q_pos) - lp.pos_lo;
        if (q >= 0 && q < n) qp = q;
    }
    const double alpha_r = __ldcg(&st->alpha_r), rq = __ldcg(&st->rq);
    if (threadIdx.x < slot) su[threadIdx.x] = __ldcg(lp.U + (int64_t)threadIdx.x * lp.ld + r);
    __syncthreads();
    for (int64_t t = t0; t < tmax; t += stride) {
        if (t < n) {
            if (t == qp) {
                // condensed tableau: this stored column is handed over to the leaving variable, whose current column is e_r
                // (see k_step_gather_cond): pivot-row entry 1/alpha_r, reduced cost 0 - d_q / alpha_r
                const double p = 1.0 / alpha_r;
                Vslot[t] = p;
                lp.dj[t] = fma(-rq, p, 0.);
            } else {
                double e = __ldcg(lp.T + t * lp.ld + r);
                for (int j = 0; j < slot; ++j) e = fma(-su[j], __ldcg(lp.V + (int64_t)j * lp.ldv + t), e);
                const double p = e / alpha_r;
                Vslot[t] = p;
                lp.dj[t] = fma(-rq, p, __ldcg(lp.dj + t));
            }
        } else if (t < lp.ldv) {
            Vslot[t] = 0.;
        }
        if (t < lp.ld) {
            Uslot[t] = (t < m ? __ldcg(lp.dcol + t) : 0.) - (t == r ? 1. : 0.);
            if (qp >= 0) lp.T[(int64_t)qp * lp.ld + t] = (t == r) ? 1. : 0.;  // stale column := e_r ...
        }
        if (qp >= 0 && t < slot) lp.V[t * lp.ldv + qp] = 0.;                 // ... with no pending correction before this slot
    }
}

__global__ void __launch_bounds__(256) k_blk_row(DevLP lp, int slot, PivotState* st) {
    __shared__ double su[kBlkMax];
    blk_row_body(lp, slot, st, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, su);
}

// ------------------------------------------------------------------------------------------------
// Helpers of the cooperative pivot kernel (k_blk_pivots_fused, peer.cuh), which runs up to `npiv` complete primal pivots of
// the blocked engine in ONE launch: the five dependent kernels of an iteration (price, select, column + ratios, ratio pick,
// row + slot) become phases of a persistent grid.  Every reduction carries (best, second best, index of the best): the tie
// folds of primal :271-286 / :379-399 are order-independent when the second best is at least 2 EPS away from the best (same
// criterion as the fast paths of k_select_primal / k_ratio_pick), so the winner is known without evaluating the fold;
// otherwise one block evaluates it (select_primal_body / ratio_pick_body).  Decisions are identical to the five-kernel path.
// ------------------------------------------------------------------------------------------------
struct Top2 {
    double a1, a2;
    int i1;
};
template <bool MAX> __device__ __forceinline__ bool top2_better(double a, double b) { return MAX ? (a > b) : (a < b); }
template <bool MAX> __device__ __forceinline__ void top2_push(Top2& t, double v, int i) {
    if (top2_better<MAX>(v, t.a1)) { t.a2 = t.a1; t.a1 = v; t.i1 = i; }
    else if (top2_better<MAX>(v, t.a2)) t.a2 = v;
}
template <bool MAX> __device__ __forceinline__ void top2_merge(Top2& t, const Top2& o) {
    if (top2_better<MAX>(o.a1, t.a1)) {
        const double second = top2_better<MAX>(t.a1, o.a2) ? t.a1 : o.a2;
        t.a1 = o.a1; t.i1 = o.i1; t.a2 = second;
    } else {
        // o.a1 does not beat t.a1: it competes for second place (also when it EQUALS t.a1 => a2 == a1 => "tie")
        if (top2_better<MAX>(o.a1, t.a2)) t.a2 = o.a1;
    }
}
__device__ __forceinline__ void blk_zero_slot(const DevLP& lp, int slot, int64_t t0, int64_t stride) {
    double* Uslot = lp.U + (int64_t)slot * lp.ld;
    double* Vslot = lp.V + (int64_t)slot * lp.ldv;
    const int64_t tmax = max(lp.ld, lp.ldv);
    for (int64_t t = t0; t < tmax; t += stride) {
        if (t < lp.ld) Uslot[t] = 0.;
        if (t < lp.ldv) Vslot[t] = 0.;
    }
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
        sU[j * kFlushSU + i] = (j < cnt && row0 + i < R) ? -U[(int64_t)j * ld + row0 + i] : 0.;
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
                if (c < C && r < R) {
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
                if (c < C && r < R) {
                    double* p = T + c * ld + r;
                    if (STREAM) st_f64x2_stream(p, acc[ct][rt]);
                    else st_f64x2(p, acc[ct][rt]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3b, version 3: k_blk_flush with one CTA per SM, TWO register tiles per warp (the 16 x 16-byte loads of column step s+1 are
// issued before the DMMA sequence of step s starts: 8 KB in flight per warp for the whole time it occupies the tensor pipe)
// and no block-wide barrier per column step.  The V tiles travel global -> shared
// memory as bulk asynchronous copies (cp.async.bulk, the TMA copy engine: one 512-byte row of the tile per copy) into
// a ring of kFlushStages buffers guarded by mbarriers: full[s] (armed with the expected byte count by the producer
// lane, completed by the copy engine) and empty[s] (one arrival per warp when it has read the tile).  Warps only
// meet through those mbarriers, so the two warps of an SM sub-partition drift out of phase and one of them occupies
// the fp64 tensor pipe while the other issues its tile loads / stores -- with a __syncthreads per step every warp hit
// the load/store phase at the same time and the tensor pipe idled (an intermediate version with a barrier per step: 69 % DMMA
// active at k = 64, profiles/r01_flush_kernel_sweeps.jsonl, flush_kernel = 2 rows).
// Requires V rows padded to a multiple of 64 columns (the engine allocates ldv that way) and 16-byte alignment.
// ------------------------------------------------------------------------------------------------
constexpr int kFlushStages = 3;
constexpr int kFlush3Threads = 288;  // 8 consumer warps (4 x 2 warp tiles of 32 x 32) + 1 producer warp
inline size_t blk_flush3_smem_bytes(int K4) { return sizeof(double) * (size_t)K4 * (kFlushSU + kFlushStages * kFlushSV) + 16 * kFlushStages; }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned spins = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0u) {  // a copy that never lands must not hang the GPU
            long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <bool STREAM>
__global__ void __launch_bounds__(kFlush3Threads, 1) k_blk_flush3(double* __restrict__ T, int64_t ld, int R, int C, const double* __restrict__ U,
                                                       const double* __restrict__ V, int64_t ldv, int cnt, int col_steps) {
    extern __shared__ __align__(16) double blk_smem[];
    const int K4 = (cnt + 3) & ~3;
    double* sU = blk_smem;                         // sU[j][row] = -U[row0 + row, j]
    double* sVr = blk_smem + K4 * kFlushSU;        // ring: sV[stage][j][col] = V[j, col0 + col]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(sVr + kFlushStages * K4 * kFlushSV);
    unsigned long long* empty = full + kFlushStages;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * kFlushRows;
    const int wr = (warp & 3) * 32, wc = (warp >> 2) * 32;
    const int fq = lane >> 2, fk = lane & 3;
    const int ksteps = K4 >> 2;
    const int64_t step0 = (int64_t)blockIdx.y * col_steps;
    const int64_t steps_total = (C + kFlushCols - 1) / kFlushCols;
    const int nsteps = (int)max((int64_t)0, min((int64_t)col_steps, steps_total - step0));
    if (nsteps == 0) return;
    const unsigned tile_bytes = (unsigned)cnt * kFlushCols * (unsigned)sizeof(double);

    // producer warp (warp 8): V tile of step t -> ring slot t % kFlushStages, one 512-byte row per lane and copy
    auto issue_v = [&](int t) {
        const int slot = t % kFlushStages;
        const int use = t / kFlushStages;
        if (lane == 0) {
            if (use > 0) mbar_wait(&empty[slot], (unsigned)((use - 1) & 1));  // every consumer warp released the previous tile of this slot
            mbar_arrive_expect_tx(&full[slot], tile_bytes);
        }
        __syncwarp();
        double* dst = sVr + (size_t)slot * K4 * kFlushSV;
        const double* src = V + (step0 + t) * kFlushCols;
        for (int j = lane; j < cnt; j += 32) bulk_g2s(dst + j * kFlushSV, src + (int64_t)j * ldv, kFlushCols * (unsigned)sizeof(double), &full[slot]);
    };
    auto load_tile = [&](double2 (&acc)[4][4], int s) {
        const int64_t col0 = (step0 + s) * kFlushCols;
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int64_t c = col0 + wc + ct * 8 + fq;
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                const int64_t r = row0 + wr + rt * 8 + 2 * fk;
                if (c < C && r < R) {
                    const double* p = T + c * ld + r;
                    acc[ct][rt] = STREAM ? ld_f64x2_stream(p) : ld_f64x2(p);
                } else {
                    acc[ct][rt] = make_double2(0., 0.);
                }
            }
        }
    };
    auto mma_store = [&](double2 (&acc)[4][4], int s) {
        const int slot = s % kFlushStages;
        mbar_wait(&full[slot], (unsigned)((s / kFlushStages) & 1));
        const double* sV = sVr + (size_t)slot * K4 * kFlushSV;
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
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);  // this warp no longer reads the slot
        const int64_t col0 = (step0 + s) * kFlushCols;
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int64_t c = col0 + wc + ct * 8 + fq;
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                const int64_t r = row0 + wr + rt * 8 + 2 * fk;
                if (c < C && r < R) {
                    double* p = T + c * ld + r;
                    if (STREAM) st_f64x2_stream(p, acc[ct][rt]);
                    else st_f64x2(p, acc[ct][rt]);
                }
            }
        }
    };

    if (tid == 0) {
        for (int i = 0; i < kFlushStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double2 accA[4][4], accB[4][4];
    if (warp < 8) load_tile(accA, 0);
    for (int e = tid; e < K4 * kFlushRows; e += kFlush3Threads) {
        const int j = e >> 7, i = e & (kFlushRows - 1);
        sU[j * kFlushSU + i] = (j < cnt && row0 + i < R) ? -U[(int64_t)j * ld + row0 + i] : 0.;
    }
    // rows cnt .. K4-1 of every ring slot are never copied: zero them once
    for (int e = tid; e < kFlushStages * (K4 - cnt) * kFlushSV; e += kFlush3Threads) {
        const int slot = e / ((K4 - cnt) * kFlushSV), rem = e - slot * (K4 - cnt) * kFlushSV;
        sVr[(size_t)slot * K4 * kFlushSV + (size_t)cnt * kFlushSV + rem] = 0.;
    }
    __syncthreads();  // barriers initialised, -U and the zero rows visible (the only block-wide barrier)
    if (warp == 8) {  // producer
        for (int t = 0; t < nsteps; ++t) issue_v(t);
        return;
    }
    for (int s = 0; s < nsteps; s += 2) {
        if (s + 1 < nsteps) load_tile(accB, s + 1);
        mma_store(accA, s);
        if (s + 1 >= nsteps) break;
        if (s + 2 < nsteps) load_tile(accA, s + 2);
        mma_store(accB, s + 1);
    }
}

// ------------------------------------------------------------------------------------------------
// K3b, version 4 (tensor-bound regime, k >= ~40): same ring / fragment layout as k_blk_flush3, but SIXTEEN consumer warps per
// SM (CTA tile 128 rows x 128 columns per step, warps 4 x 4, warp tile 32 x 32) with ONE register tile per warp.  ncu on
// version 3 showed 40 % of the warp samples in the fixed-latency wait behind each DMMA: a warp cannot issue DMMAs back to
// back, so two warps per scheduler leave the fp64 tensor pipe ~25 % idle; four warps per scheduler saturate it, and the
// tile loads of one warp hide behind the DMMA phases of the other three (no register prefetch needed: 112 registers).
// ------------------------------------------------------------------------------------------------
constexpr int kFlush4Cols = 128;
constexpr int kFlush4SV = kFlush4Cols + 4;
constexpr int kFlush4Threads = 544;  // 16 consumer warps + 1 producer warp
inline int blk_flush4_stages(int K4) { return K4 <= 40 ? 3 : 2; }
inline size_t blk_flush4_smem_bytes_st(int K4, int stages) { return sizeof(double) * (size_t)K4 * (kFlushSU + stages * kFlush4SV) + 16 * 3; }
inline size_t blk_flush4_smem_bytes(int K4) { return sizeof(double) * (size_t)K4 * (kFlushSU + blk_flush4_stages(K4) * kFlush4SV) + 16 * 3; }

// tile access modes of versions 4 / 5 (template parameter STREAM): 0 = default caching, 1 = evict-first loads and stores (.cs),
// 2 = loads that bypass L1 (.cg) + evict-first stores, 3 = .cg loads + .cg stores
template <int MODE> __device__ __forceinline__ double2 ld_tile(const double* p) {
    if (MODE == 1) return __ldcs(reinterpret_cast<const double2*>(p));
    if (MODE >= 2) return __ldcg(reinterpret_cast<const double2*>(p));
    return *reinterpret_cast<const double2*>(p);
}
template <int MODE> __device__ __forceinline__ void st_tile(double* p, double2 v) {
    if (MODE == 1 || MODE == 2) __stcs(reinterpret_cast<double2*>(p), v);
    else if (MODE == 3) __stcg(reinterpret_cast<double2*>(p), v);
    else *reinterpret_cast<double2*>(p) = v;
}

template <int STREAM>
__global__ void __launch_bounds__(kFlush4Threads, 1) k_blk_flush4(double* __restrict__ T, int64_t ld, int R, int C, const double* __restrict__ U,
                                                                  const double* __restrict__ V, int64_t ldv, int cnt, int col_steps, int stages) {
    extern __shared__ __align__(16) double blk_smem[];
    const int K4 = (cnt + 3) & ~3;
    double* sU = blk_smem;                         // sU[j][row] = -U[row0 + row, j]
    double* sVr = blk_smem + K4 * kFlushSU;        // ring: sV[stage][j][col] = V[j, col0 + col]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(sVr + (size_t)stages * K4 * kFlush4SV);
    unsigned long long* empty = full + 3;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * kFlushRows;
    const int wr = (warp & 3) * 32, wc = ((warp >> 2) & 3) * 32;
    const int fq = lane >> 2, fk = lane & 3;
    const int ksteps = K4 >> 2;
    const int64_t step0 = (int64_t)blockIdx.y * col_steps;
    const int64_t steps_total = (C + kFlush4Cols - 1) / kFlush4Cols;
    const int nsteps = (int)max((int64_t)0, min((int64_t)col_steps, steps_total - step0));
    if (nsteps == 0) return;
    const unsigned tile_bytes = (unsigned)cnt * kFlush4Cols * (unsigned)sizeof(double);
    if (tid == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < K4 * kFlushRows; e += kFlush4Threads) {
        const int j = e >> 7, i = e & (kFlushRows - 1);
        sU[j * kFlushSU + i] = (j < cnt && row0 + i < R) ? -U[(int64_t)j * ld + row0 + i] : 0.;
    }
    for (int e = tid; e < stages * (K4 - cnt) * kFlush4SV; e += kFlush4Threads) {  // rows cnt .. K4-1 are never copied: zero them once
        const int slot = e / ((K4 - cnt) * kFlush4SV), rem = e - slot * (K4 - cnt) * kFlush4SV;
        sVr[(size_t)slot * K4 * kFlush4SV + (size_t)cnt * kFlush4SV + rem] = 0.;
    }
    __syncthreads();  // the only block-wide barrier
    if (warp == 16) {  // producer: V tile of step t -> ring slot t % stages, one 1 KB row per lane and copy
        for (int t = 0; t < nsteps; ++t) {
            const int slot = t % stages, use = t / stages;
            if (lane == 0) {
                if (use > 0) mbar_wait(&empty[slot], (unsigned)((use - 1) & 1));
                mbar_arrive_expect_tx(&full[slot], tile_bytes);
            }
            __syncwarp();
            double* dst = sVr + (size_t)slot * K4 * kFlush4SV;
            const double* src = V + (step0 + t) * kFlush4Cols;
            for (int j = lane; j < cnt; j += 32) bulk_g2s(dst + j * kFlush4SV, src + (int64_t)j * ldv, kFlush4Cols * (unsigned)sizeof(double), &full[slot]);
        }
        return;
    }
    for (int s = 0; s < nsteps; ++s) {
        const int64_t col0 = (step0 + s) * kFlush4Cols;
        double2 acc[4][4];
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int64_t c = col0 + wc + ct * 8 + fq;
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                const int64_t r = row0 + wr + rt * 8 + 2 * fk;
                if (c < C && r < R) {
                    const double* p = T + c * ld + r;
                    acc[ct][rt] = ld_tile<STREAM>(p);
                } else {
                    acc[ct][rt] = make_double2(0., 0.);
                }
            }
        }
        const int slot = s % stages;
        mbar_wait(&full[slot], (unsigned)((s / stages) & 1));
        const double* sV = sVr + (size_t)slot * K4 * kFlush4SV;
#pragma unroll 2
        for (int ks = 0; ks < ksteps; ++ks) {
            double a[4], b[4];
            const int j = ks * 4 + fk;
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) a[ct] = sV[j * kFlush4SV + wc + ct * 8 + fq];
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) b[rt] = sU[j * kFlushSU + wr + rt * 8 + fq];
#pragma unroll
            for (int ct = 0; ct < 4; ++ct)
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) dmma_m8n8k4(acc[ct][rt].x, acc[ct][rt].y, a[ct], b[rt]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int64_t c = col0 + wc + ct * 8 + fq;
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                const int64_t r = row0 + wr + rt * 8 + 2 * fk;
                if (c < C && r < R) {
                    double* p = T + c * ld + r;
                    st_tile<STREAM>(p, acc[ct][rt]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3b, version 5: k_blk_flush4 with the warp's 32 x 32 register tile cut into NSPLIT parts by column tiles and software-pipelined
// inside the warp.  ncu on version 4 (profiles/r02_ncu_full_blk_flush4_k56_summary.txt): 75.6 % DMMA active, and the largest stall
// is long_scoreboard (5.3 warps per issue cycle): a warp issues its 16 tile loads and waits for ALL of them before its first DMMA,
// so a scheduler regularly has fewer than the ~3 warps in their DMMA phase that the fp64 tensor pipe needs.  Here part p of step
// s + 1 is requested right after part p of step s was stored, and the DMMAs of the OTHER parts of step s run while it travels:
// no warp ever waits on HBM with nothing to issue.  Same registers (one 32 x 32 tile per warp), same ring, same per-element
// accumulation order as versions 3 / 4 (the stored tableau is bit-identical to theirs).
// ------------------------------------------------------------------------------------------------
template <int STREAM, int NSPLIT>
__global__ void __launch_bounds__(kFlush4Threads, 1) k_blk_flush5(double* __restrict__ T, int64_t ld, int R, int C, const double* __restrict__ U,
                                                                  const double* __restrict__ V, int64_t ldv, int cnt, int col_steps, int stages) {
    constexpr int CTS = 4 / NSPLIT;  // column tiles (of 8 columns) per part
    extern __shared__ __align__(16) double blk_smem[];
    const int K4 = (cnt + 3) & ~3;
    double* sU = blk_smem;                         // sU[j][row] = -U[row0 + row, j]
    double* sVr = blk_smem + K4 * kFlushSU;        // ring: sV[stage][j][col] = V[j, col0 + col]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(sVr + (size_t)stages * K4 * kFlush4SV);
    unsigned long long* empty = full + 3;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * kFlushRows;
    const int wr = (warp & 3) * 32, wc = ((warp >> 2) & 3) * 32;
    const int fq = lane >> 2, fk = lane & 3;
    const int ksteps = K4 >> 2;
    const int64_t step0 = (int64_t)blockIdx.y * col_steps;
    const int64_t steps_total = (C + kFlush4Cols - 1) / kFlush4Cols;
    const int nsteps = (int)max((int64_t)0, min((int64_t)col_steps, steps_total - step0));
    if (nsteps == 0) return;
    const unsigned tile_bytes = (unsigned)cnt * kFlush4Cols * (unsigned)sizeof(double);
    if (tid == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double2 acc[NSPLIT][CTS][4];  // [part][column tile][row tile]
    auto load_part = [&](double2 (&a)[CTS][4], int s, int part) {
        const int64_t col0 = (step0 + s) * kFlush4Cols + wc + part * (CTS * 8);
#pragma unroll
        for (int ct = 0; ct < CTS; ++ct) {
            const int64_t c = col0 + ct * 8 + fq;
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                const int64_t r = row0 + wr + rt * 8 + 2 * fk;
                if (c < C && r < R) {
                    const double* p = T + c * ld + r;
                    a[ct][rt] = ld_tile<STREAM>(p);
                } else {
                    a[ct][rt] = make_double2(0., 0.);
                }
            }
        }
    };
    auto store_part = [&](const double2 (&a)[CTS][4], int s, int part) {
        const int64_t col0 = (step0 + s) * kFlush4Cols + wc + part * (CTS * 8);
#pragma unroll
        for (int ct = 0; ct < CTS; ++ct) {
            const int64_t c = col0 + ct * 8 + fq;
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                const int64_t r = row0 + wr + rt * 8 + 2 * fk;
                if (c < C && r < R) {
                    double* p = T + c * ld + r;
                    st_tile<STREAM>(p, a[ct][rt]);
                }
            }
        }
    };
    if (warp < 16) {  // the first tile travels while -U is staged
#pragma unroll
        for (int p = 0; p < NSPLIT; ++p) load_part(acc[p], 0, p);
    }
    for (int e = tid; e < K4 * kFlushRows; e += kFlush4Threads) {
        const int j = e >> 7, i = e & (kFlushRows - 1);
        sU[j * kFlushSU + i] = (j < cnt && row0 + i < R) ? -U[(int64_t)j * ld + row0 + i] : 0.;
    }
    for (int e = tid; e < stages * (K4 - cnt) * kFlush4SV; e += kFlush4Threads) {  // rows cnt .. K4-1 are never copied: zero them once
        const int slot = e / ((K4 - cnt) * kFlush4SV), rem = e - slot * (K4 - cnt) * kFlush4SV;
        sVr[(size_t)slot * K4 * kFlush4SV + (size_t)cnt * kFlush4SV + rem] = 0.;
    }
    __syncthreads();  // the only block-wide barrier
    if (warp == 16) {  // producer: V tile of step t -> ring slot t % stages, one 1 KB row per lane and copy
        for (int t = 0; t < nsteps; ++t) {
            const int slot = t % stages, use = t / stages;
            if (lane == 0) {
                if (use > 0) mbar_wait(&empty[slot], (unsigned)((use - 1) & 1));
                mbar_arrive_expect_tx(&full[slot], tile_bytes);
            }
            __syncwarp();
            double* dst = sVr + (size_t)slot * K4 * kFlush4SV;
            const double* src = V + (step0 + t) * kFlush4Cols;
            for (int j = lane; j < cnt; j += 32) bulk_g2s(dst + j * kFlush4SV, src + (int64_t)j * ldv, kFlush4Cols * (unsigned)sizeof(double), &full[slot]);
        }
        return;
    }
    for (int s = 0; s < nsteps; ++s) {
        const int slot = s % stages;
        mbar_wait(&full[slot], (unsigned)((s / stages) & 1));
        const double* sV = sVr + (size_t)slot * K4 * kFlush4SV;
#pragma unroll
        for (int p = 0; p < NSPLIT; ++p) {
#pragma unroll 2
            for (int ks = 0; ks < ksteps; ++ks) {
                double a[CTS], b[4];
                const int j = ks * 4 + fk;
#pragma unroll
                for (int ct = 0; ct < CTS; ++ct) a[ct] = sV[j * kFlush4SV + wc + (p * CTS + ct) * 8 + fq];
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) b[rt] = sU[j * kFlushSU + wr + rt * 8 + fq];
#pragma unroll
                for (int ct = 0; ct < CTS; ++ct)
#pragma unroll
                    for (int rt = 0; rt < 4; ++rt) dmma_m8n8k4(acc[p][ct][rt].x, acc[p][ct][rt].y, a[ct], b[rt]);
            }
            if (p == NSPLIT - 1) {  // this warp no longer reads the slot
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
            }
            store_part(acc[p], s, p);
            if (s + 1 < nsteps) load_part(acc[p], s + 1, p);  // lands behind the DMMAs of the other parts
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3b, version 6: the 16-consumer-warp kernel WITHOUT a producer warp.  512 threads instead of 544 lift the register cap from 96 to
// 128 per thread: versions 4 / 5 spill (56 - 80 bytes), and their spill reloads (LDL) sit on the same scoreboards as the
// outstanding tile loads -- on ncu's source page of version 4 long_scoreboard is 28 % of all stall samples, two thirds of them on integer
// instructions that only consume a reloaded spill (profiles/r02_ncu_full_blk_flush4_k56_summary.txt), i.e. a warp waits for its
// whole T tile before it even polls the ring, and in version 5 the reloads inside the pipelined loop cancel the overlap.
// The ring is refilled by whichever consumer warp is the LAST to finish the DMMAs of a step (a shared-memory arrival counter per
// slot replaces the `empty` mbarriers): that warp re-arms full[slot] and issues the bulk copies of step s + stages.
// NSPLIT = 1: one 32 x 32 register tile per warp, loaded as a whole (version 4's schedule); NSPLIT = 2: tile pipelined in two
// halves (version 5's schedule).  Interior tiles take a path without bounds predicates.
// Same DMMA shape and per-element accumulation order as every other version (bit-identical tableau).
// ------------------------------------------------------------------------------------------------
constexpr int kFlush6Threads = 512;
inline size_t blk_flush6_smem_bytes(int K4, int stages) { return sizeof(double) * (size_t)K4 * (kFlushSU + stages * kFlush4SV) + 48; }

template <int STREAM, int NSPLIT, bool FULL>
__device__ __forceinline__ void flush6_body(double* __restrict__ T, int64_t ld, int R, int C, const double* __restrict__ V, int64_t ldv, int cnt,
                                            int K4, int stages, int nsteps, int64_t step0, int64_t row0, const double* sU, double* sVr,
                                            unsigned long long* full, unsigned* arrived) {
    constexpr int CTS = 4 / NSPLIT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = (warp & 3) * 32, wc = (warp >> 2) * 32;
    const int fq = lane >> 2, fk = lane & 3;
    const int ksteps = K4 >> 2;
    const unsigned tile_bytes = (unsigned)cnt * kFlush4Cols * (unsigned)sizeof(double);
    const int64_t r_lane = row0 + wr + 2 * fk;                    // first row of this lane's accumulator pairs
    double* tp = T + (step0 * kFlush4Cols + wc + fq) * ld + r_lane;  // element (row r_lane, column col0 + wc + fq) of step 0
    const int64_t cstride = 8 * ld, sstride = (int64_t)kFlush4Cols * ld;
    double2 acc[NSPLIT][CTS][4];  // [part][column tile][row tile]
    auto load_part = [&](double2 (&a)[CTS][4], int s, int part) {
        const double* base = tp + (int64_t)s * sstride + (int64_t)(part * CTS) * cstride;
#pragma unroll
        for (int ct = 0; ct < CTS; ++ct) {
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                if (FULL) {
                    a[ct][rt] = ld_tile<STREAM>(base + ct * cstride + rt * 8);
                } else {
                    const int64_t c = (step0 + s) * kFlush4Cols + wc + (part * CTS + ct) * 8 + fq, r = r_lane + rt * 8;
                    a[ct][rt] = (c < C && r < R) ? ld_tile<STREAM>(base + ct * cstride + rt * 8) : make_double2(0., 0.);
                }
            }
        }
    };
    auto store_part = [&](const double2 (&a)[CTS][4], int s, int part) {
        double* base = tp + (int64_t)s * sstride + (int64_t)(part * CTS) * cstride;
#pragma unroll
        for (int ct = 0; ct < CTS; ++ct) {
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                if (FULL) {
                    st_tile<STREAM>(base + ct * cstride + rt * 8, a[ct][rt]);
                } else {
                    const int64_t c = (step0 + s) * kFlush4Cols + wc + (part * CTS + ct) * 8 + fq, r = r_lane + rt * 8;
                    if (c < C && r < R) st_tile<STREAM>(base + ct * cstride + rt * 8, a[ct][rt]);
                }
            }
        }
    };
#pragma unroll
    for (int p = 0; p < NSPLIT; ++p) load_part(acc[p], 0, p);
    for (int s = 0; s < nsteps; ++s) {
        const int slot = s % stages;
        mbar_wait(&full[slot], (unsigned)((s / stages) & 1));
        const double* sV = sVr + (size_t)slot * K4 * kFlush4SV;
#pragma unroll
        for (int p = 0; p < NSPLIT; ++p) {
#pragma unroll 2
            for (int ks = 0; ks < ksteps; ++ks) {
                double a[CTS], b[4];
                const int j = ks * 4 + fk;
#pragma unroll
                for (int ct = 0; ct < CTS; ++ct) a[ct] = sV[j * kFlush4SV + wc + (p * CTS + ct) * 8 + fq];
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) b[rt] = sU[j * kFlushSU + wr + rt * 8 + fq];
#pragma unroll
                for (int ct = 0; ct < CTS; ++ct)
#pragma unroll
                    for (int rt = 0; rt < 4; ++rt) dmma_m8n8k4(acc[p][ct][rt].x, acc[p][ct][rt].y, a[ct], b[rt]);
            }
            if (p == NSPLIT - 1) {
                // this warp no longer reads the slot; the last of the 16 warps to get here refills it with the tile of step s + stages
                __syncwarp();
                unsigned last = 0;
                if (lane == 0) {
                    __threadfence_block();
                    last = (atomicAdd(&arrived[slot], 1u) == 15u) ? 1u : 0u;
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                const int t = s + stages;
                if (last && t < nsteps) {
                    // every lane issues its rows: a single-thread loop was measured 4 % slower (the refill sits on the critical path: with
                    // two stages the tile of step s + 2 is requested when the slowest warp leaves step s)
                    if (lane == 0) {
                        arrived[slot] = 0u;  // nobody arrives on this slot again before the copies below have completed full[slot]
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_arrive_expect_tx(&full[slot], tile_bytes);
                    }
                    __syncwarp();
                    double* dst = sVr + (size_t)slot * K4 * kFlush4SV;
                    const double* src = V + (step0 + t) * kFlush4Cols;
                    for (int j = lane; j < cnt; j += 32) bulk_g2s(dst + j * kFlush4SV, src + (int64_t)j * ldv, kFlush4Cols * (unsigned)sizeof(double), &full[slot]);
                }
            }
            store_part(acc[p], s, p);
            if (s + 1 < nsteps) load_part(acc[p], s + 1, p);
        }
    }
}

template <int STREAM, int NSPLIT>
__global__ void __launch_bounds__(kFlush6Threads, 1) k_blk_flush6(double* __restrict__ T, int64_t ld, int R, int C, const double* __restrict__ U,
                                                                  const double* __restrict__ V, int64_t ldv, int cnt, int col_steps, int stages) {
    extern __shared__ __align__(16) double blk_smem[];
    const int K4 = (cnt + 3) & ~3;
    double* sU = blk_smem;                         // sU[j][row] = -U[row0 + row, j]
    double* sVr = blk_smem + K4 * kFlushSU;        // ring: sV[stage][j][col] = V[j, col0 + col]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(sVr + (size_t)stages * K4 * kFlush4SV);
    unsigned* arrived = reinterpret_cast<unsigned*>(full + 3);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * kFlushRows;
    const int64_t step0 = (int64_t)blockIdx.y * col_steps;
    const int64_t steps_total = (C + kFlush4Cols - 1) / kFlush4Cols;
    const int nsteps = (int)max((int64_t)0, min((int64_t)col_steps, steps_total - step0));
    if (nsteps == 0) return;
    if (tid == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); arrived[i] = 0u; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < K4 * kFlushRows; e += kFlush6Threads) {
        const int j = e >> 7, i = e & (kFlushRows - 1);
        sU[j * kFlushSU + i] = (j < cnt && row0 + i < R) ? -U[(int64_t)j * ld + row0 + i] : 0.;
    }
    for (int e = tid; e < stages * (K4 - cnt) * kFlush4SV; e += kFlush6Threads) {  // rows cnt .. K4-1 are never copied: zero them once
        const int slot = e / ((K4 - cnt) * kFlush4SV), rem = e - slot * (K4 - cnt) * kFlush4SV;
        sVr[(size_t)slot * K4 * kFlush4SV + (size_t)cnt * kFlush4SV + rem] = 0.;
    }
    __syncthreads();  // the only block-wide barrier
    if (warp == 0) {  // initial fill of the ring
        const unsigned tile_bytes = (unsigned)cnt * kFlush4Cols * (unsigned)sizeof(double);
        for (int t = 0; t < stages && t < nsteps; ++t) {
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive_expect_tx(&full[t], tile_bytes);
            }
            __syncwarp();
            double* dst = sVr + (size_t)t * K4 * kFlush4SV;
            const double* src = V + (step0 + t) * kFlush4Cols;
            for (int j = lane; j < cnt; j += 32) bulk_g2s(dst + j * kFlush4SV, src + (int64_t)j * ldv, kFlush4Cols * (unsigned)sizeof(double), &full[t]);
        }
    }
    const bool interior = row0 + kFlushRows <= R && (step0 + nsteps) * kFlush4Cols <= C;
    if (interior) flush6_body<STREAM, NSPLIT, true>(T, ld, R, C, V, ldv, cnt, K4, stages, nsteps, step0, row0, sU, sVr, full, arrived);
    else flush6_body<STREAM, NSPLIT, false>(T, ld, R, C, V, ldv, cnt, K4, stages, nsteps, step0, row0, sU, sVr, full, arrived);
}

// ------------------------------------------------------------------------------------------------
// K3b, version 4r: version 4 with RW x 4 consumer warps (CTA tile RW*32 rows x 128 columns per step).  RW = 3 gives 12 consumer
// warps + the producer warp = 13 warps, which the register file allocates as 16: 128 registers per thread instead of 96 (no
// spills, fragment loads of the next k-step in flight behind the DMMAs of the current one), and 96 KB of tile loads in flight
// instead of 128 KB next to a larger L1 (163 KB of shared memory at k = 56).  Keeps the dedicated producer warp of version 4 (the
// refill stays off the consumers' critical path, which is what version 6 loses).  Bit-identical to the other versions.
// ------------------------------------------------------------------------------------------------
template <int RW> constexpr int flush4r_threads() { return (4 * RW + 1) * 32; }
template <int RW> inline size_t blk_flush4r_smem_bytes(int K4, int stages) { return sizeof(double) * (size_t)K4 * ((RW * 32 + 4) + stages * kFlush4SV) + 48; }

template <int STREAM, int RW>
__global__ void __launch_bounds__((4 * RW + 1) * 32, 1) k_blk_flush4r(double* __restrict__ T, int64_t ld, int R, int C, const double* __restrict__ U,
                                                                      const double* __restrict__ V, int64_t ldv, int cnt, int col_steps, int stages) {
    constexpr int ROWS = RW * 32, SU = ROWS + 4, NCW = 4 * RW, THREADS = (NCW + 1) * 32;
    extern __shared__ __align__(16) double blk_smem[];
    const int K4 = (cnt + 3) & ~3;
    double* sU = blk_smem;                    // sU[j][row] = -U[row0 + row, j]
    double* sVr = blk_smem + K4 * SU;         // ring: sV[stage][j][col] = V[j, col0 + col]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(sVr + (size_t)stages * K4 * kFlush4SV);
    unsigned long long* empty = full + 3;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * ROWS;
    const int wr = (warp % RW) * 32, wc = (warp / RW) * 32;
    const int fq = lane >> 2, fk = lane & 3;
    const int ksteps = K4 >> 2;
    const int64_t step0 = (int64_t)blockIdx.y * col_steps;
    const int64_t steps_total = (C + kFlush4Cols - 1) / kFlush4Cols;
    const int nsteps = (int)max((int64_t)0, min((int64_t)col_steps, steps_total - step0));
    if (nsteps == 0) return;
    const unsigned tile_bytes = (unsigned)cnt * kFlush4Cols * (unsigned)sizeof(double);
    if (tid == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < K4 * ROWS; e += THREADS) {
        const int j = e / ROWS, i = e - j * ROWS;
        sU[j * SU + i] = (j < cnt && row0 + i < R) ? -U[(int64_t)j * ld + row0 + i] : 0.;
    }
    for (int e = tid; e < stages * (K4 - cnt) * kFlush4SV; e += THREADS) {  // rows cnt .. K4-1 are never copied: zero them once
        const int slot = e / ((K4 - cnt) * kFlush4SV), rem = e - slot * (K4 - cnt) * kFlush4SV;
        sVr[(size_t)slot * K4 * kFlush4SV + (size_t)cnt * kFlush4SV + rem] = 0.;
    }
    __syncthreads();  // the only block-wide barrier
    if (warp == NCW) {  // producer: V tile of step t -> ring slot t % stages, one 1 KB row per lane and copy
        for (int t = 0; t < nsteps; ++t) {
            const int slot = t % stages, use = t / stages;
            if (lane == 0) {
                if (use > 0) mbar_wait(&empty[slot], (unsigned)((use - 1) & 1));
                mbar_arrive_expect_tx(&full[slot], tile_bytes);
            }
            __syncwarp();
            double* dst = sVr + (size_t)slot * K4 * kFlush4SV;
            const double* src = V + (step0 + t) * kFlush4Cols;
            for (int j = lane; j < cnt; j += 32) bulk_g2s(dst + j * kFlush4SV, src + (int64_t)j * ldv, kFlush4Cols * (unsigned)sizeof(double), &full[slot]);
        }
        return;
    }
    const bool interior = row0 + ROWS <= R && (step0 + nsteps) * kFlush4Cols <= C;
    const int64_t r_lane = row0 + wr + 2 * fk;
    double* tp = T + (step0 * kFlush4Cols + wc + fq) * ld + r_lane;
    const int64_t cstride = 8 * ld, sstride = (int64_t)kFlush4Cols * ld;
    for (int s = 0; s < nsteps; ++s) {
        double2 acc[4][4];
        double* base = tp + (int64_t)s * sstride;
        if (interior) {
#pragma unroll
            for (int ct = 0; ct < 4; ++ct)
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) acc[ct][rt] = ld_tile<STREAM>(base + ct * cstride + rt * 8);
        } else {
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) {
                const int64_t c = (step0 + s) * kFlush4Cols + wc + ct * 8 + fq;
#pragma unroll
                for (int rt = 0; rt < 4; ++rt)
                    acc[ct][rt] = (c < C && r_lane + rt * 8 < R) ? ld_tile<STREAM>(base + ct * cstride + rt * 8) : make_double2(0., 0.);
            }
        }
        const int slot = s % stages;
        mbar_wait(&full[slot], (unsigned)((s / stages) & 1));
        const double* sV = sVr + (size_t)slot * K4 * kFlush4SV;
#pragma unroll 2
        for (int ks = 0; ks < ksteps; ++ks) {
            double a[4], b[4];
            const int j = ks * 4 + fk;
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) a[ct] = sV[j * kFlush4SV + wc + ct * 8 + fq];
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) b[rt] = sU[j * SU + wr + rt * 8 + fq];
#pragma unroll
            for (int ct = 0; ct < 4; ++ct)
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) dmma_m8n8k4(acc[ct][rt].x, acc[ct][rt].y, a[ct], b[rt]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (interior) {
#pragma unroll
            for (int ct = 0; ct < 4; ++ct)
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) st_tile<STREAM>(base + ct * cstride + rt * 8, acc[ct][rt]);
        } else {
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) {
                const int64_t c = (step0 + s) * kFlush4Cols + wc + ct * 8 + fq;
#pragma unroll
                for (int rt = 0; rt < 4; ++rt)
                    if (c < C && r_lane + rt * 8 < R) st_tile<STREAM>(base + ct * cstride + rt * 8, acc[ct][rt]);
            }
        }
    }
}

}  // namespace ellp
