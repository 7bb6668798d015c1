// dual_blocked.cuh -- the DUAL simplex loop on the blocked condensed tableau (single GPU and peer-sharded).
//
// Reference: DualSimplexSolver::solve_with_initial, src/solvers/dual/dual_simplex_solver.rs:188-334.  The reference
// recomputes an LU of A_B in every iteration (:241) to obtain rho = e_r^T B^-1 (:248-253), the pivot row
// alpha = A_N^T rho (:255) and the entering column alpha_q = B^-1 a_q (:294).  On the tableau T = B^-1 A_N both are simply
// row r and column q of the CURRENT tableau, i.e. of  T_stale - U V  (blocked.cuh): the revised engine's three HBM passes
// per pivot (A_N^T rho, B^-1 a_q, rank-1 on B^-1) become one strided row read + one column read + k pending
// corrections, and every k pivots one rank-k flush on the fp64 tensor pipe.
//
// Layout = peer.cuh: rank g of G stores the nonbasic POSITIONS [pos_lo, pos_lo + nT) of T, V and dj (dj[t] = d of the
// variable at local position t, dual :296-302); x, Bv, Nv, Ns, U and PivotState are replicated and evolve identically.
//
// Per pivot (slot = index of the new pending (U, V) pair), k_blk_dual_pivots_fused:
//   L   leaving row = FIRST basis position whose variable violates a bound by more than EPS (:200-236): every thread
//       checks the rows whose x it just updated (phase C of the previous pivot), per-block minimum position + delta,
//       published as LL words; gathering them is the first grid-wide exchange (ll_publish / ll_gather, peer.cuh).
//       x is replicated, so no cross-rank message is needed.
//   R   pivot row of the current tableau over the local positions: alpha_t = T[r,t] - sum_j U[r,j] V[j,t] (one fma per
//       pending pivot, in pivot order), eligibility by the nonbasic side (:257-269), ratio d_t / alpha~_t, lexicographic
//       minimum of (ratio, position) == Iterator::min_by's first minimum (:270-279), carried with alpha_t of the winner.
//       Second grid-wide exchange; one block per rank then stores the rank's result into every rank's mailbox and every
//       block merges the G entries: the same entering position, theta_dual and pivot element on every rank.
//   E   local: new V slot (pivot row / alpha_q[r]), d_t -= theta_dual alpha_t (:298-300), position q handed to the
//       leaving variable (d = -theta_dual, :296; stored column := e_r).  Needs nothing from the column: it runs while the
//       column is in flight.
//   C   the owner of q rebuilds the entering column (stale column + pending corrections) and stores it into every other
//       rank's column buffer; every rank: x_B -= theta_primal alpha_q (:306-312), x_q += theta_primal (:314), index swap,
//       objective, trace (:316-333) by the thread that owns row r, new U slot, and the bound check of every updated row
//       (phase L of the next pivot).
// The pivot element read from the row (phase R) and from the column (phase C) are the same bits: both are T[r,q] with the
// pending products -U[r,j] V[j,q] applied in the same order.
// y is not needed by the iteration; it is rebuilt at download time from d (engine.cu: dual_tab_export).
#pragma once
#include "peer.cuh"

namespace ellp {

constexpr double kDualSnapUlps = 16.;

struct LexAcc {
    double v, al;  // ratio, pivot-row entry of the candidate
    int p;         // global nonbasic position, -1 = no candidate
    int nan;
};
struct LexSh {
    unsigned long long k[2][32];
    double al[2][32];
    int p[2][32], nan[2][32];
};

__device__ __forceinline__ LexAcc warp_lexmin(const LexAcc& a) {
    const unsigned full = 0xffffffffu;
    const unsigned long long key = (a.p >= 0) ? ord_key(a.v) : ~0ull;
    const unsigned long long M = warp_best_u64<false>(key);
    const int pc = (a.p >= 0 && key == M) ? a.p : 0x7fffffff;
    const int pm = __reduce_min_sync(full, pc);
    const int src = __ffs(__ballot_sync(full, pc == pm)) - 1;
    LexAcc r;
    r.p = (pm == 0x7fffffff) ? -1 : pm;
    r.v = ord_val(M);
    r.al = __shfl_sync(full, a.al, src);
    r.nan = (int)__reduce_or_sync(full, (unsigned)a.nan);
    return r;
}
// result valid in every thread; `buf` alternates between calls (block-uniform)
__device__ __forceinline__ LexAcc block_lexmin(const LexAcc& a, LexSh* sh, int& buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const LexAcc w = warp_lexmin(a);
    if (lane == 0) { sh->k[buf][warp] = (w.p >= 0) ? ord_key(w.v) : ~0ull; sh->al[buf][warp] = w.al; sh->p[buf][warp] = w.p; sh->nan[buf][warp] = w.nan; }
    __syncthreads();
    LexAcc r;
    r.p = (lane < nw) ? sh->p[buf][lane] : -1;
    r.v = (lane < nw && r.p >= 0) ? ord_val(sh->k[buf][lane]) : 0.;
    r.al = (lane < nw) ? sh->al[buf][lane] : 0.;
    r.nan = (lane < nw) ? sh->nan[buf][lane] : 0;
    buf ^= 1;
    return warp_lexmin(r);
}
__device__ __forceinline__ void lex_push(LexAcc& b, double v, int p, double al) {
    if (b.p < 0 || v < b.v || (v == b.v && p < b.p)) { b.v = v; b.p = p; b.al = al; }
}

// dual :200-236 for one basic variable: true when it violates a bound by more than EPS; delta = x - violated bound
__device__ __forceinline__ bool dual_violation(int kind, double lb, double ub, double x, double* delta) {
    if (kind == ELLP_LOWER) {
        if (x < lb - kEps) { *delta = x - lb; return true; }
    } else if (kind == ELLP_UPPER) {
        if (x > ub + kEps) { *delta = x - ub; return true; }
    } else if (kind == ELLP_TWOSIDED) {
        if (x > ub + kEps) { *delta = x - ub; return true; }
        if (x < lb - kEps) { *delta = x - lb; return true; }
    }
    return false;  // Free and Fixed basic variables never leave (:204, :232)
}

// DEVEX = true: leaving row = argmax infeasibility^2 / w_i with dual Devex reference weights (ELLP_PRICE_DEVEX; no reference
// counterpart: ellp picks the first infeasible row): w starts at 1 (reference framework = the starting basis), and after a pivot
// on (r, q): w_i = max(w_i, (alpha_q[i] / alpha_q[r])^2 w_r) for i != r, w_r = max(w_r / alpha_q[r]^2, 1) -- all from the
// entering column every row thread holds anyway, so the rule costs no extra pass and no extra exchange.
template <bool DEVEX>
__global__ void __launch_bounds__(kFusedMaxThreads, 1) k_blk_dual_pivots_fused(DevLP lp, PeerLinks pl, int slot0, int npiv, uint32_t seq0, PivotState* st) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ Top2Fast s_top;
    __shared__ LexSh s_lex;
    __shared__ double s_vec[kBlkMax];
    __shared__ double s_su[kBlkMax];
    __shared__ double s_mb[kMaxPeers * kMboxFields];
    __shared__ double s_part[kLLMaxBlocks * 4];
    __shared__ double s_extra;
    __shared__ double s_delta;
    const int tid = threadIdx.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + tid, gsize = (int64_t)gridDim.x * blockDim.x;
    const int G = gridDim.x, R = pl.nranks, me = pl.rank;
    uint4* llA = reinterpret_cast<uint4*>(lp.coop);         // leaving-row partials of every block, [parity][block][4 words]
    uint4* llC = reinterpret_cast<uint4*>(lp.coop + 2560);  // entering-position partials
    const int nT = lp.nT, m = lp.m;
    bool run = (__ldcg(&st->status) == kRunning);
    PivotRegs g;
    pivot_regs_load(g, st);
    int rbuf = 0, lbuf = 0;
    bool have_L = false;  // this block's leaving-row partial of the coming pivot was already published by phase C
    int zq = -1, zcnt = 0;  // local position handed over by the previous pivot: its V entries of the slots [0, zcnt) are still to be cleared
    for (int slot = slot0; slot < slot0 + npiv; ++slot) {
        if (!run) { blk_zero_slot(lp, slot, gtid, gsize); continue; }
        const uint32_t seq = seq0 + (uint32_t)(slot - slot0) + 1u;
        const int par = (int)(seq & 1u);
        // ---- L (first pivot of the launch only; afterwards phase C did it): first infeasible basis position (dual :200-236)
        if (!have_L) {
            Top2 t{DEVEX ? -1.0 : CUDART_INF, DEVEX ? -1.0 : CUDART_INF, -1};
            double my_delta = 0.;
            for (int64_t i = gtid; i < m; i += gsize) {
                const int var = __ldcg(lp.Bv + i);
                double dl;
                if (dual_violation(lp.kind[var], lp.lb[var], lp.ub[var], __ldcg(lp.x + var), &dl)) {
                    if (DEVEX) {
                        const double score = dl * dl / __ldcg(lp.w + i);
                        if (score > t.a1) { t.a1 = score; t.i1 = (int)i; my_delta = dl; }
                    } else { t.a1 = (double)i; t.i1 = (int)i; my_delta = dl; break; }
                }
            }
            const int mine = t.i1;
            t = top2_block_fast<DEVEX>(t, &s_top, rbuf);
            if (mine >= 0 && mine == t.i1) s_delta = my_delta;
            __syncthreads();
            ll_publish(llA, par, t, (t.i1 >= 0) ? s_delta : 0., seq);
        }
        double delta;
        ll_gather(llA, par, G, s_part, seq);
        const Top2 L = (G <= 32) ? ll_reduce_w<DEVEX>(s_part, G, &delta) : ll_reduce<DEVEX>(s_part, G, &s_top, rbuf, &s_extra, &delta);
        if (L.i1 < 0) {  // no infeasible basic variable: optimal (:243-246)
            if (gtid == 0) { st->status = ELLP_OPTIMAL; st->do_update = 0; st->do_step = 0; }
            run = false;
            blk_zero_slot(lp, slot, gtid, gsize);
            continue;
        }
        const int r = L.i1;
        const double w_r = DEVEX ? __ldcg(lp.w + r) : 1.;
        const bool neg = delta < 0.;
        const int new_side = neg ? ELLP_NB_LOWER : ELLP_NB_UPPER;
        // ---- R: pivot row over the local positions, ratios, lexicographic minimum (dual :255-279)
        if (tid < slot) s_su[tid] = __ldcg(lp.U + (int64_t)tid * lp.ld + r);
        __syncthreads();
        LexAcc best{0., 0., -1, 0};
        double a_first = 0.;  // pivot-row entry of the first position this thread owns (t == gtid), reused by phase E
        for (int64_t t = gtid; t < nT; t += gsize) {
            if ((int)t == zq) {  // handed over by the previous pivot: no pending correction before that pivot's slot
                for (int j = 0; j < zcnt; ++j) lp.V[(int64_t)j * lp.ldv + t] = 0.;
            }
            const double a_raw = corr_chain(__ldcg(lp.T + t * lp.ld + r), lp.V + t, lp.ldv, s_su, slot);
            if (t == gtid) a_first = a_raw; else lp.rN[t] = a_raw;
            const double a = neg ? -a_raw : a_raw;  // :257-259
            const int side = (int)__ldcg(lp.Ns + lp.pos_lo + t);
            const bool keep = (side == ELLP_NB_LOWER) ? (a > kEps) : (side == ELLP_NB_UPPER ? (a < -kEps) : true);
            if (keep) {
                double q = __ldcg(lp.dj + t) / a;
                if (q != q) best.nan = 1;       // partial_cmp().unwrap() would panic (:279)
                if (q == 0.) q = 0.;            // -0.0 and +0.0 compare equal in the reference: one key for both
                lex_push(best, q, lp.pos_lo + (int)t, a_raw);
            }
        }
        zq = -1;
        {
            const LexAcc b = block_lexmin(best, &s_lex, lbuf);
            const Top2 pub{b.p >= 0 ? b.v : 0., (double)b.nan, b.p};
            ll_publish(llC, par, pub, b.al, seq);
        }
        ll_gather(llC, par, G, s_part, seq);
        LexAcc loc{0., 0., -1, 0};
        if (G <= 32) {  // every warp reduces all G partials itself (cf. ll_reduce_w): no block barrier
            const int b = tid & 31;
            if (b < G) {
                const int p = (int)s_part[4 * b + 2];
                if (s_part[4 * b + 1] != 0.) loc.nan = 1;
                if (p >= 0) lex_push(loc, s_part[4 * b], p, s_part[4 * b + 3]);
            }
            loc = warp_lexmin(loc);
        } else {
            for (int b = tid; b < G; b += blockDim.x) {
                const int p = (int)s_part[4 * b + 2];
                if (s_part[4 * b + 1] != 0.) loc.nan = 1;
                if (p >= 0) lex_push(loc, s_part[4 * b], p, s_part[4 * b + 3]);
            }
            loc = block_lexmin(loc, &s_lex, lbuf);
        }
        if (R > 1) {  // every rank's (ratio, nan flag, position, pivot-row entry) to every rank
            if (blockIdx.x == 0 && tid < R * kMboxFields) {
                const int dst = tid / kMboxFields, f = tid % kMboxFields;
                const double v = f == 0 ? (loc.p >= 0 ? loc.v : 0.) : (f == 1 ? (double)loc.nan : (f == 2 ? (double)loc.p : loc.al));
                ll_send(mbox_slot(pl.mbox[dst], par, 0, me, f), v, seq);
            }
            if (tid < R * kMboxFields) s_mb[tid] = ll_recv(mbox_slot(pl.mbox[me], par, 0, tid / kMboxFields, tid % kMboxFields), seq);
            __threadfence();
            __syncthreads();
            loc = LexAcc{0., 0., -1, 0};
            for (int s = 0; s < R; ++s) {
                const int p = (int)s_mb[s * kMboxFields + 2];
                if (s_mb[s * kMboxFields + 1] != 0.) loc.nan = 1;
                if (p >= 0) lex_push(loc, s_mb[s * kMboxFields], p, s_mb[s * kMboxFields + 3]);
            }
            __syncthreads();  // s_mb is rewritten by the next pivot
        }
        if (loc.nan || loc.p < 0) {  // :279 panic / :281-284 dual unbounded => primal infeasible
            if (gtid == 0) {
                if (loc.nan) st->err = kErrNaNDualRatio;
                st->status = ELLP_INFEASIBLE;
                st->r_pos = r; st->delta = delta; st->do_update = 0; st->do_step = 0;
            }
            run = false;
            blk_zero_slot(lp, slot, gtid, gsize);
            continue;
        }
        const int q_pos = loc.p;
        const double theta_d = neg ? -loc.v : loc.v;  // :286-289
        const double alpha_rq = loc.al;               // pivot element alpha_q[r] (:306), see the header
        const double theta_p = delta / alpha_rq;
        const int ql = q_pos - lp.pos_lo;
        const bool owner = (ql >= 0 && ql < nT);
        const int cnt = slot;
        double* Uslot = lp.U + (int64_t)slot * lp.ld;
        double* Vslot = lp.V + (int64_t)slot * lp.ldv;
        // the owner needs V[0..cnt, ql] for the column: issue those loads before phase E
        if (owner && tid < cnt) s_vec[tid] = __ldcg(lp.V + (int64_t)tid * lp.ldv + ql);
        // ---- E: local part of the new V slot and of the reduced costs (dual :296-302)
        for (int64_t t = gtid; t < lp.ldv; t += gsize) {
            if (t < nT) {
                if ((int)t == ql) {
                    Vslot[t] = 1.0 / alpha_rq;
                    lp.dj[t] = -theta_d;  // the position now holds the leaving variable (:296)
                } else {
                    const double a_raw = (t == gtid) ? a_first : lp.rN[t];
                    Vslot[t] = a_raw / alpha_rq;
                    const double dold = __ldcg(lp.dj + t), prod = theta_d * a_raw;
                    double dnew = dold - prod;  // :298-300
                    // A difference that cancels to a few ulps of its operands is an exact tie of the ratio test (d_t / alpha_t ==
                    // theta_dual) seen through rounding: store the exact zero.  The pivot row of the tableau carries other rounding
                    // errors than the reference's A_N^T rho, so without this the next degenerate ratio test would rank its ties by
                    // +-1e-17 / alpha_t -- i.e. prefer the SMALLEST pivot element -- instead of by position (dual :270-279).
                    if (fabs(dnew) <= kDualSnapUlps * 2.220446049250313e-16 * fmax(fabs(dold), fabs(prod))) dnew = 0.;
                    lp.dj[t] = dnew;
                }
            } else {
                Vslot[t] = 0.;
            }
        }
        if (owner) { zq = ql; zcnt = cnt; }
        __syncthreads();  // s_vec
        // ---- C: entering column (owner: rebuild + broadcast), x step, bookkeeping, new U slot, bound check of the updated rows
        PivotRegs g_next = g;
        g_next.pivots = g.pivots + 1;
        g_next.trace_len = g.trace_len + 1;
        g_next.obj = g.obj + theta_d * delta;  // :316
        if (g_next.pivots >= g.max_iter) run = false;  // :191-194 at the next loop head
        const bool want_L = run && (slot + 1 < slot0 + npiv);
        Top2 tl{DEVEX ? -1.0 : CUDART_INF, DEVEX ? -1.0 : CUDART_INF, -1};
        double my_delta = 0.;
        const uint4* colbuf = pl.col[me] + (int64_t)par * pl.col_cap;
        for (int64_t i = gtid; i < lp.ld; i += gsize) {
            int var = 0, kv = ELLP_FIXED;
            double xv = 0., lbv = 0., ubv = 0., w_i = 1.;
            if (i < m) {  // independent of the column: in flight while the column entry is rebuilt / polled
                var = __ldcg(lp.Bv + i);
                xv = __ldcg(lp.x + var);
                kv = lp.kind[var]; lbv = lp.lb[var]; ubv = lp.ub[var];
                if (DEVEX) w_i = __ldcg(lp.w + i);
            }
            double a;
            if (owner) {
                a = corr_chain(__ldcg(lp.T + (int64_t)ql * lp.ld + i), lp.U + i, lp.ld, s_vec, cnt);
                for (int d = 0; d < R; ++d)
                    if (d != me) ll_send(pl.col[d] + (int64_t)par * pl.col_cap + i, a, seq);
                lp.T[(int64_t)ql * lp.ld + i] = (i == r) ? 1. : 0.;  // the stored column now belongs to the leaving variable: e_r
            } else {
                a = ll_recv(colbuf + i, seq);
            }
            if (i < m) {
                xv = xv - theta_p * a;  // :310-312
                lp.x[var] = xv;
                if (i == r) {  // this thread does the bookkeeping of the pivot (:314-333)
                    const int leave_var = var;
                    const int q_var = __ldcg(lp.Nv + q_pos);
                    xv = __ldcg(lp.x + q_var) + theta_p;  // :314
                    lp.x[q_var] = xv;
                    lp.Bv[r] = q_var;
                    lp.Nv[q_pos] = leave_var;
                    lp.Ns[q_pos] = (uint8_t)new_side;
                    lp.cB[r] = lp.c[q_var];
                    lp.d[q_var] = 0.;  // :302
                    if (lp.trace && g.trace_len < g.trace_cap) {
                        ellp_trace_rec rec;
                        rec.phase = g.phase_tag;
                        rec.iter = (int32_t)g.pivots;
                        rec.entering = q_var;
                        rec.leaving = leave_var;
                        rec.step = theta_p;
                        rec.obj = g.obj;
                        lp.trace[g.trace_len] = rec;
                    }
                    st->trace_len = g_next.trace_len;
                    st->obj = g_next.obj;
                    st->pivots = g_next.pivots;
                    st->r_pos = r; st->q_pos = q_pos; st->q_var = q_var; st->leave_var = leave_var; st->new_side = new_side;
                    st->delta = delta; st->theta_d = theta_d; st->alpha_r = alpha_rq; st->step = theta_p;
                    st->do_update = 1; st->do_step = 1;
                    if (g_next.pivots >= g.max_iter) st->status = ELLP_MAXITER;
                    kv = lp.kind[q_var]; lbv = lp.lb[q_var]; ubv = lp.ub[q_var];
                }
                if (DEVEX) {
                    const double ratio = a / alpha_rq;
                    w_i = (i == r) ? fmax(w_r / (alpha_rq * alpha_rq), 1.) : fmax(w_i, ratio * ratio * w_r);
                    lp.w[i] = w_i;
                }
                double dl;
                if (want_L && dual_violation(kv, lbv, ubv, xv, &dl)) {
                    if (DEVEX) {
                        const double score = dl * dl / w_i;
                        if (score > tl.a1) { tl.a1 = score; tl.i1 = (int)i; my_delta = dl; }
                    } else if (tl.i1 < 0) { tl.a1 = (double)i; tl.i1 = (int)i; my_delta = dl; }
                }
            }
            Uslot[i] = (i < m ? a : 0.) - (i == r ? 1. : 0.);
        }
        g = g_next;
        have_L = false;
        if (want_L) {  // phase L of the next pivot
            const int mine = tl.i1;
            tl = top2_block_fast<DEVEX>(tl, &s_top, rbuf);
            if (mine >= 0 && mine == tl.i1) s_delta = my_delta;
            __syncthreads();
            ll_publish(llA, par ^ 1, tl, (tl.i1 >= 0) ? s_delta : 0., seq + 1u);
            have_L = true;
        }
    }
    // the hand-over of the last pivot still has V entries to clear; every block must be past its column rebuild first
    grid.sync();
    if (zq >= 0 && gtid == (int64_t)zq % gsize)
        for (int j = 0; j < zcnt; ++j) lp.V[(int64_t)j * lp.ldv + zq] = 0.;
}

// dj[t] = d of the variable at local position t (start of a dual solve on the tableau: the caller's d is authoritative)
__global__ void k_dual_tab_init(DevLP lp) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < lp.nT) lp.dj[t] = lp.d[lp.Nv[lp.pos_lo + t]];
}
// d[var at position p] = dpos[p] for every nonbasic position (dpos = dj, or the all-gathered dj of the peer engine)
__global__ void k_dual_tab_scatter_d(DevLP lp, const double* __restrict__ dpos) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < lp.nN) lp.d[lp.Nv[p]] = dpos[p];
}
// rhs_i = c_j - d_j for the variable j that was basic in row i when the tableau was built (A_B0^T y = c_B0 - d_B0)
__global__ void k_dual_tab_yrhs(DevLP lp, const int32_t* __restrict__ Bv0, double* __restrict__ rhs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < lp.ld) rhs[i] = (i < lp.m) ? lp.c[Bv0[i]] - lp.d[Bv0[i]] : 0.;
}
// starting basis = diagonal matrix diag(s): y_i = rhs_i / s_i
__global__ void k_dual_tab_y_diag(DevLP lp, const double* __restrict__ rhs, const double* __restrict__ bscale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < lp.m) lp.y[i] = rhs[i] / bscale[i];
}

// V[j, t] = -A[k0 + j, Nv[t]] (row block of A_N as the row-major operand of the rank-k kernel), zero beyond nT
__global__ void k_gather_rows_neg(const double* __restrict__ A, int64_t ld, const int32_t* __restrict__ Nv, int nT, int k0, double* __restrict__ V,
                                  int64_t ldv) {
    const int j = blockIdx.y;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ldv; t += (int64_t)gridDim.x * blockDim.x)
        V[(int64_t)j * ldv + t] = (t < nT) ? -A[(int64_t)Nv[t] * ld + k0 + j] : 0.;
}
// T[i, j] /= d[i] for every stored column (diagonal starting basis), two rows per thread
__global__ void k_scale_rows_inv(double* __restrict__ T, int64_t ld, int m, int C, const double* __restrict__ d) {
    const int64_t i = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= m) return;
    const double d0 = d[i], d1 = (i + 1 < m) ? d[i + 1] : 1.0;
    for (int j = blockIdx.y; j < C; j += gridDim.y) {
        double2 v = ld_f64x2(T + (int64_t)j * ld + i);
        v.x = v.x / d0; v.y = v.y / d1;
        st_f64x2(T + (int64_t)j * ld + i, v);
    }
}
// ---- x_B = B^-1 b - T x_N at a tableau rebuild (the reference never recomputes x; a rebuild is where the drift of the
// incrementally updated point is removed as well).  Deterministic: column chunks, partial sums added in chunk order.
constexpr int kXChunks = 32;
// part[c * ld + i] = sum over the columns j of chunk c of M[i, j] * v(j);  v(j) = vec[idx ? idx[j] : j]
__global__ void __launch_bounds__(256) k_gemv_n_chunks(const double* __restrict__ M, int64_t ld, int R, int C, const double* __restrict__ vec,
                                                       const int32_t* __restrict__ idx, double* __restrict__ part) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y, per = (C + kXChunks - 1) / kXChunks;
    const int j0 = c * per, j1 = min(C, j0 + per);
    if (i >= R) return;
    double acc = 0.;
    for (int j = j0; j < j1; ++j) acc = fma(M[(int64_t)j * ld + i], vec[idx ? idx[j] : j], acc);
    part[(int64_t)c * ld + i] = acc;
}
// x[Bv[i]] = sum_c pb[c][i] * (inv_diag ? 1 / bscale[i] ... ) - sum_c pt[c][i]
__global__ void k_xB_finish(DevLP lp, const double* __restrict__ pb, const double* __restrict__ pt, int diag_only) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lp.m) return;
    double xb = 0., xt = 0.;
    if (diag_only) xb = lp.b[i] / lp.bscale[i];
    else for (int c = 0; c < kXChunks; ++c) xb += pb[(int64_t)c * lp.ld + i];
    for (int c = 0; c < kXChunks; ++c) xt += pt[(int64_t)c * lp.ld + i];
    lp.x[lp.Bv[i]] = xb - xt;
}
// residual check of a long solve: out[0] = max_i |sum_c part[c][i] - b_i| (ordered-bit atomicMax of a non-negative double)
__global__ void __launch_bounds__(256) k_residual_max(const double* __restrict__ part, int64_t ld, int m, const double* __restrict__ b,
                                                      unsigned long long* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double r = 0.;
    if (i < m) {
        double ax = 0.;
        for (int c = 0; c < kXChunks; ++c) ax += part[(int64_t)c * ld + i];
        r = fabs(ax - b[i]);
        if (r != r) r = CUDART_INF;  // a NaN residual must trigger the rebuild
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, off));
    if ((threadIdx.x & 31) == 0 && r > 0.) atomicMax(out, (unsigned long long)__double_as_longlong(r));
}
__global__ void k_fill_const(double* __restrict__ p, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace ellp
