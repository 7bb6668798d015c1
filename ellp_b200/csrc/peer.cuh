// peer.cuh -- column-sharded blocked tableau engine with the per-pivot exchange fused INTO the pivot kernel.
//
// BASELINE.json configs[4] / north_star: "a single very large dense tableau is column-sharded, with each GPU pricing its
// own columns, a per-iteration argmin allreduce over NVLink, and the pivot column broadcast".  The first implementation
// (launch_sharded_iteration, engine.cu) calls NCCL three times per pivot between seven small kernels; here the
// arg-reduce and the column broadcast are stores into PEER MEMORY (cudaIpc-mapped buffers of the other ranks, NVLink 5 /
// NVSwitch) issued by the same persistent cooperative kernel that prices, runs the ratio test and appends the (U, V)
// slot, so a pivot costs two NVLink one-way latencies instead of three collective launches.
//
// Layout (rank g of G): the CONDENSED tableau of blocked.cuh, split by nonbasic POSITION: T, dj, V, key, rN hold the
// positions [pos_lo, pos_lo + nT) of the N list, nT = nN / G.  Everything index-level or O(m + n) is replicated and
// evolves identically on every rank: x, Bv, Nv, Ns, U, PivotState, the ratio test and its tie fold.
//
// Wire protocol: every double travels as one 16-byte word {lo32, seq, hi32, seq} written with a single vector store
// and polled by the consumer until both sequence fields match (the "LL" idea of NCCL's low-latency protocol): data and
// flag arrive together, so there is no fence, no separate flag and no ordering requirement between stores on the
// link.  seq = number of the pivot since the communicator was created (never repeats); buffers are double-buffered by
// the parity of seq: a rank can be at most one pivot ahead of the slowest reader of its previous message, because it
// cannot finish pivot p+1 without every rank's pricing message for p+1, which is sent after that rank finished pivot p.
// A consumer that waits longer than kPeerTimeoutNs traps (a dead peer must not hang the GPU).
//
// Per pivot and rank (slot = index of the pending (U, V) pair, cf. blocked.cuh), k_blk_pivots_fused:
//   A   (first pivot of a launch only; afterwards E already did it) Dantzig keys of the local positions, per-block
//       (best, second best, position, reduced cost), published as four LL words per block in a local area (ll_publish)
//   B   every block polls all blocks' words and reduces them (ll_gather / ll_reduce): an all-to-all that doubles as the
//       grid barrier and leaves the rank-local result in every thread.  One block then stores it into every rank's
//       mailbox; every block polls the G mailbox entries and merges them in rank order => the same entering position on
//       every rank.  Near-tie (best - second < 2 EPS, any two ranks): second round with the order-free rule of SURVEY
//       appendix A.1 (largest variable index within EPS of the global maximum), one mailbox entry per rank (here the
//       rank-local reduction uses the last-block ticket: the path is rare).
//   C1  the owner of the entering position rebuilds that column of the current tableau (stale column + pending
//       corrections, same arithmetic as k_ratio_prep) and stores it into every rank's column buffer
//   C2  every rank polls the column (row i by the thread that needs row i), ratios, per-block two smallest + the column
//       entry of the best row, published like in A; gathering them is the second (and last) barrier of the pivot
//   D   every thread merges the partials and derives the decision itself (leaving row, lambda, new side, still running);
//       exact tie fold (ratio_pick_body) only when the minimum is not isolated
//   E   x step, bookkeeping by the one thread that owns row r, local part of the pivot row and of the reduced-cost row,
//       new (U, V) slot, and the Dantzig keys of the NEXT pivot (phase A of pivot p+1).  No barrier: the next pivot's
//       all-to-all separates it from the next reader.
// Reference lines: pricing primal_simplex_solver.rs:189,253-292; column + ratios :295-367; fold/step/apply :379-434,
// :205-232.  Decisions equal those of the NCCL path and of the oracle's canonical mode pivot for pivot (tests).
#pragma once
#include <cooperative_groups.h>
#include "blocked.cuh"

namespace ellp {

constexpr int kMaxPeers = 8;
constexpr long long kPeerTimeoutNs = 8000000000ll;  // 8 s
constexpr int kMboxFields = 4;
constexpr int kMboxRounds = 3;  // 0: pricing, 1: near-tie round, 2: the owner's ratio-test decision
constexpr int kMboxWords = 2 /*parity*/ * kMboxRounds * kMaxPeers * kMboxFields;  // uint4 words per rank

struct PeerLinks {
    int32_t rank, nranks;
    int64_t col_cap;         // rows per parity slot of a column buffer
    uint4* mbox[kMaxPeers];  // mbox[r] = rank r's mailbox (peer-mapped unless r == rank)
    uint4* col[kMaxPeers];   // col[r]  = rank r's column buffer, 2 * col_cap words
    unsigned int* ticket;    // local: arrival counters of the last-block pattern (2)
    long long* tlog;         // optional phase-timing log (tuning key "phase_timing"): kTlogStamps clock64() values per pivot, block 0 thread 0
    int32_t tlog_cap;        // pivots the log can hold
    int32_t owner_only;      // primal kernel, R > 1: 1 = only the owner of the entering position runs the ratio test and broadcasts
                             // the decision; 0 = every rank polls the column and decides redundantly (round-1 protocol)
};
constexpr int kTlogStamps = 10;

__device__ __forceinline__ long long peer_now_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void ll_send(uint4* dst, double v, uint32_t seq) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"((uint32_t)b), "r"(seq), "r"((uint32_t)(b >> 32)), "r"(seq)
                 : "memory");
}

__device__ __forceinline__ double ll_recv(const uint4* src, uint32_t seq) {
    uint32_t lo, f0, hi, f1;
    long long t0 = 0;
    unsigned spins = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(src) : "memory");
        if (f0 == seq && f1 == seq) break;
        if ((++spins & 4095u) == 0u) {
            const long long now = peer_now_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kPeerTimeoutNs) __trap();  // a peer died or diverged: fail loudly instead of hanging the GPU
        }
    }
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

__device__ __forceinline__ uint4* mbox_slot(uint4* base, int par, int round, int src, int field) {
    return base + (((par * kMboxRounds + round) * kMaxPeers + src) * kMboxFields + field);
}

// Last-block pattern: returns true in every thread of exactly one block per call site and pivot -- the block whose
// arrival completed the grid.  All writes the other blocks made before arriving are visible to it.
__device__ __forceinline__ bool peer_arrive_last(unsigned int* ticket, int* s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        const int last = (t == gridDim.x - 1);
        if (last) { *ticket = 0u; __threadfence(); }
        *s_flag = last;
    }
    __syncthreads();
    return *s_flag != 0;
}

// ------------------------------------------------------------------------------------------------
// Version 2 of the pivot kernel: TWO barriers per pivot instead of four.
//   * decisions are replicated: after the merge of phase D every thread knows (leaving row, lambda) and derives
//     do_step / do_update / the new side / "still running" itself, so nobody waits for one thread to publish them;
//     the bookkeeping of primal :205-232 (ratio_commit: Bv / Nv / Ns swap, trace, objective, pivot count, status) is
//     carried out by ONE thread -- the one that owns row r in the x step, after it read Bv[r] -- concurrently with E;
//   * E is fused with the NEXT pivot's phase A: the thread that writes the new reduced cost d_t prices it at once, so
//     the per-block (best, second best) of pivot p+1 leave E of pivot p and the only barrier between two pivots is the
//     pricing mailbox itself.
// Per pivot: [ticket -> mailbox send] B (mailbox wait) C1 C2 | grid barrier | D E+A'.  Exact tie folds (ratio_pick_body,
// and select_primal_body for the reference rule on a single rank) keep their extra barriers; they are rare.
// Serves the single-GPU blocked engine too (nranks == 1: the mailbox and the column buffer are this GPU's own memory).
// ------------------------------------------------------------------------------------------------
// ---- block-wide (best, second best, index of the best) on the integer pipe -------------------------------------------
// Doubles are mapped to order-preserving 64-bit keys; a warp maximum is two redux.sync (high word, then low word among
// the lanes that hold the winning high word).  Second best = the same reduction with the winner lane contributing its own
// second.  ~6 warp-collective instructions instead of 5 shuffle steps over a (double, double, int) triple; one
// __syncthreads per block reduction thanks to two alternating shared-memory buffers.
__device__ __forceinline__ unsigned long long ord_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_val(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
template <bool MAX> __device__ __forceinline__ unsigned long long warp_best_u64(unsigned long long v) {
    const unsigned full = 0xffffffffu;
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    if (MAX) {
        const unsigned mh = __reduce_max_sync(full, hi);
        const unsigned ml = __reduce_max_sync(full, hi == mh ? lo : 0u);
        return ((unsigned long long)mh << 32) | ml;
    } else {
        const unsigned mh = __reduce_min_sync(full, hi);
        const unsigned ml = __reduce_min_sync(full, hi == mh ? lo : 0xffffffffu);
        return ((unsigned long long)mh << 32) | ml;
    }
}
template <bool MAX> __device__ __forceinline__ Top2 warp_top2(const Top2& t) {
    const unsigned full = 0xffffffffu;
    const unsigned long long k1 = ord_key(t.a1), k2 = ord_key(t.a2);
    const unsigned long long M = warp_best_u64<MAX>(k1);
    const int src = __ffs(__ballot_sync(full, k1 == M)) - 1;
    const unsigned long long S = warp_best_u64<MAX>(((int)(threadIdx.x & 31) == src) ? k2 : k1);
    Top2 r;
    r.a1 = ord_val(M);
    r.a2 = ord_val(S);
    r.i1 = __shfl_sync(full, t.i1, src);
    return r;
}
struct Top2Fast {
    double a1[2][32], a2[2][32];
    int i1[2][32];
};
// result valid in every thread; `buf` alternates between calls (block-uniform)
template <bool MAX> __device__ __forceinline__ Top2 top2_block_fast(const Top2& t, Top2Fast* sh, int& buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const Top2 w = warp_top2<MAX>(t);
    if (lane == 0) { sh->a1[buf][warp] = w.a1; sh->a2[buf][warp] = w.a2; sh->i1[buf][warp] = w.i1; }
    __syncthreads();
    const double worst = MAX ? -1.0 : CUDART_INF;
    Top2 r;
    r.a1 = (lane < nw) ? sh->a1[buf][lane] : worst;
    r.a2 = (lane < nw) ? sh->a2[buf][lane] : worst;
    r.i1 = (lane < nw) ? sh->i1[buf][lane] : -1;
    buf ^= 1;
    return warp_top2<MAX>(r);
}
// Uniform (replicated in every thread) scalar state of the solve, so that the one thread that does the bookkeeping needs
// no dependent loads from PivotState: refreshed from st at launch and after every tie-path commit.
struct PivotRegs {
    uint64_t pivots, max_iter;
    int64_t trace_len, trace_cap;
    double obj;
    int32_t phase_tag;
};
__device__ __forceinline__ void pivot_regs_load(PivotRegs& g, const PivotState* st) {
    g.pivots = __ldcg(&st->pivots);
    g.max_iter = __ldcg(&st->max_iter);
    g.trace_len = __ldcg(&st->trace_len);
    g.trace_cap = __ldcg(&st->trace_cap);
    g.obj = __ldcg(&st->obj);
    g.phase_tag = __ldcg(&st->phase_tag);
}

// ratio_commit (kernels.cuh) with every scalar input in registers: same stores, two independent loads.  `g` holds the
// state BEFORE this pivot.  One thread.
__device__ __forceinline__ void ratio_commit_regs(const DevLP& lp, PivotState* st, const PivotRegs& g, int nb, double lambda, bool at_lower,
                                                  int q_pos, int q_var, int q_side, double rq, double alpha_r, int side_after) {
    if (!(lambda >= 0.)) {  // :402 assert!(lambda >= 0.)
        st->err = kErrLambdaNegative; st->status = ELLP_UNBOUNDED; st->do_update = 0; st->do_step = 0;
        return;
    }
    if (isinf(lambda)) {  // :404-406
        st->status = ELLP_UNBOUNDED; st->do_update = 0; st->do_step = 0;
        return;
    }
    st->do_step = (lambda > 0.) ? 1 : 0;
    int leave_var = -1;
    int status = kRunning;
    if (nb >= 0) {  // :208-221
        leave_var = __ldcg(lp.Bv + nb);
        const double cq = lp.c[q_var];
        lp.Bv[nb] = q_var;
        lp.Nv[q_pos] = leave_var;
        lp.Ns[q_pos] = (uint8_t)side_after;
        lp.cB[nb] = cq;
        st->r_pos = nb;
        st->leave_var = leave_var;
        st->alpha_r = alpha_r;
        st->do_update = 1;
    } else {  // :223-231 bound flip
        if (q_side == ELLP_NB_FREE) { st->err = kErrFlipFree; st->status = ELLP_UNBOUNDED; status = ELLP_UNBOUNDED; }
        else lp.Ns[q_pos] = (uint8_t)side_after;
        st->r_pos = -1;
        st->do_update = 0;
    }
    if (lp.trace && g.trace_len < g.trace_cap) {
        ellp_trace_rec rec;
        rec.phase = g.phase_tag;
        rec.iter = (int32_t)g.pivots;
        rec.entering = q_var;
        rec.leaving = leave_var;
        rec.step = lambda;
        rec.obj = g.obj;
        lp.trace[g.trace_len] = rec;
    }
    st->trace_len = g.trace_len + 1;
    st->step = lambda;
    st->obj = g.obj + rq * (at_lower ? lambda : -lambda);
    st->pivots = g.pivots + 1;
    if (status == kRunning && g.pivots + 1 >= g.max_iter) st->status = ELLP_MAXITER;  // :163-166 at the next loop head
}

struct PivotDec {
    int do_step, do_update, r, leave_var;  // leave_var >= 0: the bookkeeping already happened (tie path), use it for row r
    int q_pos, q_var, q_side, side_after;  // side_after: side of nonbasic position q_pos after this pivot (ELLP_NB_*)
    bool at_lower;
    double lambda, alpha_r, rq;
    double wq;  // Devex reference weight of the entering position (ELLP_PRICE_DEVEX)
};

__device__ __forceinline__ double dantzig_key(double r, int side) {
    double k = -1.0;
    if (!(fabs(r) < kEps)) {
        if (r > 0. && side == ELLP_NB_UPPER) k = r;
        else if (!(r > 0.) && side == ELLP_NB_LOWER) k = -r;
        else if (side == ELLP_NB_FREE) k = fabs(r);
    }
    return k;
}

// ELLP_PRICE_DEVEX (no reference counterpart: ellp prices with Dantzig's rule): key = |r| / sqrt(w) over the same candidates (the
// ordering of r^2 / w at the magnitude of a reduced cost; the EPS-tolerant tie rules of the reference are NOT applied to these keys --
// keys far below EPS would all "tie" and the index rule would replace the pricing; exact ties keep the first candidate), with
// primal Devex reference weights w per nonbasic position: after a pivot on (r, q) with scaled pivot row p_t = alpha_rt / alpha_rq,
// w_t = max(w_t, p_t^2 w_q) and the position handed to the leaving variable gets max(w_q / alpha_rq^2, 1).  p_t is what the row
// phase computes anyway, so the rule adds one load and one store per position and no exchange.
__device__ __forceinline__ double devex_key(double r, int side, double w) {
    const double k = dantzig_key(r, side);
    return (k == -1.0) ? -1.0 : fabs(r) / sqrt(w);
}

// acc - sum_j g[j * stride] * coef[j], j ascending (one fma per pending pivot, i.e. the roundings of j rank-1 updates), with the
// loads issued 16 at a time: the chain of dependent fmas is short, the exposed latency was one L2 round trip per 4 terms.
__device__ __forceinline__ double corr_chain(double acc, const double* g, int64_t stride, const double* coef, int cnt) {
    int j = 0;
    for (; j + 16 <= cnt; j += 16) {
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = __ldcg(g + (int64_t)(j + q) * stride);
#pragma unroll
        for (int q = 0; q < 16; ++q) acc = fma(-v[q], coef[j + q], acc);
    }
    for (; j + 4 <= cnt; j += 4) {
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = __ldcg(g + (int64_t)(j + q) * stride);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc = fma(-v[q], coef[j + q], acc);
    }
    for (; j < cnt; ++j) acc = fma(-__ldcg(g + (int64_t)j * stride), coef[j], acc);
    return acc;
}

// What phase C2 already loaded for the first row a thread owns (row == its global thread index): phase E reuses it
// instead of walking Bv -> x again.
struct RowCache {
    int64_t row;  // -1: nothing cached
    double a;     // entering-column entry (valid for row < ld)
    double xv;    // x[var] before the step (valid for row < m)
    int var;      // Bv[row] before the bookkeeping
};

// x step, (optionally) the bookkeeping, local pivot row / reduced costs / new (U, V) slot, and -- when `price` -- the
// Dantzig keys of the local positions for the next pivot (returned as this thread's Top2).
template <bool DEVEX>
__device__ __forceinline__ Top2 blk_row_price_body(const DevLP& lp, int slot, PivotState* st, const PivotDec& d, const PivotRegs& g, bool commit_here,
                                                   bool price, int64_t t0, int64_t stride, double* su, const RowCache& rc, const uint4* colsrc,
                                                   uint32_t seq) {
    const int n = lp.nT, m = lp.m;
    // entering-column entry of row t: phase C2 left it in dcol on a deciding rank; the other ranks poll the owner's LL word
    auto col_at = [&](int64_t t) { return colsrc ? ll_recv(colsrc + t, seq) : __ldcg(lp.dcol + t); };
    const int64_t tmax = max(lp.ld, lp.ldv);
    const int r = d.r;
    // the corrections of the pivot row need U[r, 0..slot): issue those loads before anything else of this phase
    if (d.do_update && threadIdx.x < slot) su[threadIdx.x] = __ldcg(lp.U + (int64_t)threadIdx.x * lp.ld + r);
    if (d.do_step) {  // primal :408-417
        for (int64_t t = t0; t < m; t += stride) {
            if (t == rc.row) {  // Bv[t] (before the bookkeeping), x and the column entry are already in registers
                const double d_i = d.at_lower ? -rc.a : rc.a;
                lp.x[rc.var] = rc.xv + d.lambda * d_i;
                continue;
            }
            const double a = col_at(t);
            const double d_i = d.at_lower ? -a : a;
            const int var = (d.leave_var >= 0 && t == r) ? d.leave_var : __ldcg(lp.Bv + t);
            lp.x[var] = __ldcg(lp.x + var) + d.lambda * d_i;
        }
        if (t0 == 0) lp.x[d.q_var] = d.at_lower ? __ldcg(lp.x + d.q_var) + d.lambda : __ldcg(lp.x + d.q_var) - d.lambda;
    }
    if (commit_here)  // this thread owns row r (or is thread 0 for a flip): it read Bv[r] above, now it may overwrite it
        ratio_commit_regs(lp, st, g, d.do_update ? r : -1, d.lambda, d.at_lower, d.q_pos, d.q_var, d.q_side, d.rq, d.alpha_r, d.side_after);
    double* Uslot = lp.U + (int64_t)slot * lp.ld;
    double* Vslot = lp.V + (int64_t)slot * lp.ldv;
    Top2 best{-1.0, -1.0, -1};
    const int qp_any = d.q_pos - lp.pos_lo;  // local index of the entering position (may be out of range on this rank)
    if (!d.do_update) {
        for (int64_t t = t0; t < tmax; t += stride) {
            if (t < lp.ld) Uslot[t] = 0.;
            if (t < lp.ldv) Vslot[t] = 0.;
            if (price && t < n) {
                const double rc = __ldcg(lp.dj + t);
                const int side = (t == qp_any) ? d.side_after : (int)__ldcg(lp.Ns + lp.pos_lo + t);
                const double k = DEVEX ? devex_key(rc, side, __ldcg(lp.wN + t)) : dantzig_key(rc, side);
                lp.key[t] = k;
                lp.rN[t] = rc;
                if (k != -1.0) top2_push<true>(best, k, (int)t);
            }
        }
        return best;
    }
    const int qp = (lp.condensed && qp_any >= 0 && qp_any < n) ? qp_any : -1;
    __syncthreads();
    for (int64_t t = t0; t < tmax; t += stride) {
        if (t < n) {
            double dnew, wnew = 1.;
            const double dold = __ldcg(lp.dj + t);
            const double wold = DEVEX ? __ldcg(lp.wN + t) : 1.;
            const int side_t = price ? (int)__ldcg(lp.Ns + lp.pos_lo + t) : 0;
            if (t == qp) {  // handed over to the leaving variable, whose current column is e_r
                const double p = 1.0 / d.alpha_r;
                Vslot[t] = p;
                dnew = fma(-d.rq, p, 0.);
                if (DEVEX) wnew = fmax(d.wq / (d.alpha_r * d.alpha_r), 1.);
            } else {
                const double e = corr_chain(__ldcg(lp.T + t * lp.ld + r), lp.V + t, lp.ldv, su, slot);
                const double p = e / d.alpha_r;
                Vslot[t] = p;
                dnew = fma(-d.rq, p, dold);
                if (DEVEX) wnew = fmax(wold, p * p * d.wq);
            }
            lp.dj[t] = dnew;
            if (DEVEX) lp.wN[t] = wnew;
            if (price) {
                const int side = (t == qp) ? d.side_after : side_t;
                const double k = DEVEX ? devex_key(dnew, side, wnew) : dantzig_key(dnew, side);
                lp.key[t] = k;
                lp.rN[t] = dnew;
                if (k != -1.0) top2_push<true>(best, k, (int)t);
            }
        } else if (t < lp.ldv) {
            Vslot[t] = 0.;
        }
        if (t < lp.ld) {
            const double a = (t == rc.row) ? rc.a : (t < m ? col_at(t) : 0.);
            Uslot[t] = (t < m ? a : 0.) - (t == r ? 1. : 0.);
            if (qp >= 0) lp.T[(int64_t)qp * lp.ld + t] = (t == r) ? 1. : 0.;
        }
        if (qp >= 0 && t < slot) lp.V[t * lp.ldv + qp] = 0.;
    }
    return best;
}

constexpr int kFusedMaxThreads = 512;
constexpr int kLLMaxBlocks = 160;  // blocks of k_blk_pivots_fused (local all-to-all areas in DevLP::coop: 2 x 2 x 160 x 4 words)

// Intra-GPU barrier WITH payload: every block stores its (best, second, index, extra) as four LL words into its slot of a
// local area; every block then polls all G slots (ll_gather) and reduces them itself.  Compared with "grid.sync + read the
// partials" this is one store -> poll hop instead of an atomic ticket, a flag spin and a dependent load; like grid.sync it
// orders all earlier global writes of every block before every later read (fence before the publish, fence after the poll).
__device__ __forceinline__ void ll_publish(uint4* area, int par, const Top2& t, double extra, uint32_t seq) {
    if (threadIdx.x < 4) {
        const int f = threadIdx.x;
        const double v = f == 0 ? t.a1 : (f == 1 ? t.a2 : (f == 2 ? (double)t.i1 : extra));
        __threadfence();
        ll_send(area + ((size_t)par * kLLMaxBlocks + blockIdx.x) * 4 + f, v, seq);
    }
}
// all threads; s_part[4 b + f] = field f of block b.  Ends with a block barrier.
__device__ __forceinline__ void ll_gather(const uint4* area, int par, int nblocks, double* s_part, uint32_t seq) {
    const uint4* base = area + (size_t)par * kLLMaxBlocks * 4;
    for (int w = threadIdx.x; w < 4 * nblocks; w += blockDim.x) s_part[w] = ll_recv(base + w, seq);
    __threadfence();
    __syncthreads();
}
// (best, second, index) over the gathered slots + the `extra` word of the winning slot; result valid in every thread
template <bool MAX> __device__ __forceinline__ Top2 ll_reduce(const double* s_part, int nblocks, Top2Fast* sh, int& buf, double* s_extra, double* extra) {
    const double worst = MAX ? -1.0 : CUDART_INF;
    Top2 t{worst, worst, -1};
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
        const Top2 o{s_part[4 * b], s_part[4 * b + 1], (int)s_part[4 * b + 2]};
        top2_merge<MAX>(t, o);
    }
    t = top2_block_fast<MAX>(t, sh, buf);
    if (threadIdx.x == 0) *s_extra = 0.;
    __syncthreads();
    if (t.i1 >= 0)
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
            if ((int)s_part[4 * b + 2] == t.i1 && s_part[4 * b] == t.a1) *s_extra = s_part[4 * b + 3];  // indices are unique across blocks
    __syncthreads();
    *extra = *s_extra;
    return t;
}
// The same reduction done by EVERY WARP on its own (lane l takes a contiguous chunk of the slots, then one warp-wide top-2): no
// block barrier at all -- s_part is complete after the __syncthreads that ends ll_gather, and a few redundant shared-memory
// reads per lane are cheaper than the three barriers of ll_reduce on this latency chain.  Ties keep the lowest slot.
template <bool MAX> __device__ __forceinline__ Top2 ll_reduce_w(const double* s_part, int nblocks, double* extra) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const double worst = MAX ? -1.0 : CUDART_INF;
    const int per = (nblocks + 31) >> 5;
    Top2 t{worst, worst, -1};
    int slot = -1;
    for (int b = lane * per; b < min(nblocks, (lane + 1) * per); ++b) {
        const Top2 o{s_part[4 * b], s_part[4 * b + 1], (int)s_part[4 * b + 2]};
        if (top2_better<MAX>(o.a1, t.a1)) slot = b;
        top2_merge<MAX>(t, o);
    }
    const unsigned long long k1 = ord_key(t.a1), k2 = ord_key(t.a2);
    const unsigned long long M = warp_best_u64<MAX>(k1);
    const int src = __ffs(__ballot_sync(full, k1 == M)) - 1;
    const unsigned long long S = warp_best_u64<MAX>((lane == src) ? k2 : k1);
    Top2 r;
    r.a1 = ord_val(M);
    r.a2 = ord_val(S);
    r.i1 = __shfl_sync(full, t.i1, src);
    const int wslot = __shfl_sync(full, slot, src);
    *extra = (r.i1 >= 0 && wslot >= 0) ? s_part[4 * wslot + 3] : 0.;
    return r;
}
template <bool DEVEX>
__global__ void __launch_bounds__(kFusedMaxThreads, 1) k_blk_pivots_fused(DevLP lp, PeerLinks pl, int tie_rule, int slot0, int npiv, uint32_t seq0,
                                                                      PivotState* st) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char scan_smem[];
    __shared__ Top2Fast s_top;
    __shared__ double s_vec[kBlkMax];
    __shared__ double s_mb[kMaxPeers * kMboxFields];
    __shared__ long long s_ll[32];
    __shared__ int s_flag;
    __shared__ double s_part[kLLMaxBlocks * 4];
    __shared__ double s_extra;
    const int tid = threadIdx.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + tid, gsize = (int64_t)gridDim.x * blockDim.x;
    const int G = gridDim.x, R = pl.nranks, me = pl.rank;
    uint4* llA = reinterpret_cast<uint4*>(lp.coop);          // pricing partials of every block, [parity][block][4 words]
    uint4* llC = reinterpret_cast<uint4*>(lp.coop + 2560);   // ratio partials
    long long* partT = reinterpret_cast<long long*>(lp.coop + 5120);  // near-tie round: per-block (variable << 32 | local position)
    const int nT = lp.nT, m = lp.m;
    bool run = (__ldcg(&st->status) == kRunning);
    PivotRegs g;
    pivot_regs_load(g, st);
    int rbuf = 0;  // alternating buffer of the block reductions
    bool priced = false;  // partA of the coming pivot already written by the previous pivot's E
    for (int slot = slot0; slot < slot0 + npiv; ++slot) {
        if (!run) { blk_zero_slot(lp, slot, gtid, gsize); continue; }
        const uint32_t seq = seq0 + (uint32_t)(slot - slot0) + 1u;
        const int par = (int)(seq & 1u);
        long long* tl = (pl.tlog && gtid == 0 && (int)(seq - 1u) < pl.tlog_cap) ? pl.tlog + (int64_t)(seq - 1u) * kTlogStamps : nullptr;
        if (tl) tl[0] = clock64();
        // ---- A (first pivot of the launch only): pricing of the local positions (primal :189, :253-270)
        if (!priced) {
            Top2 t{-1.0, -1.0, -1};
            for (int64_t j = gtid; j < nT; j += gsize) {
                const double r = __ldcg(lp.dj + j);
                const double k = DEVEX ? devex_key(r, __ldcg(lp.Ns + lp.pos_lo + j), __ldcg(lp.wN + j)) : dantzig_key(r, __ldcg(lp.Ns + lp.pos_lo + j));
                lp.key[j] = k;
                lp.rN[j] = r;
                if (k != -1.0) top2_push<true>(t, k, (int)j);
            }
            t = top2_block_fast<true>(t, &s_top, rbuf);
            ll_publish(llA, par, t, (t.i1 >= 0) ? __ldcg(lp.dj + t.i1) : 0., seq);
        }
        // every block gathers every block's pricing partial: the local (best, second, position, reduced cost) without a grid barrier
        double loc_rq;
        ll_gather(llA, par, G, s_part, seq);
        // few blocks (<= one slot per lane): every warp reduces on its own, no block barrier; many blocks: block-wide reduction
        // (measured at 148 blocks: the per-warp variant costs ~1 us more per reduction, at <= 32 blocks it saves ~0.2 us)
        const Top2 loc = (G <= 32) ? ll_reduce_w<true>(s_part, G, &loc_rq) : ll_reduce<true>(s_part, G, &s_top, rbuf, &s_extra, &loc_rq);
        if (R > 1 && blockIdx.x == 0 && tid < R * kMboxFields) {  // one block per rank tells the other ranks
            const int dst = tid / kMboxFields, f = tid % kMboxFields;
            double v;
            if (f == 0) v = loc.a1;
            else if (f == 1) v = loc.a2;
            else if (f == 2) v = (loc.i1 >= 0) ? (double)(lp.pos_lo + loc.i1) : -1.0;
            else v = loc_rq;
            ll_send(mbox_slot(pl.mbox[dst], par, 0, me, f), v, seq);
        }
        if (tl) tl[1] = clock64();
        // ---- B: entering position, identical on every rank (primal :271-292)
        int q_pos;
        double rq, wq = 1.;
        {
            Top2 t{-1.0, -1.0, -1};
            rq = 0.;
            if (R == 1) {  // single GPU: the local result is the result
                t = loc;
                rq = loc_rq;
            } else {
                if (tid < R * kMboxFields) s_mb[tid] = ll_recv(mbox_slot(pl.mbox[me], par, 0, tid / kMboxFields, tid % kMboxFields), seq);
                __threadfence();
                __syncthreads();
                for (int s = 0; s < R; ++s) {
                    const Top2 o{s_mb[s * kMboxFields], s_mb[s * kMboxFields + 1], (int)s_mb[s * kMboxFields + 2]};
                    if (o.a1 > t.a1) rq = s_mb[s * kMboxFields + 3];
                    top2_merge<true>(t, o);
                }
            }
            if (t.a1 == -1.0) {  // no candidate on any rank: optimal (:289-292)
                if (gtid == 0) { st->status = ELLP_OPTIMAL; st->do_update = 0; st->do_step = 0; }
                run = false;
                blk_zero_slot(lp, slot, gtid, gsize);
                continue;
            }
            q_pos = t.i1;
            const bool near_tie = !DEVEX && (t.a1 - t.a2 < 2. * kEps);
            if (near_tie && R == 1 && tie_rule == ELLP_TIES_REFERENCE) {
                // the reference's sequential max_by fold over N (primal :271-286), one block, then a grid barrier
                __syncthreads();
                if (blockIdx.x == 0) {
                    if (tid == 0) { st->do_update = 0; st->do_step = 0; }
                    select_primal_body(lp.key, lp.rN, lp.Nv, lp.Ns, lp.nN, tie_rule, st, scan_smem);
                }
                grid.sync();
                q_pos = __ldcg(&st->q_pos);
                rq = __ldcg(&st->rq);
                if (DEVEX) wq = __ldcg(lp.wN + q_pos);  // R == 1: every position is local
            } else {
                if (DEVEX) { const double sq = rq / t.a1; wq = sq * sq; }  // key = |r| / sqrt(w) of the winning position; the same value on every rank
                if (near_tie) {
                    // largest variable index among the keys within EPS of the global maximum
                    const double kmax = t.a1;
                    long long best = -1;
                    for (int64_t j = gtid; j < nT; j += gsize) {
                        const double k = lp.key[j];  // written by this thread (phase A / fused pricing)
                        if (k != -1.0 && (kmax - k < kEps)) {
                            const long long cand = ((long long)__ldcg(lp.Nv + lp.pos_lo + j) << 32) | (long long)j;
                            best = cand > best ? cand : best;
                        }
                    }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const long long o = __shfl_xor_sync(0xffffffffu, best, off);
                        best = o > best ? o : best;
                    }
                    __syncthreads();
                    if ((tid & 31) == 0) s_ll[tid >> 5] = best;
                    __syncthreads();
                    if (tid == 0) {
                        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = s_ll[w] > best ? s_ll[w] : best;
                        partT[blockIdx.x] = best;
                    }
                    if (peer_arrive_last(pl.ticket + 1, &s_flag)) {
                        if (tid == 0) {
                            long long b = -1;
                            for (int k = 0; k < G; ++k) { const long long o = __ldcg(partT + k); b = o > b ? o : b; }
                            s_ll[0] = b;
                        }
                        __syncthreads();
                        const long long b = s_ll[0];
                        if (tid < R * kMboxFields) {
                            const int dst = tid / kMboxFields, f = tid % kMboxFields;
                            const int jl = (int)(b & 0xffffffffll);
                            double v;
                            if (f == 0) v = (b >= 0) ? (double)(b >> 32) : -1.0;
                            else if (f == 1) v = (b >= 0) ? (double)(lp.pos_lo + jl) : -1.0;
                            else if (f == 2) v = (b >= 0) ? __ldcg(lp.dj + jl) : 0.;
                            else v = (DEVEX && b >= 0) ? __ldcg(lp.wN + jl) : 1.;
                            ll_send(mbox_slot(pl.mbox[dst], par, 1, me, f), v, seq);
                        }
                    }
                    __syncthreads();
                    if (tid < R * kMboxFields) s_mb[tid] = ll_recv(mbox_slot(pl.mbox[me], par, 1, tid / kMboxFields, tid % kMboxFields), seq);
                    __threadfence();
                    __syncthreads();
                    double bv = -1.0;
                    for (int s = 0; s < R; ++s)
                        if (s_mb[s * kMboxFields] > bv) { bv = s_mb[s * kMboxFields]; q_pos = (int)s_mb[s * kMboxFields + 1]; rq = s_mb[s * kMboxFields + 2]; wq = s_mb[s * kMboxFields + 3]; }
                }
                if (gtid == 0) {
                    st->q_pos = q_pos;
                    st->q_var = __ldcg(lp.Nv + q_pos);
                    st->q_side = __ldcg(lp.Ns + q_pos);
                    st->rq = rq;
                    st->do_update = 0;
                    st->do_step = 0;
                }
            }
        }
        const int q_var = __ldcg(lp.Nv + q_pos);
        const int q_side = __ldcg(lp.Ns + q_pos);
        const bool at_lower = (q_side == ELLP_NB_LOWER);
        const int cnt = slot;
        if (tl) tl[2] = clock64();
        // ---- C1: the owner rebuilds the entering column of the CURRENT tableau and stores it into every OTHER rank's buffer; its own
        // copy goes straight to dcol (phase C2 walks the rows with the same thread mapping, so no store -> poll round trip through L2)
        const bool own_col = (q_pos >= lp.pos_lo && q_pos < lp.pos_lo + nT);
        double a_own = 0.;  // entry of the first row this thread owns (row == gtid)
        {
            const int ql = q_pos - lp.pos_lo;
            if (own_col) {
                __syncthreads();
                if (tid < cnt) s_vec[tid] = __ldcg(lp.V + (int64_t)tid * lp.ldv + ql);
                __syncthreads();
                for (int64_t i = gtid; i < lp.ld; i += gsize) {
                    const double a = corr_chain(__ldcg(lp.T + (int64_t)ql * lp.ld + i), lp.U + i, lp.ld, s_vec, cnt);
                    for (int d = 0; d < R; ++d)
                        if (d != me) ll_send(pl.col[d] + (int64_t)par * pl.col_cap + i, a, seq);
                    if (i == gtid) a_own = a; else lp.dcol[i] = a;
                }
            }
        }
        if (tl) tl[3] = clock64();
        // ---- C2 + D on the rank that OWNS the entering position only (R > 1; a single rank owns everything): column, direction,
        // ratios (primal :295-367), leaving row / bound flip (:305-434).  The other ranks do not wait for the column: they wait for
        // the owner's decision (three words in their mailbox, one NVLink hop after the owner's local ratio test) and pick the
        // column words up in phase E, by which time they have been in flight for the whole of the owner's phases C2 / D.  All
        // ranks then derive do_step / do_update / sides / "still running" from the same (lambda, row, pivot element).
        RowCache rc;
        rc.row = -1;
        const bool decide_here = (R == 1) || !pl.owner_only || (q_pos >= lp.pos_lo && q_pos < lp.pos_lo + nT);
        const int kq = lp.kind[q_var];  // :305-311, loaded before the barrier that phase D waits behind
        const double lambda0 = (kq == ELLP_TWOSIDED) ? (lp.ub[q_var] - lp.lb[q_var]) : (kq == ELLP_FIXED ? 0. : CUDART_INF);
        PivotDec dec;
        dec.q_pos = q_pos; dec.q_var = q_var; dec.q_side = q_side; dec.at_lower = at_lower; dec.rq = rq; dec.wq = wq;
        PivotRegs g_next = g;
        bool commit_here = false;
        // the decision from (lambda, candidate row or -1, column entry of that row): primal :402-434 + :205-232, replicated
        auto decide = [&](double lambda, int nb, double alpha_nb) {
            dec.lambda = lambda;
            dec.leave_var = -1;
            dec.r = nb;
            if (!(lambda >= 0.) || isinf(lambda)) {  // :402-406: ratio_commit records Unbounded (+ the assert)
                dec.do_step = 0; dec.do_update = 0; dec.alpha_r = 1.; dec.side_after = q_side;
                run = false;
            } else {
                dec.do_step = (lambda > 0.) ? 1 : 0;
                dec.do_update = (nb >= 0) ? 1 : 0;
                if (nb >= 0) {
                    dec.alpha_r = alpha_nb;  // = dcol[nb]
                    const double d_nb = at_lower ? -dec.alpha_r : dec.alpha_r;
                    dec.side_after = (d_nb > 0.) ? ELLP_NB_UPPER : ELLP_NB_LOWER;  // :208-221
                } else {  // :223-231 bound flip of the entering variable
                    dec.alpha_r = 1.;
                    dec.side_after = (q_side == ELLP_NB_LOWER) ? ELLP_NB_UPPER : (q_side == ELLP_NB_UPPER ? ELLP_NB_LOWER : ELLP_NB_FREE);
                    if (q_side == ELLP_NB_FREE) run = false;  // ratio_commit: kErrFlipFree
                }
                g_next.pivots = g.pivots + 1;
                g_next.trace_len = g.trace_len + 1;
                g_next.obj = g.obj + rq * (at_lower ? lambda : -lambda);
                if (g_next.pivots >= g.max_iter) run = false;  // ratio_commit: MaxIter at the next loop head
            }
            commit_here = (gtid == ((dec.do_update && nb >= 0) ? (int64_t)nb % gsize : 0));
        };
        if (decide_here) {
            {
                Top2 t{CUDART_INF, CUDART_INF, -1};
                const uint4* colbuf = pl.col[me] + (int64_t)par * pl.col_cap;
                for (int64_t i = gtid; i < lp.ld; i += gsize) {
                    // Bv -> x / bounds do not depend on the column: walk them while the column word is in flight
                    int var = 0;
                    double xv = 0., lbv = 0., ubv = 0.;
                    int kv = ELLP_FIXED;
                    if (i < m) {
                        var = __ldcg(lp.Bv + i);
                        xv = __ldcg(lp.x + var);
                        kv = lp.kind[var]; lbv = lp.lb[var]; ubv = lp.ub[var];
                    }
                    const double a = own_col ? (i == gtid ? a_own : __ldcg(lp.dcol + i)) : ll_recv(colbuf + i, seq);
                    lp.dcol[i] = a;
                    if (i == gtid) { rc.row = i; rc.a = a; rc.var = var; rc.xv = xv; }
                    if (i < m) {
                        const double d_i = at_lower ? -a : a;  // :296-300
                        double lam = kLamSkipped;               // skipped (|d_i| < EPS, :321)
                        if (!(fabs(d_i) < kEps)) lam = primal_ratio(kv, lbv, ubv, xv, d_i);
                        lp.lam[i] = lam;
                        if (lam != kLamSkipped && lam < CUDART_INF) top2_push<false>(t, lam, (int)i);
                    }
                }
                t = top2_block_fast<false>(t, &s_top, rbuf);
                ll_publish(llC, par, t, (t.i1 >= 0) ? __ldcg(lp.dcol + t.i1) : 1., seq);  // extra = column entry of the block's best row
            }
            if (tl) tl[4] = clock64();
            ll_gather(llC, par, G, s_part, seq);  // replaces the grid barrier: every block waits for every block's ratios
            if (tl) tl[5] = clock64();
            double alpha_best;
            Top2 t = (G <= 32) ? ll_reduce_w<false>(s_part, G, &alpha_best) : ll_reduce<false>(s_part, G, &s_top, rbuf, &s_extra, &alpha_best);
            const double lmin_basic = t.a1;
            if (lambda0 < CUDART_INF) {
                if (lambda0 < t.a1) { t.a2 = t.a1; t.a1 = lambda0; t.i1 = -1; }
                else if (lambda0 < t.a2) t.a2 = lambda0;
            }
            const bool fast = !(t.a1 < CUDART_INF) || !(t.a2 < t.a1 + 2. * kEps);
            double msg_lambda, msg_alpha;
            int msg_nb;
            if (fast) {
                decide(t.a1, t.i1, alpha_best);
                msg_lambda = t.a1; msg_nb = t.i1; msg_alpha = alpha_best;
            } else {
                __syncthreads();
                if (blockIdx.x == 0) {
                    if (tid == 0) st->lmin_bits = __double_as_longlong(lmin_basic);
                    __syncthreads();
                    ratio_pick_body(lp, tie_rule, st, scan_smem);  // exact fold of :379-399, commits through ratio_commit
                }
                grid.sync();
                dec.do_step = __ldcg(&st->do_step);
                dec.do_update = __ldcg(&st->do_update);
                dec.r = __ldcg(&st->r_pos);
                dec.leave_var = dec.do_update ? __ldcg(&st->leave_var) : -1;
                dec.lambda = __ldcg(&st->step);
                dec.alpha_r = __ldcg(&st->alpha_r);
                dec.side_after = __ldcg(lp.Ns + q_pos);
                run = (__ldcg(&st->status) == kRunning);
                pivot_regs_load(g_next, st);
                g = g_next;  // the tie path committed through PivotState before E
                msg_lambda = (__ldcg(&st->err) == kErrLambdaNegative) ? -1.0 : dec.lambda;  // the other ranks raise the same assert
                msg_nb = dec.do_update ? dec.r : -1;
                msg_alpha = dec.alpha_r;
            }
            if (R > 1 && pl.owner_only && blockIdx.x == 0 && tid < R * 3) {  // the decision to every other rank
                const int dst = tid / 3, f = tid % 3;
                if (dst != me) ll_send(mbox_slot(pl.mbox[dst], par, 2, 0, f), f == 0 ? msg_lambda : (f == 1 ? (double)msg_nb : msg_alpha), seq);
            }
        } else {
            if (tl) { tl[4] = clock64(); tl[5] = tl[4]; }
            __syncthreads();
            if (tid < 3) s_mb[tid] = ll_recv(mbox_slot(pl.mbox[me], par, 2, 0, tid), seq);
            __threadfence();
            __syncthreads();
            decide(s_mb[0], (int)s_mb[1], s_mb[2]);
            __syncthreads();  // s_mb is reused by the next pivot's pricing merge
        }
        // ---- E + A': step, bookkeeping, local pivot row, reduced costs, new slot, keys of the next pivot
        const bool price = run && (slot + 1 < slot0 + npiv);
        if (tl) tl[6] = clock64();
        if (rc.row >= m) { rc.var = 0; rc.xv = 0.; }
        const uint4* colsrc = decide_here ? nullptr : pl.col[me] + (int64_t)par * pl.col_cap;  // not a deciding rank: the column is still in its LL words
        Top2 t = blk_row_price_body<DEVEX>(lp, slot, st, dec, g, commit_here, price, gtid, gsize, s_vec, rc, colsrc, seq);
        g = g_next;
        priced = false;
        if (tl) tl[7] = clock64();
        if (price) {
            t = top2_block_fast<true>(t, &s_top, rbuf);
            ll_publish(llA, par ^ 1, t, (t.i1 >= 0) ? __ldcg(lp.dj + t.i1) : 0., seq + 1u);  // phase A of the next pivot
            priced = true;
        }
        if (tl) tl[8] = clock64();
    }
}

// reduced costs of the locally stored positions of a fresh tableau: dj holds c_B^T (B^-1 a_p) on entry
__global__ void k_redcost_pos_local(const double* __restrict__ c, const int32_t* __restrict__ Nv, int pos_lo, int nT, double* __restrict__ dj) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nT) dj[p] = c[Nv[pos_lo + p]] - dj[p];
}

}  // namespace ellp
