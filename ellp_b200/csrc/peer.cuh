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
// Per pivot and rank (slot = index of the pending (U, V) pair, cf. blocked.cuh):
//   A   Dantzig keys of the local positions, per-block (best, second best); the LAST block to arrive (atomic ticket)
//       merges them and stores (best key, second key, global position, reduced cost) into every rank's mailbox
//   B   every block polls the G mailbox entries and merges them in rank order => the same entering position on every
//       rank without a grid barrier (the mailbox wait IS the barrier).  Near-tie (best - second < 2 EPS, any two
//       ranks): second round with the order-free rule of SURVEY appendix A.1 (largest variable index within EPS of the
//       global maximum), again one mailbox entry per rank.
//   C1  the owner of the entering position rebuilds that column of the current tableau (stale column + pending
//       corrections, same arithmetic as k_ratio_prep) and stores it into every rank's column buffer
//   C2  every rank polls the column (row i by the thread that needs row i), ratios, per-block two smallest | grid barrier
//   D   merge, entering variable's own range, commit (ratio_commit) or the exact tie fold (ratio_pick_body) | grid barrier
//   E   x step, local part of the pivot row and of the reduced-cost row, new (U, V) slot (blk_row_body).  No barrier:
//       the next pivot's phase A/B separates it from the next reader.
// Reference lines: pricing primal_simplex_solver.rs:189,253-292; column + ratios :295-367; fold/step/apply :379-434,
// :205-232.  Decisions equal those of the NCCL path and of the oracle's canonical mode pivot for pivot (tests).
#pragma once
#include <cooperative_groups.h>
#include "blocked.cuh"

namespace ellp {

constexpr int kMaxPeers = 8;
constexpr long long kPeerTimeoutNs = 8000000000ll;  // 8 s
constexpr int kMboxFields = 4;
constexpr int kMboxWords = 2 /*parity*/ * 2 /*round*/ * kMaxPeers * kMboxFields;  // uint4 words per rank

struct PeerLinks {
    int32_t rank, nranks;
    int64_t col_cap;         // rows per parity slot of a column buffer
    uint4* mbox[kMaxPeers];  // mbox[r] = rank r's mailbox (peer-mapped unless r == rank)
    uint4* col[kMaxPeers];   // col[r]  = rank r's column buffer, 2 * col_cap words
    unsigned int* ticket;    // local: arrival counters of the last-block pattern (2)
};

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
    return base + (((par * 2 + round) * kMaxPeers + src) * kMboxFields + field);
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

__global__ void __launch_bounds__(kScanThreads, 1) k_blk_pivots_peer(DevLP lp, PeerLinks pl, int tie_rule, int slot0, int npiv, uint32_t seq0,
                                                                     PivotState* st) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char scan_smem[];
    __shared__ Top2Smem s_top;
    __shared__ double s_vec[kBlkMax];
    __shared__ double s_mb[kMaxPeers * kMboxFields];
    __shared__ long long s_ll[32];
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + tid, gsize = (int64_t)gridDim.x * blockDim.x;
    const int G = gridDim.x, R = pl.nranks, me = pl.rank;
    double* partA = lp.coop;
    double* partC = lp.coop + 3 * 1024;
    long long* partT = reinterpret_cast<long long*>(lp.coop);  // near-tie round: per-block (variable << 32 | local position), reuses partA
    const int nT = lp.nT, m = lp.m;
    bool run = (__ldcg(&st->status) == kRunning);
    for (int slot = slot0; slot < slot0 + npiv; ++slot) {
        if (!run) { blk_zero_slot(lp, slot, gtid, gsize); continue; }
        const uint32_t seq = seq0 + (uint32_t)(slot - slot0) + 1u;
        const int par = (int)(seq & 1u);
        // ---- A: pricing of the local positions (primal :189, :253-270)
        {
            Top2 t{-1.0, -1.0, -1};
            for (int64_t j = gtid; j < nT; j += gsize) {
                const double r = __ldcg(lp.dj + j);
                const int side = __ldcg(lp.Ns + lp.pos_lo + j);
                double k = -1.0;
                if (!(fabs(r) < kEps)) {
                    if (r > 0. && side == ELLP_NB_UPPER) k = r;
                    else if (!(r > 0.) && side == ELLP_NB_LOWER) k = -r;
                    else if (side == ELLP_NB_FREE) k = fabs(r);
                }
                lp.key[j] = k;
                lp.rN[j] = r;
                if (k != -1.0) top2_push<true>(t, k, (int)j);
            }
            t = top2_block<true>(t, &s_top);
            if (tid == 0) { partA[3 * blockIdx.x] = t.a1; partA[3 * blockIdx.x + 1] = t.a2; partA[3 * blockIdx.x + 2] = (double)t.i1; }
            if (peer_arrive_last(pl.ticket, &s_flag)) {
                const Top2 loc = top2_grid<true>(partA, G, &s_top);
                if (tid < R * kMboxFields) {
                    const int dst = tid / kMboxFields, f = tid % kMboxFields;
                    double v;
                    if (f == 0) v = loc.a1;
                    else if (f == 1) v = loc.a2;
                    else if (f == 2) v = (loc.i1 >= 0) ? (double)(lp.pos_lo + loc.i1) : -1.0;
                    else v = (loc.i1 >= 0) ? __ldcg(lp.dj + loc.i1) : 0.;
                    ll_send(mbox_slot(pl.mbox[dst], par, 0, me, f), v, seq);
                }
            }
        }
        // ---- B: entering position, identical on every rank (primal :271-292, order-free tie rule)
        int q_pos;
        double rq;
        {
            if (tid < R * kMboxFields) s_mb[tid] = ll_recv(mbox_slot(pl.mbox[me], par, 0, tid / kMboxFields, tid % kMboxFields), seq);
            __threadfence();
            __syncthreads();
            Top2 t{-1.0, -1.0, -1};
            rq = 0.;
            for (int s = 0; s < R; ++s) {
                const Top2 o{s_mb[s * kMboxFields], s_mb[s * kMboxFields + 1], (int)s_mb[s * kMboxFields + 2]};
                if (o.a1 > t.a1) rq = s_mb[s * kMboxFields + 3];
                top2_merge<true>(t, o);
            }
            if (t.a1 == -1.0) {  // no candidate on any rank: optimal (:289-292)
                if (gtid == 0) { st->status = ELLP_OPTIMAL; st->do_update = 0; st->do_step = 0; }
                run = false;
                blk_zero_slot(lp, slot, gtid, gsize);
                continue;
            }
            q_pos = t.i1;
            if (t.a1 - t.a2 < 2. * kEps) {
                // near-tie: largest variable index among the keys within EPS of the global maximum
                const double kmax = t.a1;
                long long best = -1;
                for (int64_t j = gtid; j < nT; j += gsize) {
                    const double k = lp.key[j];  // written by this thread in phase A
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
                    if (tid < R * 3) {
                        const int dst = tid / 3, f = tid % 3;
                        const int jl = (int)(b & 0xffffffffll);
                        double v;
                        if (f == 0) v = (b >= 0) ? (double)(b >> 32) : -1.0;
                        else if (f == 1) v = (b >= 0) ? (double)(lp.pos_lo + jl) : -1.0;
                        else v = (b >= 0) ? __ldcg(lp.dj + jl) : 0.;
                        ll_send(mbox_slot(pl.mbox[dst], par, 1, me, f), v, seq);
                    }
                }
                __syncthreads();
                if (tid < R * 3) s_mb[(tid / 3) * kMboxFields + (tid % 3)] = ll_recv(mbox_slot(pl.mbox[me], par, 1, tid / 3, tid % 3), seq);
                __threadfence();
                __syncthreads();
                double bv = -1.0;
                for (int s = 0; s < R; ++s)
                    if (s_mb[s * kMboxFields] > bv) { bv = s_mb[s * kMboxFields]; q_pos = (int)s_mb[s * kMboxFields + 1]; rq = s_mb[s * kMboxFields + 2]; }
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
        const int q_var = __ldcg(lp.Nv + q_pos);
        const bool at_lower = (__ldcg(lp.Ns + q_pos) == ELLP_NB_LOWER);
        const int cnt = slot;  // pending slots of this block of pivots
        // ---- C1: the owner rebuilds the entering column of the CURRENT tableau and stores it into every rank's buffer
        {
            const int ql = q_pos - lp.pos_lo;
            if (ql >= 0 && ql < nT) {
                __syncthreads();
                if (tid < cnt) s_vec[tid] = __ldcg(lp.V + (int64_t)tid * lp.ldv + ql);
                __syncthreads();
                for (int64_t i = gtid; i < lp.ld; i += gsize) {
                    double a = __ldcg(lp.T + (int64_t)ql * lp.ld + i);
                    for (int j = 0; j < cnt; ++j) a = fma(-__ldcg(lp.U + (int64_t)j * lp.ld + i), s_vec[j], a);
                    for (int d = 0; d < R; ++d) ll_send(pl.col[d] + (int64_t)par * pl.col_cap + i, a, seq);
                }
            }
        }
        // ---- C2: every rank: column, direction, ratios (primal :295-367)
        double lmin_basic;
        {
            Top2 t{CUDART_INF, CUDART_INF, -1};
            const uint4* colbuf = pl.col[me] + (int64_t)par * pl.col_cap;
            for (int64_t i = gtid; i < lp.ld; i += gsize) {
                const double a = ll_recv(colbuf + i, seq);
                lp.dcol[i] = a;
                if (i < m) {
                    const double d_i = at_lower ? -a : a;  // :296-300
                    double lam = -1.0;                      // -1 = skipped (|d_i| < EPS, :321)
                    if (!(fabs(d_i) < kEps)) {
                        const int var = __ldcg(lp.Bv + i);
                        lam = primal_ratio(lp.kind[var], lp.lb[var], lp.ub[var], __ldcg(lp.x + var), d_i);
                    }
                    lp.lam[i] = lam;
                    if (lam != -1.0 && lam < CUDART_INF) top2_push<false>(t, lam, (int)i);
                }
            }
            t = top2_block<false>(t, &s_top);
            if (tid == 0) { partC[3 * blockIdx.x] = t.a1; partC[3 * blockIdx.x + 1] = t.a2; partC[3 * blockIdx.x + 2] = (double)t.i1; }
        }
        grid.sync();
        // ---- D: leaving row / bound flip (primal :305-434, :205-232); replicated, bit-identical on every rank
        {
            Top2 t = top2_grid<false>(partC, G, &s_top);
            lmin_basic = t.a1;
            const int kq = lp.kind[q_var];  // :305-311
            const double lambda0 = (kq == ELLP_TWOSIDED) ? (lp.ub[q_var] - lp.lb[q_var]) : (kq == ELLP_FIXED ? 0. : CUDART_INF);
            if (lambda0 < CUDART_INF) {
                if (lambda0 < t.a1) { t.a2 = t.a1; t.a1 = lambda0; t.i1 = -1; }
                else if (lambda0 < t.a2) t.a2 = lambda0;
            }
            const bool fast = !(t.a1 < CUDART_INF) || !(t.a2 < t.a1 + 2. * kEps);
            if (fast) {
                if (gtid == 0) ratio_commit(lp, st, t.i1, t.a1, at_lower, q_var);
            } else if (blockIdx.x == 0) {
                if (tid == 0) st->lmin_bits = __double_as_longlong(lmin_basic);
                __syncthreads();
                ratio_pick_body(lp, tie_rule, st, scan_smem);
            }
        }
        grid.sync();
        // ---- E: step, local pivot row, reduced costs, new slot (primal :408-417 + the deferred row reduction)
        blk_row_body(lp, slot, st, gtid, gsize, s_vec);
        run = (__ldcg(&st->status) == kRunning);
    }
}

// reduced costs of the locally stored positions of a fresh tableau: dj holds c_B^T (B^-1 a_p) on entry
__global__ void k_redcost_pos_local(const double* __restrict__ c, const int32_t* __restrict__ Nv, int pos_lo, int nT, double* __restrict__ dj) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nT) dj[p] = c[Nv[pos_lo + p]] - dj[p];
}

}  // namespace ellp
