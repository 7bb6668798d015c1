// engine.cu -- host driver of the device-resident pivot loop + the C ABI of include/ellp_b200.h.
//
// The host only launches kernels and reads back the 16-byte index-level status; every number of the
// iteration (B^-1, x, y, d, reduced costs, ratios) stays in HBM.  No CPU fallback exists: without a
// CUDA device ellp_b200_create() fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "engine.hpp"
#include "kernels.cuh"
#include "batch.cuh"
#include "refactor.cuh"
#include "blocked.cuh"
#include "peer.cuh"
#include "dual_blocked.cuh"
#include "small.cuh"

using namespace ellp;

#define CUDA_TRY(expr)                                                                           \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                      \
            return ELLP_E_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define LAUNCH(kernel, grid, block, ...)                                                         \
    do {                                                                                         \
        kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__);                                \
        ctx->launches++;                                                                         \
    } while (0)

#define LAUNCH_SMEM(kernel, grid, block, smem, ...)                                              \
    do {                                                                                         \
        kernel<<<(grid), (block), (smem), ctx->stream>>>(__VA_ARGS__);                           \
        ctx->launches++;                                                                         \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Arena {
    size_t off = 0;
    char* base = nullptr;
    template <class T> T* take(size_t count) {
        off = align_up(off, 256);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct ellp_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    char* arena = nullptr;
    size_t arena_bytes = 0;
    bool resident = false;
    int solver = ELLP_PRIMAL;
    DevLP lp{};
    int KS = 1, kc = 64;
    bool binv_valid = false;
    bool tableau = false;  // ELLP_ENGINE_TABLEAU resident
    // column sharding (one rank per GPU)
    bool sharded = false;
    int rank = 0, nranks = 1;
    void* nccl_comm = nullptr;
    double* sendcol = nullptr;  // ld doubles staged for the pivot-column all-reduce (inside the arena)
    uint8_t* d_sides = nullptr; // n_glob bytes: colstat of every column (gathered) / scatter source
    int* d_flag = nullptr;
    SelScratch* d_sel = nullptr;   // scratch of the multi-block selection kernels (kernels.cuh)
    // CUDA graph of `graph_iters` iterations of the revised engine (the 8-9 dependent launches of an iteration are
    // launch-latency bound: replaying them from a graph removes most of the gaps).  Rebuilt when the resident LP or the
    // rules change; not used with profile = 1 (events around the row reduction).
    cudaGraphExec_t graph_exec = nullptr;
    int graph_iters = 0;
    uint64_t graph_key = 0, graph_launches = 0;
    uint64_t lp_generation = 0;   // bumped by every upload / generate
    int use_graphs = 1;           // tuning key "cuda_graphs"
    int small_path = 1;           // tuning key "small_path": single-CTA kernels for netlib-sized LPs on the revised engine (small.cuh)
    bool small_attr_set = false;
    uint64_t pivots_since_refactor = 0;
    PivotState* d_st = nullptr;
    PivotState* h_st = nullptr;  // pinned
    int64_t trace_cap = 0;
    double dual_obj0 = 0.;
    std::vector<cudaEvent_t> ev;  // profile=1: pairs around rank-1 launches
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // K6 batch of small LPs (device resident)
    struct Batch {
        int nlp = 0, m = 0, n0 = 0, nc = 0, ld = 0, trace_cap = 0;
        double *A = nullptr, *c = nullptr, *b = nullptr, *lb = nullptr, *ub = nullptr, *x = nullptr, *obj = nullptr;
        uint8_t *kind = nullptr, *Ns = nullptr;
        int32_t *B = nullptr, *N = nullptr, *status = nullptr, *iters = nullptr, *err = nullptr, *trace_len = nullptr;
        ellp_trace_rec* trace = nullptr;
    } batch;
    // tuning (ellp_b200_set_tuning)
    int rank1_cols_per_cta = 8;
    int rank1_stream_min_mb = 96;
    int blk_kmax = 0;             // slots allocated for the blocked (deferred rank-k) tableau engine; 0 = rank-1 engine only
    int blk_fill = 0;             // slots used since the last flush
    int flush_col_steps = 8;      // column steps (of 64 columns) per CTA of k_blk_flush
    int flush_kernel = 0;         // tuning: 0 = auto (see launch_rankk), 1 = k_blk_flush (2 CTAs/SM, also the fallback for an unpadded V),
                                  // 3 = k_blk_flush3 (register prefetch + bulk-copy ring), 4 = k_blk_flush4 (16 consumer warps), 5 / 6 = k_blk_flush5<2 / 4>,
                                  // 7 / 8 = k_blk_flush6<1 / 2> (no producer warp), 9 = k_blk_flush4r<3> (12 consumer warps, 128 registers)
    int last_flush_kernel = 0;    // version launch_rankk launched last (ellp_b200_last_flush_kernel)
    int flush4_min_k = 24;        // auto: the wide kernels (versions 4r / 6) from this many pending pairs on, version 3 below
    int flush_ld = -1;            // tuning key "flush_ld": tile access mode of versions 4 / 5 (-1 = auto, see ld_tile in blocked.cuh)
    int flush_stages = 0;         // tuning key "flush_stages": ring depth of versions 4 / 5 (0 = blk_flush4_stages)
    bool flush_attrs_set = false;
    int flush2_col_steps = 32;    // column steps per CTA of k_blk_flush3 / k_blk_flush4
    int flush_waves = 6;          // tuning key "flush_waves": keep at least this many waves of CTAs (narrow shards); 0 = take flush2_col_steps as given.
                                  // Sweep on the shard shapes 32768 x {4096, 8192, 16384} (profiles/r02_flush_shard_sweep.jsonl): ~1000 CTAs (7 waves) is the
                                  // optimum for both kernels; the round-1 value 8 cut narrow shards into 14 waves of short CTAs (0.694 vs 0.618 ms at 32768 x 4096, k = 56)
    int coop_pivots = 1;          // blocked engine: 1 = k_blk_pivots_fused (one cooperative launch per block of pivots), 0 = five kernels per pivot
    // peer-memory sharded engine (peer.cuh): condensed tableau split by nonbasic position, exchange fused into the pivot kernel
    bool a_resident = true;       // false after the condensed fast upload: only T = A_N is on the device
    bool peer_mode = false;       // the resident LP uses the peer layout
    int peer_exchange = 1;        // tuning: sharded + block_k > 1 => peer layout (1) or the NCCL path on the full tableau (0)
    PeerLinks pl{};               // peer-mapped mailboxes / column buffers (passed to the kernel by value)
    void* peer_own = nullptr;     // this rank's exchange buffer (cudaMalloc, exported with cudaIpcGetMemHandle)
    void* peer_map[kMaxPeers] = {nullptr};  // cudaIpcOpenMemHandle mappings of the other ranks' buffers
    int64_t peer_cap = 0;         // rows per parity slot of the column buffers
    uint32_t xseq = 0;            // pivots exchanged since the communicator was created (wire sequence number)
    int coop_grid[4] = {0, 0, 0, 0}, coop_threads_cached[4] = {0, 0, 0, 0};  // k_blk_pivots_fused<false / true>, k_blk_dual_pivots_fused<false / true>
    // tableau engines with a general (non-identity) starting basis: B^-1 of the basis the tableau was built from lives in
    // lp.G / lp.Binv (has_binv), with the scratch the blocked LU needs next to the tableau's own U / V
    bool has_binv = false;
    double* rf_V = nullptr;       // kPanel x rf_ldv block row of the LU
    int64_t rf_ldv = 0;
    double* rf_coop = nullptr;    // publication slots of the cooperative panel kernel
    cudaStream_t copy_stream = nullptr;   // H2D of the pipelined batch path (ellp_b200_primal_solve_batch)
    cudaStream_t d2h_stream = nullptr;    // D2H of the same
    std::vector<cudaEvent_t> chunk_ev;
    int residual_every = -1;      // tuning key "residual_every": pivots between checks of |A x - b| on rebuildable tableau LPs (-1 = default: 1024 for m > 512, 0 = off)
    double residual_tol = 1e-9;   // relative to 1 + |b|_inf (tuning key "residual_tol_1e12": tolerance in units of 1e-12)
    double last_residual = 0.;
    double b_inf = -1.;           // |b|_inf of the resident LP (computed on first use)
    int fast_upload = 1;          // tuning key "fast_upload": 0 = always upload the whole A (keeps the LP rebuildable: refactor_every works)
    int owner_ratio = 0;          // tuning key "owner_ratio": 0 every rank runs the primal ratio test (default), 1 only the owner of the entering column + decision broadcast
    int batch_pipeline = 1;       // tuning key "batch_pipeline": 0 = upload, run, download one after the other
    bool recompute_x = false;     // set by ellp_b200_run around mid-solve rebuilds of the tableau (not at the start of a run: the caller's x is authoritative)
    bool devex_live = false;      // lp.w holds Devex reference weights of the resident solve (reset by upload / generate / refactor)
    bool dj_live = false;         // dual on the tableau: dj (not lp.d) holds the current reduced costs of the nonbasic positions
    bool tab_from_binv = false;   // T was built as B^-1 A_N (y at download = B^-T (c_B0 - d_B0)); false: diagonal starting basis (bscale)
    long long* tlog = nullptr;    // phase-timing log of k_blk_pivots_fused (tuning key "phase_timing")
    int tlog_cap = 0;
    uint32_t tlog_seq0 = 0;
    int coop_threads = 256;       // tuning: threads per block of k_blk_pivots_fused (64..512)
    int coop_ctas_per_sm = 1;     // tuning: resident blocks per SM the fused kernel may use
    int lu_panel_grid = 0;        // co-resident CTAs of k_lu_panel_coop (0 = not yet queried, -1 = unavailable)
    int refactor_panel = 0;       // tuning: 1 = single-CTA panel kernel
    int refactor_mode = 0;        // 0 auto (blocked LU + DMMA for m >= 128, Gauss-Jordan below), 1 Gauss-Jordan, 2 blocked LU  // evict-first policy when the updated matrix is larger than this
};

extern "C" { static void batch_free(ellp_b200_ctx* ctx); }
static void peer_release(ellp_b200_ctx* ctx);

// ---- NCCL, bound at run time (torch ships libnccl.so.2; the library must also load on boxes without it) -----------
namespace nccl {
struct UniqueId { char internal[128]; };
typedef void* Comm;
enum { kUint8 = 1, kFloat64 = 8, kSum = 0 };
struct Api {
    void* handle = nullptr;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static Api api;
static bool load(const char* path, std::string* err) {
    if (api.handle) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h && path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { *err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    api.handle = h;
    api.GetUniqueId = (int (*)(UniqueId*))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(Comm*, int, UniqueId, int))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (int (*)(Comm))dlsym(h, "ncclCommDestroy");
    api.AllGather = (int (*)(const void*, void*, size_t, int, Comm, cudaStream_t))dlsym(h, "ncclAllGather");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, Comm, cudaStream_t))dlsym(h, "ncclAllReduce");
    api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.AllReduce) { *err = "libnccl.so.2 lacks required symbols"; api.handle = nullptr; return false; }
    return true;
}
}  // namespace nccl

#define NCCL_TRY(expr)                                                                                       \
    do {                                                                                                     \
        int r__ = (expr);                                                                                    \
        if (r__ != 0) {                                                                                      \
            ctx->err = std::string(#expr) + ": " + (nccl::api.GetErrorString ? nccl::api.GetErrorString(r__) : "nccl error"); \
            return ELLP_E_CUDA;                                                                              \
        }                                                                                                    \
    } while (0)

namespace {

int set_err(ellp_b200_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

int ensure_arena(ellp_b200_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->arena_bytes) return ELLP_OK;
    if (ctx->arena) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaFree(ctx->arena));
        ctx->arena = nullptr;
        ctx->arena_bytes = 0;
    }
    bytes = align_up(bytes + (bytes >> 3), (size_t)1 << 21);
    CUDA_TRY(cudaMalloc(&ctx->arena, bytes));
    ctx->arena_bytes = bytes;
    return ELLP_OK;
}

// true iff column B[i] of the column-major m x n matrix A is a multiple d_i e_i (d_i != 0) of the unit vector e_i for every basis
// position i (slack bases: +1 for Lte / Eq artificial columns, -1 for Gte rows); diag receives d.  Memory-bound scan of m^2
// doubles, split over host threads (it overlaps the DMA of the nonbasic columns).
bool host_basis_is_diagonal(const double* A, int m, const int32_t* B, double* diag, bool* all_ones) {
    const size_t total = (size_t)m * m;
    unsigned nthreads = 1;
    if (total >= ((size_t)1 << 22)) nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::atomic<bool> ok{true}, ones{true};
    auto work = [&](int i0, int i1) {
        for (int i = i0; i < i1 && ok.load(std::memory_order_relaxed); ++i) {
            const double* col = A + (size_t)B[i] * m;
            const double d = col[i];
            const bool good = (d != 0.0) && (d == d) && !std::isinf(d);
            double acc = 0.;
            for (int k = 0; k < m; ++k) acc += (col[k] != 0.0) ? 1.0 : 0.0;  // branch-free count of nonzeros (NaN counts too)
            if (!good || acc != 1.0) ok.store(false, std::memory_order_relaxed);
            if (d != 1.0) ones.store(false, std::memory_order_relaxed);
            diag[i] = d;
        }
    };
    if (nthreads == 1) { work(0, m); }
    else {
        std::vector<std::thread> th;
        const int chunk = (m + (int)nthreads - 1) / (int)nthreads;
        for (unsigned t = 0; t < nthreads; ++t) {
            const int i0 = (int)t * chunk, i1 = std::min(m, i0 + chunk);
            if (i0 < i1) th.emplace_back(work, i0, i1);
        }
        for (auto& t : th) t.join();
    }
    *all_ones = ones.load();
    return ok.load();
}

struct RefactorScratch {
    double* V = nullptr;
    int64_t ldv = 0;
    double* coop = nullptr;
};

void carve(Arena& a, DevLP& lp, int KS, int64_t trace_cap, bool tableau, int blk_kmax, bool sharded = false, int nranks = 1,
           double** sendcol = nullptr, uint8_t** d_sides = nullptr, bool with_binv = false, RefactorScratch* rf = nullptr) {
    // n = locally stored columns (A / T, dj, key, prow); ng = length of the replicated per-variable vectors
    const size_t ld = (size_t)lp.ld, m = (size_t)lp.m, n = (size_t)lp.n, ng = (size_t)lp.n_glob;
    const size_t nN = sharded ? n : (size_t)lp.nN;
    lp.A = a.take<double>(ld * n);
    lp.c = a.take<double>(ng);
    lp.b = a.take<double>(std::max<size_t>(m, 1));
    lp.lb = a.take<double>(ng);
    lp.ub = a.take<double>(ng);
    lp.kind = a.take<uint8_t>(ng);
    lp.x = a.take<double>(ng);
    lp.Bv = a.take<int32_t>(std::max<size_t>(m, 1));
    lp.Nv = a.take<int32_t>(std::max<size_t>(nN, 1));
    lp.Ns = a.take<uint8_t>(std::max<size_t>(nN, 1));
    lp.y = a.take<double>(ld);
    lp.d = a.take<double>(ng);
    if (tableau) {  // condensed: T is separate from A; NCCL-sharded: T overwrites A in place
        lp.G = nullptr;
        lp.Binv = nullptr;
        if (with_binv && !sharded) {  // general starting basis: B^-1 (revised-engine refactorisation), then T = B^-1 A_N
            lp.G = a.take<double>(ld * 2 * m);
            lp.Binv = lp.G ? lp.G + ld * m : nullptr;
            RefactorScratch r;
            r.ldv = (int64_t)align_up(std::max(2 * m, nN), 128);
            r.V = a.take<double>((size_t)kPanel * (size_t)r.ldv);
            r.coop = a.take<double>(lu_panel_pub_doubles(160));
            if (rf) *rf = r;
        }
        lp.Bv0 = a.take<int32_t>(std::max<size_t>(m, 1));
        lp.bscale = a.take<double>(ld);
        lp.dpos = a.take<double>(std::max<size_t>(nN, 1));
        lp.wN = a.take<double>(std::max<size_t>(nN, 1));
        lp.xpart = a.take<double>(2 * 32 * ld);
        lp.condensed = sharded ? 0 : 1;
        lp.nT = lp.condensed ? (int32_t)nN : (int32_t)n;
        lp.T = lp.condensed ? a.take<double>(ld * std::max<size_t>(nN, 1)) : const_cast<double*>(lp.A);
        lp.dj = a.take<double>(n);
        lp.ldv = (int64_t)align_up((size_t)std::max<int32_t>(lp.nT, 1), 128);  // whole 64-column tiles: k_blk_flush3 bulk-copies V rows unguarded
        lp.coop = a.take<double>(6 * 1024);
        lp.U = blk_kmax > 0 ? a.take<double>(ld * (size_t)blk_kmax) : nullptr;
        lp.V = blk_kmax > 0 ? a.take<double>((size_t)lp.ldv * (size_t)blk_kmax) : nullptr;
    } else {
        // revised engine: V = row-major block row of the blocked LU (operand of the rank-kPanel trailing update),
        // coop = publication slots of the cooperative panel factorisation (refactor.cuh)
        lp.U = nullptr;
        lp.ldv = (int64_t)align_up(2 * m, 128);
        lp.V = a.take<double>((size_t)kPanel * (size_t)lp.ldv);
        lp.coop = a.take<double>(lu_panel_pub_doubles(160));
        lp.condensed = 0;
        lp.nT = 0;
        lp.G = a.take<double>(ld * 2 * m);
        lp.Binv = lp.G ? lp.G + ld * m : nullptr;
        lp.T = nullptr;
        lp.dj = nullptr;
    }
    lp.cB = a.take<double>(ld);
    lp.u = a.take<double>(ld);
    lp.rN = a.take<double>(std::max<size_t>(nN, 1));
    lp.key = a.take<double>(std::max<size_t>(nN, 1));
    lp.dcol = a.take<double>(ld);
    lp.rho = a.take<double>(ld);
    lp.prow = a.take<double>(std::max(tableau ? n : 2 * m, (tableau && with_binv) ? 2 * m : (size_t)0) + 8);
    lp.part = a.take<double>(tableau ? 8 : (size_t)KS * ld);
    lp.lam = a.take<double>(std::max<size_t>(m, 1));
    lp.lu_piv = a.take<int32_t>(std::max<size_t>(m, 1));
    if (!tableau) {
        lp.w = a.take<double>(ld);
        lp.npart = a.take<double>((m / kNormCols + 2) * ld);
    } else {
        lp.w = a.take<double>(ld);
        lp.npart = nullptr;
    }
    lp.trace = trace_cap > 0 ? a.take<ellp_trace_rec>((size_t)trace_cap) : nullptr;
    if (sharded) {
        lp.colstat = a.take<uint8_t>(n);
        lp.xchg = a.take<double>(64 + 4 * (size_t)nranks);
        double* sc = a.take<double>(ld);
        uint8_t* sd = a.take<uint8_t>(ng);
        if (sendcol) *sendcol = sc;
        if (d_sides) *d_sides = sd;
    } else {
        lp.colstat = nullptr;
        lp.xchg = nullptr;
    }
}

// slots of the blocked (deferred rank-k) tableau engine requested by the caller (ellp_opts::block_k), 0 = rank-1 engine
int blk_slots(const ellp_opts* o, bool tableau, int solver = ELLP_PRIMAL) {
    if (!tableau || !o) return 0;
    if (o->block_k <= 1) return solver == ELLP_DUAL ? 32 : 0;  // the dual runs only on the blocked engine (dual_blocked.cuh)
    return std::min<int>(o->block_k, kBlkMax);
}

const char* dev_err_message(int e) {
    switch (e) {
        case kErrNaNPricing: return "NaN detected";
        case kErrLambdaNegative: return "assertion failed: lambda >= 0.";
        case kErrFlipFree: return "pivot should have been unbounded";
        case kErrNaNDualRatio: return "called `Option::unwrap()` on a `None` value (partial_cmp)";
        case kErrSingular: return "invalid B, A_B is not invertible";
        default: return "unknown device error";
    }
}

int read_state(ellp_b200_ctx* ctx) {
    CUDA_TRY(cudaMemcpyAsync(ctx->h_st, ctx->d_st, sizeof(PivotState), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ELLP_OK;
}

int write_state(ellp_b200_ctx* ctx) {
    CUDA_TRY(cudaMemcpyAsync(ctx->d_st, ctx->h_st, sizeof(PivotState), cudaMemcpyHostToDevice, ctx->stream));
    return ELLP_OK;
}

void launch_rank1(ellp_b200_ctx* ctx, double* E, int64_t ld, int R, int C, const double* alpha, const double* prow,
                  const PivotState* st, int r_fixed, double* dj = nullptr, double* npart = nullptr) {
    if (C <= 0 || R <= 0) return;
    const int cpc = npart ? kNormCols : std::max(kColsInFlight, ctx->rank1_cols_per_cta);
    dim3 grid((unsigned)((R + 2 * kRank1Threads - 1) / (2 * kRank1Threads)), (unsigned)((C + cpc - 1) / cpc));
    const bool stream = (double)ld * C * 8.0 > (double)ctx->rank1_stream_min_mb * 1048576.0;
    if (npart) {
        if (stream) LAUNCH((k_rank1<true, true>), grid, kRank1Threads, E, ld, R, C, alpha, prow, st, r_fixed, cpc, dj, npart);
        else LAUNCH((k_rank1<false, true>), grid, kRank1Threads, E, ld, R, C, alpha, prow, st, r_fixed, cpc, dj, npart);
    } else {
        if (stream) LAUNCH((k_rank1<true, false>), grid, kRank1Threads, E, ld, R, C, alpha, prow, st, r_fixed, cpc, dj, npart);
        else LAUNCH((k_rank1<false, false>), grid, kRank1Threads, E, ld, R, C, alpha, prow, st, r_fixed, cpc, dj, npart);
    }
}

int gemv_grid(int ncols) { return std::max(1, std::min((ncols + 7) / 8, 148 * 32)); }

// One launch of the rank-k row reduction E -= U V (K3b).  Kernel choice (tuning key "flush_kernel"): 3 = bulk-copy /
// mbarrier ring, one CTA per SM (needs V rows padded to whole tiles, 16-byte aligned); 4 = the same ring with 16 consumer warps;
// 1 = two CTAs per SM without register prefetch (any V); 5 / 6 = version 4 with the warp tile pipelined in 2 / 4 parts; 7 / 8 = the
// producer-less 512-thread kernel (k_blk_flush6) with the schedule of version 4 / 5.
void launch_rankk(ellp_b200_ctx* ctx, double* E, int64_t ld, int R, int C, const double* U, const double* V, int64_t ldv, int cnt) {
    const int K4 = (cnt + 3) & ~3;
    int kern = ctx->flush_kernel;
    const bool automatic = (kern == 0);
    // auto, first part (profiles/r02_flush5_sweep.jsonl, r02_flush_lowk_sweep.jsonl, 32768^2 tableau): version 4r (12 consumer warps, 128
    // registers) is ahead of versions 3 / 4 from k = 24 on (k = 32: 22.6 vs 19.0 TFLOP/s, k = 56: 29.8 vs 27.4, k = 64: 30.8 vs 27.4 / 24.3);
    // below that the kernels are HBM-bound and level, version 3 stays
    if (automatic) kern = (cnt >= ctx->flush4_min_k) ? 9 : 3;
    const bool base_ok = (ldv % 2 == 0) && ((reinterpret_cast<uintptr_t>(V) & 15) == 0);
    const bool wide = kern >= 4;  // 128-column steps (4: 16 consumer warps; 5 / 6: tile pipelined in 2 / 4 parts; 7 / 8: no producer warp; 9: 12 warps)
    if (wide && !(base_ok && (int64_t)((C + kFlush4Cols - 1) / kFlush4Cols) * kFlush4Cols <= ldv)) kern = 3;
    if (kern == 3 && !(base_ok && (int64_t)((C + kFlushCols - 1) / kFlushCols) * kFlushCols <= ldv)) kern = 1;
    if (kern == 2) kern = 1;
    const int cols_per_step = kern >= 4 ? kFlush4Cols : kFlushCols;
    const int steps_total = (C + cols_per_step - 1) / cols_per_step;
    const size_t smem = kern >= 4 ? blk_flush4_smem_bytes(K4) : (kern == 3 ? blk_flush3_smem_bytes(K4) : blk_flush_smem_bytes(K4));
    // column steps per CTA.  One CTA per SM: keep enough waves of CTAs that the last partial wave stays small (narrow shards) -- 6 waves
    // of 128-row CTAs (profiles/r02_flush_shard_sweep.jsonl), 4 waves of the 96-row CTAs of version 4r (32768 x 4096, k = 64: 16 steps per
    // CTA = 684 CTAs 28.0 TFLOP/s, 8 steps 26.0, 32 steps 24.5)
    auto plan = [&](int kn) {
        int cs = std::max(1, std::min(kn == 1 ? ctx->flush_col_steps : ctx->flush2_col_steps, steps_total));
        if (kn != 1) {
            const int rows = (kn == 9) ? 96 : kFlushRows;
            const int waves = (kn == 9) ? std::min(ctx->flush_waves, 4) : ctx->flush_waves;
            const int64_t row_blocks = (R + rows - 1) / rows;
            while (cs > 4 && row_blocks * ((steps_total + cs - 1) / cs) < (int64_t)waves * 148) cs >>= 1;
        }
        return cs;
    };
    int col_steps = plan(kern);
    // auto, second part: where a CTA is short (<= 8 column steps: small tableaus such as 4096 x 8192, where it leads at every k -- k = 48:
    // 23.1 vs 20.2 (4r) vs 17.6 (3) TFLOP/s) the producer-less kernel with the tile pipelined in two halves wins; its 203 KB of shared
    // memory at k > 56 leave too little L1, so version 4r keeps those
    if (automatic && kern == 9 && col_steps <= 8 && cnt <= 56) { kern = 8; col_steps = plan(8); }
    dim3 grid((unsigned)((R + kFlushRows - 1) / kFlushRows), (unsigned)((steps_total + col_steps - 1) / col_steps));
    const bool stream = (double)R * C * 8.0 > (double)ctx->rank1_stream_min_mb * 1048576.0;
    ctx->last_flush_kernel = kern;
    if (kern == 9) {  // version 4r: 3 x 4 consumer warps, CTA tile 96 rows x 128 columns
        const int mode = ctx->flush_ld >= 0 ? (ctx->flush_ld ? 1 : 0) : (stream ? 1 : 0);
        int stages = blk_flush4_stages(K4);
        if (ctx->flush_stages > 0 && blk_flush4r_smem_bytes<3>(K4, ctx->flush_stages) <= (size_t)227 * 1024) stages = std::min(3, ctx->flush_stages);
        dim3 grid9((unsigned)((R + 95) / 96), grid.y);
        if (mode) LAUNCH_SMEM((k_blk_flush4r<1, 3>), grid9, flush4r_threads<3>(), blk_flush4r_smem_bytes<3>(K4, stages), E, ld, R, C, U, V, ldv, cnt, col_steps, stages);
        else LAUNCH_SMEM((k_blk_flush4r<0, 3>), grid9, flush4r_threads<3>(), blk_flush4r_smem_bytes<3>(K4, stages), E, ld, R, C, U, V, ldv, cnt, col_steps, stages);
        return;
    }
    if (kern >= 4) {
        // tile access mode (tuning key "flush_ld"): -1 = auto (evict-first when the matrix exceeds L2, default caching otherwise)
        const int mode = ctx->flush_ld >= 0 ? ctx->flush_ld : (stream ? 1 : 0);
        int stages = blk_flush4_stages(K4);
        if (ctx->flush_stages > 0 && blk_flush4_smem_bytes_st(K4, ctx->flush_stages) <= (size_t)227 * 1024) stages = std::min(3, ctx->flush_stages);
        const size_t smem_w = blk_flush4_smem_bytes_st(K4, stages);
#define ELLP_FLUSH_WIDE(KERNEL)                                                                                              \
        switch (mode) {                                                                                                      \
            case 0: LAUNCH_SMEM((KERNEL(0)), grid, kFlush4Threads, smem_w, E, ld, R, C, U, V, ldv, cnt, col_steps, stages); break; \
            case 2: LAUNCH_SMEM((KERNEL(2)), grid, kFlush4Threads, smem_w, E, ld, R, C, U, V, ldv, cnt, col_steps, stages); break; \
            case 3: LAUNCH_SMEM((KERNEL(3)), grid, kFlush4Threads, smem_w, E, ld, R, C, U, V, ldv, cnt, col_steps, stages); break; \
            default: LAUNCH_SMEM((KERNEL(1)), grid, kFlush4Threads, smem_w, E, ld, R, C, U, V, ldv, cnt, col_steps, stages); break; \
        }
#define ELLP_K4(M) k_blk_flush4<M>
#define ELLP_K5(M) k_blk_flush5<M, 2>
#define ELLP_K6(M) k_blk_flush5<M, 4>
        if (kern == 7 || kern == 8) {
            const int md = mode ? 1 : 0;
            if (kern == 7) {
                if (md) LAUNCH_SMEM((k_blk_flush6<1, 1>), grid, kFlush6Threads, smem_w, E, ld, R, C, U, V, ldv, cnt, col_steps, stages);
                else LAUNCH_SMEM((k_blk_flush6<0, 1>), grid, kFlush6Threads, smem_w, E, ld, R, C, U, V, ldv, cnt, col_steps, stages);
            } else {
                if (md) LAUNCH_SMEM((k_blk_flush6<1, 2>), grid, kFlush6Threads, smem_w, E, ld, R, C, U, V, ldv, cnt, col_steps, stages);
                else LAUNCH_SMEM((k_blk_flush6<0, 2>), grid, kFlush6Threads, smem_w, E, ld, R, C, U, V, ldv, cnt, col_steps, stages);
            }
        } else if (kern == 4) { ELLP_FLUSH_WIDE(ELLP_K4) }
        else if (kern == 5) { ELLP_FLUSH_WIDE(ELLP_K5) }
        else { ELLP_FLUSH_WIDE(ELLP_K6) }
#undef ELLP_K4
#undef ELLP_K5
#undef ELLP_K6
#undef ELLP_FLUSH_WIDE
    } else if (kern == 3) {
        if (stream) LAUNCH_SMEM(k_blk_flush3<true>, grid, kFlush3Threads, smem, E, ld, R, C, U, V, ldv, cnt, col_steps);
        else LAUNCH_SMEM(k_blk_flush3<false>, grid, kFlush3Threads, smem, E, ld, R, C, U, V, ldv, cnt, col_steps);
    } else {
        if (stream) LAUNCH_SMEM(k_blk_flush<true>, grid, 256, smem, E, ld, R, C, U, V, ldv, cnt, col_steps);
        else LAUNCH_SMEM(k_blk_flush<false>, grid, 256, smem, E, ld, R, C, U, V, ldv, cnt, col_steps);
    }
}

int flush_attrs(ellp_b200_ctx* ctx) {
    if (ctx->flush_attrs_set) return ELLP_OK;
    const int smem1 = (int)std::max(blk_flush_smem_bytes(48), blk_flush_smem_bytes(kBlkMax));
    CUDA_TRY(cudaFuncSetAttribute(k_blk_flush<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    CUDA_TRY(cudaFuncSetAttribute(k_blk_flush<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    CUDA_TRY(cudaFuncSetAttribute(k_blk_flush3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)blk_flush3_smem_bytes(kBlkMax)));
    CUDA_TRY(cudaFuncSetAttribute(k_blk_flush3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)blk_flush3_smem_bytes(kBlkMax)));
    const int smem4 = 227 * 1024;  // versions 4 / 5: the ring depth is a tuning parameter, allow the whole carve-out
#define ELLP_ATTR(K) CUDA_TRY(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4))
    ELLP_ATTR(k_blk_flush4<0>); ELLP_ATTR(k_blk_flush4<1>); ELLP_ATTR(k_blk_flush4<2>); ELLP_ATTR(k_blk_flush4<3>);
    ELLP_ATTR((k_blk_flush5<0, 2>)); ELLP_ATTR((k_blk_flush5<1, 2>)); ELLP_ATTR((k_blk_flush5<2, 2>)); ELLP_ATTR((k_blk_flush5<3, 2>));
    ELLP_ATTR((k_blk_flush5<0, 4>)); ELLP_ATTR((k_blk_flush5<1, 4>)); ELLP_ATTR((k_blk_flush5<2, 4>)); ELLP_ATTR((k_blk_flush5<3, 4>));
    ELLP_ATTR((k_blk_flush6<0, 1>)); ELLP_ATTR((k_blk_flush6<1, 1>)); ELLP_ATTR((k_blk_flush6<0, 2>)); ELLP_ATTR((k_blk_flush6<1, 2>));
    ELLP_ATTR((k_blk_flush4r<0, 3>)); ELLP_ATTR((k_blk_flush4r<1, 3>));
#undef ELLP_ATTR
    ctx->flush_attrs_set = true;
    return ELLP_OK;
}

// panel factorisation of the blocked LU: cooperative multi-CTA kernel while a CTA's slice of the panel fits shared
// memory, the single-CTA kernel otherwise
int launch_lu_panel(ellp_b200_ctx* ctx, const DevLP& lp, double* G, int64_t ld, int m, int k0, int nb) {
    if (ctx->lu_panel_grid == 0) {
        int sms = 0, coop = 0, per_sm = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
        CUDA_TRY(cudaFuncSetAttribute(k_lu_panel_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, kLuPanelSmemMax));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lu_panel_coop, 256, kLuPanelSmemMax));
        ctx->lu_panel_grid = (coop && per_sm > 0) ? std::min(sms, 159) : -1;
    }
    const int rows = m - k0;
    int nctas = ctx->lu_panel_grid > 0 ? std::max(1, std::min(ctx->lu_panel_grid, rows / 64)) : 0;
    int rpc = nctas > 0 ? (rows + nctas - 1) / nctas : 0;
    const size_t smem = (size_t)rpc * (kPanel + 1) * sizeof(double);
    if (nctas == 0 || smem > (size_t)kLuPanelSmemMax || ctx->refactor_panel == 1) {
        LAUNCH(k_lu_panel, 1, 1024, G, ld, m, k0, nb, lp.lu_piv, ctx->d_st);
        return ELLP_OK;
    }
    nctas = (rows + rpc - 1) / rpc;  // drop CTAs that would hold no row
    LuPanelPub* pub = reinterpret_cast<LuPanelPub*>(lp.coop);
    int32_t* piv = lp.lu_piv;
    PivotState* st = ctx->d_st;
    void* args[] = {(void*)&G, (void*)&ld, (void*)&m, (void*)&k0, (void*)&nb, (void*)&rpc, (void*)&piv, (void*)&pub, (void*)&st};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_lu_panel_coop, dim3(nctas), dim3(256), args, smem, ctx->stream));
    ctx->launches++;
    return ELLP_OK;
}

// B^-1 of the basis lp.Bv, from the resident constraint matrix lp.A, into lp.Binv = right half of G = [A_B | I] (see kernels.cuh /
// refactor.cuh).  The revised engine passes its own DevLP; the tableau engines pass a copy whose V / ldv / coop point to the LU
// scratch (ctx->rf_*).  Does not touch PivotState except err / gj_piv (the caller saves and restores it).
int refactor_binv(ellp_b200_ctx* ctx, const DevLP& lp) {
    const int m = lp.m;
    const bool use_lu = (ctx->refactor_mode == 2 || (ctx->refactor_mode == 0 && m >= 128));
    bool diagonal = false;
    if (ctx->refactor_mode == 0) {  // slack / artificial bases: B is diagonal, invert it directly
        int flags[2] = {0, 0};
        CUDA_TRY(cudaMemsetAsync(ctx->d_flag, 0, 2 * sizeof(int), ctx->stream));
        LAUNCH(k_check_diag_basis, m, 128, lp.A, lp.ld, m, lp.Bv, ctx->d_flag);
        CUDA_TRY(cudaMemcpyAsync(flags, ctx->d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        diagonal = (flags[0] == 0);
        if (diagonal && flags[1]) return set_err(ctx, ELLP_E_ELLP, dev_err_message(kErrSingular));
    }
    if (diagonal) {
        LAUNCH(k_diag_inverse, m, 256, lp);
    } else if (use_lu) {
        // K4: blocked partial-pivot LU of [A_B | I] with DMMA trailing updates, then the blocked back substitution (refactor.cuh)
        LAUNCH(k_gj_init, 2 * m, 256, lp);
        double* G = lp.G;
        const int64_t ld = lp.ld;
        const int ncols = 2 * m;
        if (int rc = flush_attrs(ctx)) return rc;
        for (int k0 = 0; k0 < m; k0 += kPanel) {
            const int nb = std::min(kPanel, m - k0), c0 = k0 + nb;
            if (int rc = launch_lu_panel(ctx, lp, G, ld, m, k0, nb)) return rc;
            LAUNCH(k_lu_swap_rows, (ncols - nb + 255) / 256, 256, G, ld, ncols, k0, nb, lp.lu_piv, ctx->d_st);
            if (ncols > c0) LAUNCH(k_lu_trsm_lower, (ncols - c0 + 127) / 128, 128, G, ld, ncols, k0, nb, c0, ctx->d_st, lp.V, lp.ldv);
            // G22 -= L21 U12 over rows [c0, ld) (padding rows are zero) and ALL remaining columns: fp64 tensor pipe (k_blk_flush3)
            if (m > c0) launch_rankk(ctx, G + (int64_t)c0 * ld + c0, ld, (int)(ld - c0), ncols - c0, G + (int64_t)k0 * ld + c0, lp.V, lp.ldv, nb);
        }
        for (int k0 = ((m - 1) / kPanel) * kPanel; k0 >= 0; k0 -= kPanel) {
            const int nb = std::min(kPanel, m - k0);
            LAUNCH(k_lu_trsm_upper, (m + 127) / 128, 128, G, ld, ncols, k0, nb, m, ctx->d_st, lp.V, lp.ldv);
            // X[0:k0, :] -= U[0:k0, k0:k0+nb] X[k0:k0+nb, :]
            if (k0 > 0) launch_rankk(ctx, G + (int64_t)m * ld, ld, k0, m, G + (int64_t)k0 * ld, lp.V, lp.ldv, nb);
        }
    } else if (ctx->small_path && m <= kSmallMaxM && gj_small_smem_bytes(lp.ld, m) <= (size_t)kGjSmallSmemMax) {
        // netlib-sized basis: the whole Gauss-Jordan inverse in one single-CTA launch (small.cuh)
        if (!ctx->small_attr_set) {
            CUDA_TRY(cudaFuncSetAttribute(k_gj_small, cudaFuncAttributeMaxDynamicSharedMemorySize, kGjSmallSmemMax));
            ctx->small_attr_set = true;
        }
        LAUNCH_SMEM(k_gj_small, 1, kSmallThreads, gj_small_smem_bytes(lp.ld, m), lp, ctx->d_st);
    } else {
        LAUNCH(k_gj_init, 2 * m, 256, lp);
        for (int k = 0; k < m; ++k) {
            LAUNCH(k_gj_pivot, 1, 1024, lp.G, lp.ld, m, (const int32_t*)nullptr, k, lp.dcol, ctx->d_st);
            const int cols = 2 * m - k;
            LAUNCH(k_gj_swap_gather, (cols + 255) / 256, 256, lp.G, lp.ld, k, 2 * m, k, lp.prow, ctx->d_st);
            launch_rank1(ctx, lp.G + (int64_t)k * lp.ld, lp.ld, m, cols, lp.dcol, lp.prow, ctx->d_st, 0);
        }
    }
    return ELLP_OK;
}

// Condensed tableau from a general basis: B^-1 by refactor_binv, then T = B^-1 A_N as m / kb rank-kb updates on the fp64 tensor
// pipe (the row-reduction kernel K3b with U = a column block of B^-1 and V = minus the matching row block of A_N).
// Replaces m sequential full-tableau Gauss-Jordan sweeps; leaves A intact and B^-1 behind for the dual's y.
int tableau_from_binv(ellp_b200_ctx* ctx) {
    DevLP& lp = ctx->lp;
    DevLP rl = lp;
    rl.V = ctx->rf_V;
    rl.ldv = ctx->rf_ldv;
    rl.coop = ctx->rf_coop;
    if (int rc = refactor_binv(ctx, rl)) return rc;
    if (int rc = flush_attrs(ctx)) return rc;
    const int m = lp.m, nT = lp.nT;
    // row blocks of A_N go through the tableau's own V (blk_kmax rows) when the blocked engine is active, else through the LU's block row
    double* Vb = lp.V ? lp.V : ctx->rf_V;
    const int64_t ldvb = lp.V ? lp.ldv : ctx->rf_ldv;
    const int kb = lp.V ? std::min(ctx->blk_kmax, kBlkMax) : kPanel;
    if (!lp.V && ldvb < nT) return set_err(ctx, ELLP_E_ARG, "tableau start from a general basis needs n - m <= 2 m without block_k (use block_k > 1)");
    CUDA_TRY(cudaMemsetAsync(lp.T, 0, sizeof(double) * (size_t)lp.ld * (size_t)nT, ctx->stream));
    for (int k0 = 0; k0 < m; k0 += kb) {
        const int nb = std::min(kb, m - k0);
        dim3 gg((unsigned)((nT + 255) / 256), (unsigned)nb);
        LAUNCH(k_gather_rows_neg, gg, 256, lp.A, lp.ld, lp.Nv + lp.pos_lo, nT, k0, Vb, ldvb);
        launch_rankk(ctx, lp.T, lp.ld, (int)lp.ld, nT, lp.Binv + (int64_t)k0 * lp.ld, Vb, ldvb, nb);
    }
    if (lp.V) CUDA_TRY(cudaMemsetAsync(lp.V, 0, sizeof(double) * (size_t)lp.ldv * (size_t)ctx->blk_kmax, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(lp.Bv0, lp.Bv, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->tab_from_binv = true;
    return ELLP_OK;
}

// |A x - b|_inf of the resident point (tableau engines that kept A): one chunked GEMV over A, deterministic partial sums.
int primal_residual(ellp_b200_ctx* ctx, double* out) {
    DevLP& lp = ctx->lp;
    const int m = lp.m;
    unsigned long long* slot = reinterpret_cast<unsigned long long*>(ctx->d_flag + 2);  // d_flag holds 4 ints; [0], [1] belong to refactor()
    CUDA_TRY(cudaMemsetAsync(slot, 0, sizeof(unsigned long long), ctx->stream));
    dim3 gx((unsigned)((m + 255) / 256), (unsigned)kXChunks);
    LAUNCH(k_gemv_n_chunks, gx, 256, lp.A, lp.ld, m, lp.n, (const double*)lp.x, (const int32_t*)nullptr, lp.xpart);
    LAUNCH(k_residual_max, (m + 255) / 256, 256, (const double*)lp.xpart, lp.ld, m, lp.b, slot);
    unsigned long long bits = 0;
    CUDA_TRY(cudaMemcpyAsync(&bits, slot, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    double r;
    std::memcpy(&r, &bits, sizeof(r));
    *out = r;
    return ELLP_OK;
}

// Refactorisation.  Revised engine: B^-1 from the current basis on G = [A_B | I].  Tableau engine: T = B^-1 A_N rebuilt from the
// resident A (gathered when the basis columns are the identity, through B^-1 otherwise), then -- primal -- the reduced-cost row
// d = c - c_B^T T; the dual keeps its own d (dual :296-302 never recomputes it).
int refactor(ellp_b200_ctx* ctx, uint64_t* count) {
    DevLP& lp = ctx->lp;
    const int m = lp.m;
    // preserve the iteration's index-level state around the factorisation
    if (int rc = read_state(ctx)) return rc;
    PivotState saved = *ctx->h_st;
    if (!ctx->tableau) {
        if (int rc = refactor_binv(ctx, lp)) return rc;
    } else {
        // F = A (sharded in-place engine: the tableau itself); the condensed engine keeps only the nonbasic columns in T
        double* F = const_cast<double*>(lp.A);
        if (ctx->dj_live) LAUNCH(k_dual_tab_scatter_d, (lp.nN + 255) / 256, 256, lp, (const double*)lp.dj);  // mid-solve: d follows dj
        int mismatch = 0;
        if (!ctx->sharded) {  // a sharded tableau is only accepted with an identity starting basis (checked at upload)
            CUDA_TRY(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
            LAUNCH(k_check_identity_basis, m, 256, F, lp.ld, m, lp.Bv, ctx->d_flag);
            CUDA_TRY(cudaMemcpyAsync(&mismatch, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
        bool built = false;
        if (mismatch && lp.condensed && ctx->has_binv) {
            if (int rc = tableau_from_binv(ctx)) return rc;
            built = true;
        } else if (mismatch) {  // no room for B^-1 was reserved: m in-place Gauss-Jordan sweeps over the full tableau (destroys A)
            for (int k = 0; k < m; ++k) {
                LAUNCH(k_gj_pivot, 1, 1024, F, lp.ld, m, (const int32_t*)lp.Bv, k, lp.dcol, ctx->d_st);
                LAUNCH(k_gj_swap_gather, (lp.n + 255) / 256, 256, F, lp.ld, 0, lp.n, k, lp.prow, ctx->d_st);
                launch_rank1(ctx, F, lp.ld, m, lp.n, lp.dcol, lp.prow, ctx->d_st, 0);
            }
            ctx->a_resident = false;
        }
        LAUNCH(k_init_cB, (int)((lp.ld + 255) / 256), 256, lp);
        if (lp.condensed) {
            if (!built) {
                dim3 gg((unsigned)std::max<int64_t>(1, std::min<int64_t>(8, (lp.ld + 255) / 256)), (unsigned)lp.nN);
                LAUNCH(k_gather_cols, gg, 256, F, lp.ld, lp.Nv, lp.nN, lp.T);
                CUDA_TRY(cudaMemcpyAsync(lp.Bv0, lp.Bv, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToDevice, ctx->stream));
                LAUNCH(k_fill_const, (int)((lp.ld + 255) / 256), 256, lp.bscale, (int64_t)lp.ld, 1.0);
                ctx->tab_from_binv = false;
            }
            if (ctx->recompute_x && lp.xpart && !ctx->peer_mode) {
                // mid-solve rebuild: x_B = B^-1 b - T x_N removes the drift of the incrementally updated point
                static_assert(kXChunks == 32, "lp.xpart is carved for 32 chunks");
                dim3 gx((unsigned)((m + 255) / 256), (unsigned)kXChunks);
                if (built) LAUNCH(k_gemv_n_chunks, gx, 256, (const double*)lp.Binv, lp.ld, m, m, lp.b, (const int32_t*)nullptr, lp.xpart);
                LAUNCH(k_gemv_n_chunks, gx, 256, (const double*)lp.T, lp.ld, m, lp.nT, (const double*)lp.x, (const int32_t*)lp.Nv, lp.xpart + (int64_t)kXChunks * lp.ld);
                LAUNCH(k_xB_finish, (m + 255) / 256, 256, lp, (const double*)lp.xpart, (const double*)(lp.xpart + (int64_t)kXChunks * lp.ld), built ? 0 : 1);
            }
            if (ctx->solver == ELLP_DUAL) {
                LAUNCH(k_dual_tab_init, (lp.nT + 255) / 256, 256, lp);
                ctx->dj_live = true;
            } else {
                LAUNCH(k_gemv_t<EPI_PLAIN>, gemv_grid(lp.nN), 256, lp.T, lp.ld, (const int32_t*)nullptr, lp.nN, lp.cB, lp.dj, (const double*)nullptr,
                       (const uint8_t*)nullptr, (double*)nullptr, (PivotState*)nullptr, 0);
                LAUNCH(k_redcost_pos, (lp.nN + 255) / 256, 256, lp.c, lp.Nv, lp.nN, lp.dj);
            }
        } else {
            LAUNCH(k_gemv_t<EPI_REDCOST>, gemv_grid(lp.n), 256, lp.T, lp.ld, (const int32_t*)nullptr, lp.n, lp.cB, lp.dj, lp.c + lp.col_lo,
                   (const uint8_t*)nullptr, (double*)nullptr, (PivotState*)nullptr, 0);
        }
    }
    if (int rc = read_state(ctx)) return rc;
    const int err = ctx->h_st->err;
    *ctx->h_st = saved;
    ctx->h_st->err = 0;
    if (int rc = write_state(ctx)) return rc;
    if (err == kErrSingular) return set_err(ctx, ELLP_E_ELLP, dev_err_message(err));
    ctx->binv_valid = true;
    ctx->pivots_since_refactor = 0;
    if (ctx->tableau && ctx->devex_live) {  // new reference framework
        LAUNCH(k_fill_const, (int)((lp.ld + 255) / 256), 256, lp.w, (int64_t)lp.ld, 1.0);
        if (lp.wN) LAUNCH(k_fill_const, (lp.nT + 255) / 256, 256, lp.wN, (int64_t)lp.nT, 1.0);
    }
    if (count) ++*count;
    return ELLP_OK;
}

// tableau engine, primal: 5 launches per pivot, the rank-1 update of T is >99 % of the bytes
// blocked engine: T -= U V over the slots filled since the last flush (k_blk_flush, fp64 tensor pipe), timed like K3
void launch_flush(ellp_b200_ctx* ctx, bool profile, size_t* ev_used) {
    DevLP& lp = ctx->lp;
    const int cnt = ctx->blk_fill;
    ctx->blk_fill = 0;
    if (cnt <= 0) return;
    if (profile && ev_used && *ev_used + 2 <= ctx->ev.size()) cudaEventRecord(ctx->ev[(*ev_used)++], ctx->stream);
    launch_rankk(ctx, lp.T, lp.ld, (int)lp.ld, lp.nT, lp.U, lp.V, lp.ldv, cnt);
    if (profile && ev_used && (*ev_used & 1)) cudaEventRecord(ctx->ev[(*ev_used)++], ctx->stream);
}

// ---- peer-memory sharded engine (peer.cuh) -------------------------------------------------------------------------
// Exchange buffer of one rank: [ kMboxWords mailbox words | 2 * cap column words | 2 tickets ], 16 bytes per word.
// Collective: every rank calls it with the same `rows`.  The cudaIpc handles travel through one ncclAllGather.
int peer_setup(ellp_b200_ctx* ctx, int64_t rows) {
    if (ctx->peer_own && ctx->peer_cap >= rows) return ELLP_OK;
    const int R = ctx->nranks;
    if (R > kMaxPeers) return set_err(ctx, ELLP_E_ARG, "the peer-memory engine supports at most 8 ranks");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    peer_release(ctx);
    const int64_t cap = (int64_t)align_up((size_t)std::max<int64_t>(rows, 1024), 1024);
    const size_t words = (size_t)kMboxWords + 2 * (size_t)cap;
    const size_t bytes = words * sizeof(uint4) + 256;
    CUDA_TRY(cudaMalloc(&ctx->peer_own, bytes));
    CUDA_TRY(cudaMemsetAsync(ctx->peer_own, 0, bytes, ctx->stream));
    std::vector<cudaIpcMemHandle_t> handles((size_t)R);
    if (R > 1) {
        cudaIpcMemHandle_t mine;
        CUDA_TRY(cudaIpcGetMemHandle(&mine, ctx->peer_own));
        unsigned char* d_h = nullptr;
        CUDA_TRY(cudaMalloc(&d_h, sizeof(cudaIpcMemHandle_t) * (size_t)(R + 1)));
        CUDA_TRY(cudaMemcpyAsync(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
        // the all-gather also orders every rank's memset before any rank's first remote store
        NCCL_TRY(nccl::api.AllGather(d_h, d_h + sizeof(mine), sizeof(mine), nccl::kUint8, ctx->nccl_comm, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(handles.data(), d_h + sizeof(mine), sizeof(mine) * (size_t)R, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaFree(d_h));
    }
    PeerLinks pl{};
    pl.rank = ctx->rank;
    pl.nranks = R;
    pl.col_cap = cap;
    for (int r = 0; r < R; ++r) {
        void* base = ctx->peer_own;
        if (r != ctx->rank) {
            CUDA_TRY(cudaIpcOpenMemHandle(&ctx->peer_map[r], handles[(size_t)r], cudaIpcMemLazyEnablePeerAccess));
            base = ctx->peer_map[r];
        }
        pl.mbox[r] = reinterpret_cast<uint4*>(base);
        pl.col[r] = reinterpret_cast<uint4*>(base) + kMboxWords;
    }
    pl.ticket = reinterpret_cast<unsigned int*>(reinterpret_cast<uint4*>(ctx->peer_own) + words);
    ctx->pl = pl;
    ctx->peer_cap = cap;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ELLP_OK;
}

// layout of the peer engine: replicated vectors + the local slice [pos_lo, pos_lo + nT) of the condensed tableau
void carve_peer(Arena& a, DevLP& lp, int64_t trace_cap, int blk_kmax) {
    const size_t ld = (size_t)lp.ld, m = (size_t)lp.m, ng = (size_t)lp.n_glob, nN = (size_t)lp.nN, nT = (size_t)lp.nT;
    lp.c = a.take<double>(ng);
    lp.b = a.take<double>(m);
    lp.lb = a.take<double>(ng);
    lp.ub = a.take<double>(ng);
    lp.kind = a.take<uint8_t>(ng);
    lp.x = a.take<double>(ng);
    lp.Bv = a.take<int32_t>(m);
    lp.Nv = a.take<int32_t>(nN);
    lp.Ns = a.take<uint8_t>(nN);
    lp.y = a.take<double>(ld);
    lp.d = a.take<double>(ng);
    lp.G = lp.Binv = nullptr;
    lp.condensed = 1;
    lp.T = a.take<double>(ld * nT);
    lp.A = lp.T;  // the starting basis is the identity: T = A_N; the constraint matrix is not kept separately
    lp.dj = a.take<double>(nT + 8);
    lp.ldv = (int64_t)align_up(nT, 128);
    lp.coop = a.take<double>(6 * 1024);
    lp.U = a.take<double>(ld * (size_t)blk_kmax);
    lp.V = a.take<double>((size_t)lp.ldv * (size_t)blk_kmax);
    lp.cB = a.take<double>(ld);
    lp.u = a.take<double>(ld);
    lp.rN = a.take<double>(nT + 8);
    lp.key = a.take<double>(nT + 8);
    lp.dcol = a.take<double>(ld);
    lp.rho = a.take<double>(ld);
    lp.prow = a.take<double>(nT + 8);
    lp.part = a.take<double>(8);
    lp.lam = a.take<double>(m);
    lp.lu_piv = a.take<int32_t>(m);
    lp.w = a.take<double>(ld);
    lp.npart = nullptr;
    lp.Bv0 = a.take<int32_t>(m);
    lp.bscale = a.take<double>(ld);
    lp.dpos = a.take<double>(nN);
    lp.wN = a.take<double>(nT + 8);
    lp.trace = trace_cap > 0 ? a.take<ellp_trace_rec>((size_t)trace_cap) : nullptr;
    lp.colstat = nullptr;
    lp.xchg = nullptr;
}

int peer_prepare(ellp_b200_ctx* ctx, int32_t m, int32_t n_glob, const ellp_opts* o, DevLP* out, int solver = ELLP_PRIMAL) {
    if (!ctx->nccl_comm) return set_err(ctx, ELLP_E_ARG, "call ellp_b200_comm_init first");
    const int nN = n_glob - m;
    if (nN <= 0 || nN % ctx->nranks != 0) return set_err(ctx, ELLP_E_ARG, "peer sharding needs the number of nonbasic columns (n - m) divisible by the number of ranks");
    if (m % 4 != 0) return set_err(ctx, ELLP_E_ARG, "column sharding needs m % 4 == 0");
    const int blk = blk_slots(o, true, solver);
    if (blk <= 1) return set_err(ctx, ELLP_E_ARG, "the peer-memory engine needs ellp_opts::block_k > 1");
    DevLP lp{};
    lp.m = m;
    lp.n_glob = n_glob;
    lp.n = n_glob;
    lp.col_lo = 0;
    lp.nN = nN;
    lp.nT = nN / ctx->nranks;
    lp.pos_lo = ctx->rank * lp.nT;
    lp.ld = m;
    const int64_t tcap = o->trace ? o->trace_cap : 0;
    Arena probe;
    carve_peer(probe, lp, tcap, blk);
    if (int rc = ensure_arena(ctx, probe.off + 256)) return rc;
    Arena a;
    a.base = ctx->arena;
    carve_peer(a, lp, tcap, blk);
    if (int rc = peer_setup(ctx, lp.ld)) return rc;
    CUDA_TRY(cudaMemsetAsync(lp.coop, 0, sizeof(double) * 6 * 1024, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(lp.cB, 0, (size_t)((char*)lp.lam - (char*)lp.cB), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(lp.y, 0, sizeof(double) * lp.ld, ctx->stream));
    ctx->blk_kmax = blk;
    ctx->blk_fill = 0;
    ctx->KS = 1;
    ctx->kc = 64;
    ctx->trace_cap = tcap;
    ctx->solver = solver;
    ctx->tableau = true;
    ctx->sharded = true;
    ctx->peer_mode = true;
    ctx->has_binv = false;
    ctx->tab_from_binv = false;
    ctx->dj_live = false;
    ctx->devex_live = false;
    ctx->b_inf = -1.;
    ctx->a_resident = true;  // lp.A aliases the local slice of T (= A_N while the tableau is fresh): download_std_form reads it
    ctx->dual_obj0 = 0.;
    *out = lp;
    return ELLP_OK;
}

// reduced-cost row of the local positions of the fresh tableau (identity basis: T = A_N)
// (primal) / dj = d of the local positions (dual).  diag_scaled: lp.bscale holds the diagonal of the starting basis and T still
// holds A_N: divide its rows (T = D^-1 A_N).
int peer_finish_init(ellp_b200_ctx* ctx, bool diag_scaled = false) {
    DevLP& lp = ctx->lp;
    LAUNCH(k_init_cB, (int)((lp.ld + 255) / 256), 256, lp);
    CUDA_TRY(cudaMemcpyAsync(lp.Bv0, lp.Bv, sizeof(int32_t) * (size_t)lp.m, cudaMemcpyDeviceToDevice, ctx->stream));
    if (diag_scaled) {
        dim3 gs((unsigned)((lp.ld / 2 + 255) / 256), (unsigned)std::min(lp.nT, 4096));
        LAUNCH(k_scale_rows_inv, gs, 256, lp.T, lp.ld, lp.m, lp.nT, (const double*)lp.bscale);
    }
    if (ctx->solver == ELLP_DUAL) {
        LAUNCH(k_dual_tab_init, (lp.nT + 255) / 256, 256, lp);
        ctx->dj_live = true;
    } else {
        LAUNCH(k_gemv_t<EPI_PLAIN>, gemv_grid(lp.nT), 256, lp.T, lp.ld, (const int32_t*)nullptr, lp.nT, lp.cB, lp.dj, (const double*)nullptr,
               (const uint8_t*)nullptr, (double*)nullptr, (PivotState*)nullptr, 0);
        LAUNCH(k_redcost_pos_local, (lp.nT + 255) / 256, 256, lp.c, lp.Nv, lp.pos_lo, lp.nT, lp.dj);
    }
    ctx->resident = true;
    ctx->lp_generation++;
    ctx->binv_valid = true;
    ctx->pivots_since_refactor = 0;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    return ELLP_OK;
}

// `npiv` pivots (slots blk_fill .. blk_fill + npiv - 1) in one cooperative launch of the peer-memory pivot kernel; every
// rank issues the same sequence of launches.  self_only: single-GPU blocked engine (the exchange buffers are this GPU's own).
int launch_coop_pivots_peer(ellp_b200_ctx* ctx, const ellp_opts* o, int npiv, bool self_only) {
    DevLP& lp = ctx->lp;
    const bool dual = (ctx->solver == ELLP_DUAL);  // dual_blocked.cuh: same layout, same exchange buffers, no tie folds (no dynamic smem)
    const bool devex = o->pricing == ELLP_PRICE_DEVEX;
    const void* fn = dual ? (devex ? (const void*)k_blk_dual_pivots_fused<true> : (const void*)k_blk_dual_pivots_fused<false>)
                          : (devex ? (const void*)k_blk_pivots_fused<true> : (const void*)k_blk_pivots_fused<false>);
    const size_t smem = dual ? 0 : (size_t)kScanSmemBytes;
    int& grid_cap = ctx->coop_grid[(dual ? 2 : 0) + (devex ? 1 : 0)];
    int& threads_cached = ctx->coop_threads_cached[(dual ? 2 : 0) + (devex ? 1 : 0)];
    // the fused kernel runs with small blocks: its phases are latency-bound and every block-wide reduction / barrier costs
    // issue slots per resident warp (measured: 256 threads per block beat 1024)
    const int threads = std::max(64, std::min(kFusedMaxThreads, ctx->coop_threads & ~31));
    if (grid_cap == 0 || threads_cached != threads) {
        if (smem) CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int sms = 0, per_sm = 0, coop = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
        grid_cap = (coop && per_sm > 0) ? std::min(kLLMaxBlocks, sms * std::min(per_sm, std::max(1, ctx->coop_ctas_per_sm))) : -1;
        threads_cached = threads;
    }
    if (grid_cap < 0) return set_err(ctx, ELLP_E_CUDA, "cooperative launch unavailable");
    const int64_t work = std::max<int64_t>(std::max<int64_t>(lp.ld, lp.ldv), lp.nT);
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(grid_cap, (work + threads - 1) / threads));
    int tie = o->tie_rule, slot0 = ctx->blk_fill;
    uint32_t seq0 = ctx->xseq;
    PivotState* st = ctx->d_st;
    PeerLinks pl = ctx->pl;
    pl.tlog = ctx->tlog ? ctx->tlog - (int64_t)ctx->tlog_seq0 * kTlogStamps : nullptr;  // the kernel indexes the log by (seq - 1)
    pl.tlog_cap = ctx->tlog ? (int32_t)(ctx->tlog_seq0 + (uint32_t)ctx->tlog_cap) : 0;
    pl.owner_only = ctx->owner_ratio > 0 ? 1 : 0;  // every rank launches with the same value; measured at 8 GPUs: the redundant test is 2 % faster
    if (self_only) {
        pl.mbox[0] = ctx->pl.mbox[ctx->pl.rank];
        pl.col[0] = ctx->pl.col[ctx->pl.rank];
        pl.rank = 0;
        pl.nranks = 1;
    }
    if (dual) {
        void* args[] = {(void*)&lp, (void*)&pl, (void*)&slot0, (void*)&npiv, (void*)&seq0, (void*)&st};
        CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, smem, ctx->stream));
    } else {
        void* args[] = {(void*)&lp, (void*)&pl, (void*)&tie, (void*)&slot0, (void*)&npiv, (void*)&seq0, (void*)&st};
        CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, smem, ctx->stream));
    }
    ctx->launches++;
    ctx->blk_fill += npiv;
    ctx->xseq += (uint32_t)npiv;
    return ELLP_OK;
}

void launch_tableau_primal_iteration(ellp_b200_ctx* ctx, const ellp_opts* o, bool profile, size_t* ev_used, int blk) {
    DevLP& lp = ctx->lp;
    PivotState* st = ctx->d_st;
    const int m = lp.m, nN = lp.nN, nT = lp.nT;
    LAUNCH(k_price_tab, (nN + 255) / 256, 256, lp.dj, lp.Nv, lp.Ns, nN, lp.rN, lp.key, st, lp.condensed);  // primal :189, :253-270
    LAUNCH_SMEM(k_select_primal, 1, kScanThreads, kScanSmemBytes, lp.key, lp.rN, lp.Nv, lp.Ns, nN, o->tie_rule, st);           // :271-292
    LAUNCH(k_ratio_prep, (m + 255) / 256, 256, lp, 0, blk > 0 ? ctx->blk_fill : 0, st);         // :295-367
    if (o->ratio == ELLP_RATIO_HARRIS) LAUNCH(k_ratio_pick_harris, 1, 1024, lp, 1e-9, st);      // Harris two-pass test (opt-in)
    else LAUNCH_SMEM(k_ratio_pick, 1, kScanThreads, kScanSmemBytes, lp, o->tie_rule, st);        // :379-434, :205-232
    if (blk > 0) {  // deferred row reduction: new (U, V) slot now, T -= U V every blk pivots
        LAUNCH(k_blk_row, (int)((std::max<int64_t>(lp.ld, lp.ldv) + 255) / 256), 256, lp, ctx->blk_fill, st);
        if (++ctx->blk_fill >= blk) launch_flush(ctx, profile, ev_used);
        return;
    }
    LAUNCH(k_step_gather_cond, (int)((std::max<int64_t>(lp.ld, nT) + 255) / 256), 256, lp, st);  // :408-417 + pivot row
    if (profile && *ev_used + 2 <= ctx->ev.size()) cudaEventRecord(ctx->ev[(*ev_used)++], ctx->stream);
    launch_rank1(ctx, lp.T, lp.ld, m, nT, lp.dcol, lp.prow, st, 0, lp.dj);
    if (profile && (*ev_used & 1)) cudaEventRecord(ctx->ev[(*ev_used)++], ctx->stream);
}

// column-sharded tableau: 7 kernels + 3 NCCL collectives per pivot, no host involvement
int launch_sharded_iteration(ellp_b200_ctx* ctx, const ellp_opts* o, bool profile, size_t* ev_used, int blk) {
    DevLP& lp = ctx->lp;
    PivotState* st = ctx->d_st;
    const int m = lp.m, n = lp.n, G = ctx->nranks;
    LAUNCH(k_price_shard, (n + 255) / 256, 256, lp, st);
    LAUNCH(k_shard_local_max, 1, 1024, lp, st);
    NCCL_TRY(nccl::api.AllGather(lp.xchg + 0, lp.xchg + 16, 1, nccl::kFloat64, ctx->nccl_comm, ctx->stream));
    LAUNCH(k_shard_pick, 1, 1024, lp, G, st);
    NCCL_TRY(nccl::api.AllGather(lp.xchg + 8, lp.xchg + 32, 3, nccl::kFloat64, ctx->nccl_comm, ctx->stream));
    LAUNCH(k_shard_stage_column, (int)((lp.ld + 255) / 256), 256, lp, G, blk > 0 ? ctx->blk_fill : 0, st, ctx->sendcol);
    NCCL_TRY(nccl::api.AllReduce(ctx->sendcol, lp.dcol, (size_t)lp.ld, nccl::kFloat64, nccl::kSum, ctx->nccl_comm, ctx->stream));
    LAUNCH(k_ratio_prep, (m + 255) / 256, 256, lp, -1, 0, st);
    LAUNCH_SMEM(k_ratio_pick, 1, kScanThreads, kScanSmemBytes, lp, o->tie_rule, st);
    if (blk > 0) {
        LAUNCH(k_blk_row, (int)((std::max<int64_t>(lp.ld, lp.ldv) + 255) / 256), 256, lp, ctx->blk_fill, st);
        if (++ctx->blk_fill >= blk) launch_flush(ctx, profile, ev_used);
        return ELLP_OK;
    }
    LAUNCH(k_step_gather, (std::max(m, n) + 255) / 256, 256, lp, lp.T, n, st);
    if (profile && *ev_used + 2 <= ctx->ev.size()) cudaEventRecord(ctx->ev[(*ev_used)++], ctx->stream);
    launch_rank1(ctx, lp.T, lp.ld, m, n, lp.dcol, lp.prow, st, 0, lp.dj);
    if (profile && (*ev_used & 1)) cudaEventRecord(ctx->ev[(*ev_used)++], ctx->stream);
    return ELLP_OK;
}

void launch_primal_iteration(ellp_b200_ctx* ctx, const ellp_opts* o, bool profile, size_t* ev_used) {
    DevLP& lp = ctx->lp;
    PivotState* st = ctx->d_st;
    const int m = lp.m, nN = lp.nN;
    // BTRAN u = B^-T c_B  (primal :184-187)
    LAUNCH(k_gemv_t<EPI_PLAIN>, gemv_grid(m), 256, lp.Binv, lp.ld, (const int32_t*)nullptr, m, lp.cB, lp.u,
           (const double*)nullptr, (const uint8_t*)nullptr, (double*)nullptr, st, 1);
    // pricing r = c_N - A_N^T u + Dantzig keys  (primal :189, :253-270)
    LAUNCH(k_gemv_t<EPI_PRIMAL_PRICE>, gemv_grid(nN), 256, lp.A, lp.ld, lp.Nv, nN, lp.u, lp.rN, lp.c, lp.Ns, lp.key, st, 0);
    LAUNCH_SMEM(k_select_primal, 1, kScanThreads, kScanSmemBytes, lp.key, lp.rN, lp.Nv, lp.Ns, nN, o->tie_rule, st);
    // FTRAN d = B^-1 a_q  (primal :295)
    dim3 fg((unsigned)((lp.ld + 255) / 256), (unsigned)ctx->KS);
    LAUNCH(k_ftran_partial, fg, 128, lp.Binv, lp.ld, m, lp.A, st, lp.part, ctx->kc);
    LAUNCH(k_ratio_prep, (m + 255) / 256, 256, lp, ctx->KS, 0, st);
    if (o->ratio == ELLP_RATIO_HARRIS) LAUNCH(k_ratio_pick_harris, 1, 1024, lp, 1e-9, st);
    else LAUNCH_SMEM(k_ratio_pick, 1, kScanThreads, kScanSmemBytes, lp, o->tie_rule, st);
    LAUNCH(k_step_gather, (m + 255) / 256, 256, lp, lp.Binv, m, st);
    if (profile && *ev_used + 2 <= ctx->ev.size()) cudaEventRecord(ctx->ev[(*ev_used)++], ctx->stream);
    launch_rank1(ctx, lp.Binv, lp.ld, m, m, lp.dcol, lp.prow, st, 0);
    if (profile && (*ev_used & 1)) cudaEventRecord(ctx->ev[(*ev_used)++], ctx->stream);
}

void launch_dual_iteration(ellp_b200_ctx* ctx, const ellp_opts* o, bool profile, size_t* ev_used) {
    DevLP& lp = ctx->lp;
    PivotState* st = ctx->d_st;
    const int m = lp.m, nN = lp.nN;
    const bool dse = o->pricing == ELLP_PRICE_STEEPEST_EDGE;
    const int sel_m = std::max(1, std::min(kSelMaxBlocks, (m + 255) / 256)), sel_n = std::max(1, std::min(kSelMaxBlocks, (nN + 255) / 256));
    if (dse) LAUNCH(k_dual_leaving_dse, sel_m, 256, lp, st, ctx->d_sel);                      // steepest edge (no reference counterpart)
    else LAUNCH(k_dual_leaving, sel_m, 256, lp, st, ctx->d_sel);                              // dual :200-236
    LAUNCH(k_gather_row, (m + 255) / 256, 256, lp.Binv, lp.ld, m, st, lp.rho, 0);             // rho = e_r^T B^-1 (:248-253)
    LAUNCH(k_gemv_t<EPI_PLAIN>, gemv_grid(nN), 256, lp.A, lp.ld, lp.Nv, nN, lp.rho, lp.rN,  // alpha = A_N^T rho (:255)
           (const double*)nullptr, (const uint8_t*)nullptr, (double*)nullptr, st, 0);
    if (o->ratio == ELLP_RATIO_HARRIS) LAUNCH(k_select_dual_harris, 1, 1024, lp, 1e-9, st);
    else LAUNCH(k_select_dual, sel_n, 256, lp, st, ctx->d_sel);                               // :257-289
    dim3 fg((unsigned)((lp.ld + 255) / 256), (unsigned)ctx->KS);
    LAUNCH(k_ftran_partial, fg, 128, lp.Binv, lp.ld, m, lp.A, st, lp.part, ctx->kc);          // :294
    LAUNCH(k_dual_update_vec, (std::max(m, nN) + 255) / 256, 256, lp, ctx->KS, st);           // :296-316
    LAUNCH(k_dual_update_tail, 1, 32, lp, ctx->KS, st);                                       // :296, :302, :314-333
    if (profile && *ev_used + 2 <= ctx->ev.size()) cudaEventRecord(ctx->ev[(*ev_used)++], ctx->stream);
    launch_rank1(ctx, lp.Binv, lp.ld, m, m, lp.dcol, lp.prow, st, 0, nullptr, dse ? lp.npart : nullptr);
    if (profile && (*ev_used & 1)) cudaEventRecord(ctx->ev[(*ev_used)++], ctx->stream);
    if (dse) LAUNCH(k_sum_norms, (m + 255) / 256, 256, lp.npart, lp.ld, m, (m + kNormCols - 1) / kNormCols, lp.w, st, 1);
}

// exact dual steepest-edge weights from the current basis inverse (start of a run / after a refactorisation)
void launch_row_norms(ellp_b200_ctx* ctx) {
    DevLP& lp = ctx->lp;
    const int m = lp.m;
    dim3 grid((unsigned)((m + 2 * kRank1Threads - 1) / (2 * kRank1Threads)), (unsigned)((m + kNormCols - 1) / kNormCols));
    LAUNCH(k_rownorms_partial, grid, kRank1Threads, lp.Binv, lp.ld, m, m, kNormCols, lp.npart);
    LAUNCH(k_sum_norms, (m + 255) / 256, 256, lp.npart, lp.ld, m, (m + kNormCols - 1) / kNormCols, lp.w, ctx->d_st, 0);
}

// dual on the tableau: lp.d and lp.y of the current point.  d: the maintained row dj scattered to the variables (basic variables
// keep 0 / their starting value, dual :302); y from d through the basis B0 the tableau was built from: a_j^T y = c_j - d_j for
// j in B0, i.e. y = B0^-T (c_B0 - d_B0) -- a diagonal scaling when B0 was diagonal (slack basis), one transposed GEMV with the
// kept B0^-1 otherwise.  The reference accumulates y += theta_dual rho instead (dual :304); both satisfy d = c - A^T y.
int dual_tab_export(ellp_b200_ctx* ctx) {
    DevLP& lp = ctx->lp;
    const double* dpos = lp.dj;
    if (ctx->peer_mode && ctx->nranks > 1) {
        NCCL_TRY(nccl::api.AllGather(lp.dj, lp.dpos, (size_t)lp.nT, nccl::kFloat64, ctx->nccl_comm, ctx->stream));
        dpos = lp.dpos;
    }
    LAUNCH(k_dual_tab_scatter_d, (lp.nN + 255) / 256, 256, lp, dpos);
    LAUNCH(k_dual_tab_yrhs, (int)((lp.ld + 255) / 256), 256, lp, (const int32_t*)lp.Bv0, lp.u);
    if (ctx->tab_from_binv)
        LAUNCH(k_gemv_t<EPI_PLAIN>, gemv_grid(lp.m), 256, lp.Binv, lp.ld, (const int32_t*)nullptr, lp.m, lp.u, lp.y, (const double*)nullptr,
               (const uint8_t*)nullptr, (double*)nullptr, (PivotState*)nullptr, 0);
    else
        LAUNCH(k_dual_tab_y_diag, (lp.m + 255) / 256, 256, lp, (const double*)lp.u, (const double*)lp.bscale);
    return ELLP_OK;
}

double host_dual_obj(const ellp_std_form* sf, const double* y, const double* d) {  // standard_form.rs:52-68
    double o = 0.;
    for (int i = 0; i < sf->m; ++i) o += sf->b[i] * y[i];
    for (int i = 0; i < sf->n; ++i) {
        switch (sf->kind[i]) {
            case ELLP_FREE: break;
            case ELLP_LOWER: o += sf->lb[i] * d[i]; break;
            case ELLP_UPPER: o += sf->ub[i] * d[i]; break;
            case ELLP_TWOSIDED: o += (d[i] > 0.) ? sf->lb[i] * d[i] : sf->ub[i] * d[i]; break;
            default: o += sf->lb[i] * d[i];
        }
    }
    return o;
}

}  // namespace

// solve_trivial_problem (solvers/trivial/solve_trivial_problem.rs:5-96), host side: O(n) scalar work.
int ellp::host_solve_trivial(const ellp_std_form* sf, ellp_point* pt, bool minimize) {
    int nN = 0;
    for (int i = 0; i < sf->n; ++i) {
        const double c_i = sf->c[i];
        auto push = [&](int side) {
            if (pt->N) { pt->N[nN] = i; pt->N_side[nN] = (uint8_t)side; }
            ++nN;
        };
        switch (sf->kind[i]) {
            case ELLP_FREE:
                push(ELLP_NB_FREE);
                if (c_i != 0.) { pt->nN = nN; return ELLP_UNBOUNDED; }
                pt->x[i] = 0.;
                break;
            case ELLP_LOWER:
                push(ELLP_NB_LOWER);
                if (c_i > 0.) { if (minimize) pt->x[i] = sf->lb[i]; else { pt->nN = nN; return ELLP_UNBOUNDED; } }
                else if (minimize) pt->x[i] = sf->lb[i];
                else { if (c_i != 0.) { pt->nN = nN; return ELLP_UNBOUNDED; } pt->x[i] = sf->lb[i]; }
                break;
            case ELLP_UPPER:
                push(ELLP_NB_UPPER);
                if (c_i > 0.) { if (minimize) { pt->nN = nN; return ELLP_UNBOUNDED; } pt->x[i] = sf->ub[i]; }
                else if (minimize) { if (c_i != 0.) { pt->nN = nN; return ELLP_UNBOUNDED; } pt->x[i] = sf->ub[i]; }
                else pt->x[i] = sf->ub[i];
                break;
            case ELLP_TWOSIDED:
                if ((c_i > 0.) == minimize) { push(ELLP_NB_LOWER); pt->x[i] = sf->lb[i]; }
                else { push(ELLP_NB_UPPER); pt->x[i] = sf->ub[i]; }
                break;
            default:
                push(ELLP_NB_LOWER);
                pt->x[i] = sf->lb[i];
        }
    }
    pt->nN = nN;
    return ELLP_OPTIMAL;
}

static void peer_release(ellp_b200_ctx* ctx) {
    for (int r = 0; r < kMaxPeers; ++r)
        if (ctx->peer_map[r]) { cudaIpcCloseMemHandle(ctx->peer_map[r]); ctx->peer_map[r] = nullptr; }
    if (ctx->peer_own) { cudaFree(ctx->peer_own); ctx->peer_own = nullptr; }
    ctx->peer_cap = 0;
    ctx->pl = PeerLinks{};
}

extern "C" {

const char* ellp_b200_version(void) { return "ellp_b200 0.1 (sm_100a)"; }

int ellp_b200_create(int device, ellp_b200_ctx** out) {
    if (!out) return ELLP_E_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return ELLP_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return ELLP_E_CUDA;
    auto* ctx = new ellp_b200_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&ctx->d_st, sizeof(PivotState)) != cudaSuccess || cudaMalloc(&ctx->d_flag, 4 * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&ctx->d_sel, sizeof(SelScratch)) != cudaSuccess || cudaMemset(ctx->d_sel, 0, sizeof(SelScratch)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_st, sizeof(PivotState)) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        delete ctx;
        return ELLP_E_CUDA;
    }
    cudaDeviceSynchronize();  // the memset above ran on the legacy stream; ctx->stream is non-blocking
    std::memset(ctx->h_st, 0, sizeof(PivotState));
    if (cudaFuncSetAttribute(k_select_primal, cudaFuncAttributeMaxDynamicSharedMemorySize, kScanSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(k_ratio_pick, cudaFuncAttributeMaxDynamicSharedMemorySize, kScanSmemBytes) != cudaSuccess) {
        ellp_b200_destroy(ctx);
        return ELLP_E_CUDA;
    }
    *out = ctx;
    return ELLP_OK;
}

void ellp_b200_destroy(ellp_b200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    batch_free(ctx);
    peer_release(ctx);
    if (ctx->nccl_comm && nccl::api.CommDestroy) nccl::api.CommDestroy(ctx->nccl_comm);
    for (auto e : ctx->ev) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->d_st) cudaFree(ctx->d_st);
    if (ctx->d_flag) cudaFree(ctx->d_flag);
    if (ctx->d_sel) cudaFree(ctx->d_sel);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    if (ctx->tlog) cudaFree(ctx->tlog);
    for (auto e : ctx->chunk_ev) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->h_st) cudaFreeHost(ctx->h_st);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* ellp_b200_last_error(const ellp_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
void ellp_b200_set_error_(ellp_b200_ctx* ctx, const char* msg) { if (ctx) ctx->err = msg ? msg : ""; }
uint64_t ellp_b200_launch_count(const ellp_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
void ellp_b200_reset_launch_count(ellp_b200_ctx* ctx) { if (ctx) ctx->launches = 0; }

int ellp_b200_set_tuning(ellp_b200_ctx* ctx, const char* key, int value) {
    if (!ctx || !key) return ELLP_E_ARG;
    if (!std::strcmp(key, "rank1_cols_per_cta")) ctx->rank1_cols_per_cta = value;
    else if (!std::strcmp(key, "rank1_stream_min_mb")) ctx->rank1_stream_min_mb = value;
    else if (!std::strcmp(key, "refactor_mode")) ctx->refactor_mode = value;
    else if (!std::strcmp(key, "refactor_panel")) ctx->refactor_panel = value;
    else if (!std::strcmp(key, "flush_col_steps")) { ctx->flush_col_steps = std::max(1, value); ctx->flush2_col_steps = std::max(1, value); }
    else if (!std::strcmp(key, "flush_kernel")) ctx->flush_kernel = value;
    else if (!std::strcmp(key, "flush_waves")) ctx->flush_waves = value;
    else if (!std::strcmp(key, "flush4_min_k")) ctx->flush4_min_k = value;
    else if (!std::strcmp(key, "flush_ld")) ctx->flush_ld = value;
    else if (!std::strcmp(key, "flush_stages")) ctx->flush_stages = value;
    else if (!std::strcmp(key, "coop_pivots")) ctx->coop_pivots = value;
    else if (!std::strcmp(key, "cuda_graphs")) ctx->use_graphs = value;
    else if (!std::strcmp(key, "small_path")) ctx->small_path = value;
    else if (!std::strcmp(key, "peer_exchange")) ctx->peer_exchange = value;
    else if (!std::strcmp(key, "batch_pipeline")) ctx->batch_pipeline = value;
    else if (!std::strcmp(key, "owner_ratio")) ctx->owner_ratio = value;
    else if (!std::strcmp(key, "fast_upload")) ctx->fast_upload = value;
    else if (!std::strcmp(key, "residual_every")) ctx->residual_every = value;
    else if (!std::strcmp(key, "residual_tol_1e12")) ctx->residual_tol = 1e-12 * (double)value;
    else if (!std::strcmp(key, "coop_threads")) ctx->coop_threads = value;
    else if (!std::strcmp(key, "coop_ctas_per_sm")) { ctx->coop_ctas_per_sm = value; for (int& v : ctx->coop_threads_cached) v = 0; }
    else if (!std::strcmp(key, "phase_timing")) {  // value = pivots to log (0 = off); read back with ellp_b200_phase_log
        if (ctx->tlog) { cudaFree(ctx->tlog); ctx->tlog = nullptr; }
        ctx->tlog_cap = 0;
        if (value > 0) {
            CUDA_TRY(cudaMalloc(&ctx->tlog, sizeof(long long) * kTlogStamps * (size_t)value));
            CUDA_TRY(cudaMemset(ctx->tlog, 0, sizeof(long long) * kTlogStamps * (size_t)value));
            CUDA_TRY(cudaDeviceSynchronize());  // legacy-stream memset vs. the non-blocking ctx->stream
            ctx->tlog_cap = value;
            ctx->tlog_seq0 = ctx->xseq;
        }
    }
    else return set_err(ctx, ELLP_E_ARG, std::string("unknown tuning key ") + key);
    return ELLP_OK;
}

int ellp_b200_last_flush_kernel(ellp_b200_ctx* ctx) { return ctx ? ctx->last_flush_kernel : 0; }

int ellp_b200_phase_log(ellp_b200_ctx* ctx, int64_t* out, int32_t pivots) {
    if (!ctx || !out || !ctx->tlog || pivots > ctx->tlog_cap) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaMemcpy(out, ctx->tlog, sizeof(long long) * kTlogStamps * (size_t)pivots, cudaMemcpyDeviceToHost));
    return ELLP_OK;
}

void ellp_b200_default_opts(ellp_opts* o) {
    std::memset(o, 0, sizeof(*o));
    o->max_iter = 1000;  // {Primal,Dual}SimplexSolver::default() (primal :19-23, dual :20-24)
    o->tie_rule = ELLP_TIES_REFERENCE;
    o->engine = ELLP_ENGINE_AUTO;
}

int ellp_b200_upload(ellp_b200_ctx* ctx, const ellp_std_form* sf, const ellp_point* pt, int solver, const ellp_opts* o) {
    if (!ctx || !sf || !pt) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int m = sf->m, n = sf->n;
    if (m <= 0 || n < m) return set_err(ctx, ELLP_E_ARG, "upload needs 0 < m <= n");
    // the copies below read m / n - m entries: the lengths must match (solve_with_initial reports the reference's Err first)
    if (pt->nB != m || pt->nN != n - m) return set_err(ctx, ELLP_E_ARG, "upload: B / N lengths do not match the standard form");
    if (!pt->x || !pt->B || (n > m && (!pt->N || !pt->N_side))) return set_err(ctx, ELLP_E_ARG, "upload: x / B / N / N_side must not be NULL");
    if (solver == ELLP_DUAL && (!pt->y || !pt->d)) return set_err(ctx, ELLP_E_ARG, "dual solve needs y and d");
    for (int i = 0; i < pt->nB; ++i)  // the reference would panic on the first out-of-range column access (primal :144-148)
        if (pt->B[i] < 0 || pt->B[i] >= n) return set_err(ctx, ELLP_E_PANIC, "index out of bounds: basic variable index outside the standard form");
    for (int j = 0; j < pt->nN; ++j)
        if (pt->N[j] < 0 || pt->N[j] >= n) return set_err(ctx, ELLP_E_PANIC, "index out of bounds: nonbasic variable index outside the standard form");
    DevLP lp{};
    lp.m = m;
    lp.n = n;
    lp.nN = n - m;
    lp.n_glob = n;
    lp.col_lo = 0;
    lp.ld = (int64_t)align_up((size_t)m, 4);
    int kc = std::max(64, (m + 63) / 64);
    kc = std::min(kc, kFtranMaxKc);
    const int KS = (m + kc - 1) / kc;
    const int64_t tcap = (o && o->trace) ? o->trace_cap : 0;
    const bool tableau = o && o->engine == ELLP_ENGINE_TABLEAU;
    const int blk = blk_slots(o, tableau, solver);
    // tableau engines: room for B^-1 of a general starting basis (T = B^-1 A_N, y of the dual) unless the LP is so large that
    // the caller is expected to start from the slack basis (the fast condensed path below)
    const bool with_binv = tableau && ((double)m * 2.0 * m * 8.0 <= 24.0 * 1073741824.0);
    Arena probe;
    carve(probe, lp, KS, tcap, tableau, blk, false, 1, nullptr, nullptr, with_binv);
    if (int rc = ensure_arena(ctx, probe.off + 256)) return rc;
    Arena a;
    a.base = ctx->arena;
    RefactorScratch rf;
    carve(a, lp, KS, tcap, tableau, blk, false, 1, nullptr, nullptr, with_binv, &rf);
    ctx->has_binv = with_binv;
    ctx->rf_V = rf.V; ctx->rf_ldv = rf.ldv; ctx->rf_coop = rf.coop;
    ctx->tab_from_binv = false;
    ctx->dj_live = false;
    ctx->devex_live = false;
    ctx->b_inf = -1.;
    cudaStream_t s = ctx->stream;
    // zero the padded scratch once (padding rows must stay zero)
    CUDA_TRY(cudaMemsetAsync(lp.cB, 0, (size_t)((char*)lp.lam - (char*)lp.cB), s));
    CUDA_TRY(cudaMemsetAsync(lp.y, 0, sizeof(double) * lp.ld, s));
    if (tableau && lp.coop) CUDA_TRY(cudaMemsetAsync(lp.coop, 0, sizeof(double) * 6 * 1024, s));  // LL words of the fused pivot kernel: no stale sequence numbers
    // Condensed tableau, large LP: only the nonbasic columns are ever needed on the device when the starting basis is the
    // identity (slack / artificial basis -- what the reference's phase builders produce, primal_problem.rs:234-246).  Their
    // DMA goes straight into T while host threads verify that the basis columns are unit vectors; if they are, the basis
    // half of A never crosses PCIe.  Otherwise (or for small LPs) the whole matrix is uploaded and T is built on the device.
    bool fast_condensed = false, diag_ones = true;
    std::vector<double> diag0((size_t)std::max(m, 1), 1.0);
    if (tableau && lp.condensed && lp.nN > 0 && ctx->fast_upload && (double)m * n * 8.0 >= 32.0 * 1048576.0) {
        if (lp.ld != m) CUDA_TRY(cudaMemsetAsync(lp.T, 0, sizeof(double) * (size_t)lp.ld * lp.nN, s));
        for (int p0 = 0; p0 < lp.nN;) {  // runs of consecutive variable indices in N travel as one copy
            int p1 = p0 + 1;
            while (p1 < lp.nN && pt->N[p1] == pt->N[p1 - 1] + 1) ++p1;
            const double* src = sf->A + (size_t)pt->N[p0] * m;
            double* dst = lp.T + (size_t)p0 * lp.ld;
            if (lp.ld == m) CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)m * (p1 - p0), cudaMemcpyHostToDevice, s));
            else CUDA_TRY(cudaMemcpy2DAsync(dst, sizeof(double) * lp.ld, src, sizeof(double) * m, sizeof(double) * m, p1 - p0, cudaMemcpyHostToDevice, s));
            p0 = p1;
        }
        fast_condensed = host_basis_is_diagonal(sf->A, m, pt->B, diag0.data(), &diag_ones);
    }
    if (fast_condensed) {
        // A stays unpopulated on the device (ellp_b200_download_std_form reports that)
    } else if (lp.ld == m) {
        CUDA_TRY(cudaMemcpyAsync((void*)lp.A, sf->A, sizeof(double) * (size_t)m * n, cudaMemcpyHostToDevice, s));
    } else {
        CUDA_TRY(cudaMemsetAsync((void*)lp.A, 0, sizeof(double) * (size_t)lp.ld * n, s));
        CUDA_TRY(cudaMemcpy2DAsync((void*)lp.A, sizeof(double) * lp.ld, sf->A, sizeof(double) * m, sizeof(double) * m, n,
                                   cudaMemcpyHostToDevice, s));
    }
    CUDA_TRY(cudaMemcpyAsync((void*)lp.c, sf->c, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.b, sf->b, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.lb, sf->lb, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.ub, sf->ub, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.kind, sf->kind, (size_t)n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(lp.x, pt->x, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(lp.Bv, pt->B, sizeof(int32_t) * m, cudaMemcpyHostToDevice, s));
    if (lp.nN > 0) {
        CUDA_TRY(cudaMemcpyAsync(lp.Nv, pt->N, sizeof(int32_t) * lp.nN, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(lp.Ns, pt->N_side, (size_t)lp.nN, cudaMemcpyHostToDevice, s));
    }
    ctx->dual_obj0 = 0.;
    if (solver == ELLP_DUAL) {
        if (!pt->y || !pt->d) return set_err(ctx, ELLP_E_ARG, "dual solve needs y and d");
        CUDA_TRY(cudaMemcpyAsync(lp.y, pt->y, sizeof(double) * m, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(lp.d, pt->d, sizeof(double) * n, cudaMemcpyHostToDevice, s));
        ctx->dual_obj0 = host_dual_obj(sf, pt->y, pt->d);
    }
    ctx->lp = lp;
    ctx->KS = KS;
    ctx->kc = kc;
    ctx->trace_cap = tcap;
    ctx->solver = solver;
    ctx->resident = true;
    ctx->lp_generation++;
    ctx->tableau = tableau;
    ctx->blk_kmax = blk;
    ctx->blk_fill = 0;
    ctx->sharded = false;
    ctx->peer_mode = false;
    ctx->binv_valid = false;
    ctx->a_resident = !fast_condensed;
    if (fast_condensed) {  // T = A_N already sits in place: reduced-cost row d_p = c_p - c_B^T a_p, as at the end of refactor()
        LAUNCH(k_init_cB, (int)((lp.ld + 255) / 256), 256, lp);
        CUDA_TRY(cudaMemcpyAsync(lp.Bv0, lp.Bv, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToDevice, s));
        LAUNCH(k_fill_const, (int)((lp.ld + 255) / 256), 256, lp.bscale, (int64_t)lp.ld, 1.0);
        if (!diag_ones) {  // diagonal basis D (Gte rows: -1 slacks): T = D^-1 A_N, row by row
            CUDA_TRY(cudaMemcpyAsync(lp.bscale, diag0.data(), sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, s));
            dim3 gs((unsigned)((lp.ld / 2 + 255) / 256), (unsigned)std::min(lp.nN, 4096));
            LAUNCH(k_scale_rows_inv, gs, 256, lp.T, lp.ld, m, lp.nN, (const double*)lp.bscale);
        }
        if (solver == ELLP_DUAL) {
            LAUNCH(k_dual_tab_init, (lp.nT + 255) / 256, 256, lp);
            ctx->dj_live = true;
        } else {
            LAUNCH(k_gemv_t<EPI_PLAIN>, gemv_grid(lp.nN), 256, lp.T, lp.ld, (const int32_t*)nullptr, lp.nN, lp.cB, lp.dj, (const double*)nullptr,
                   (const uint8_t*)nullptr, (double*)nullptr, (PivotState*)nullptr, 0);
            LAUNCH(k_redcost_pos, (lp.nN + 255) / 256, 256, lp.c, lp.Nv, lp.nN, lp.dj);
        }
        ctx->binv_valid = true;
        ctx->pivots_since_refactor = 0;
    }
    // host buffers are only borrowed for the duration of the call
    CUDA_TRY(cudaStreamSynchronize(s));
    CUDA_TRY(cudaGetLastError());
    return ELLP_OK;
}

int ellp_b200_generate_dense(ellp_b200_ctx* ctx, int32_t m, int32_t n_struct, uint64_t seed, const ellp_opts* o) {
    return ellp_b200_generate_dense_ex(ctx, m, n_struct, seed, 0, o);
}

int ellp_b200_generate_dense_ex(ellp_b200_ctx* ctx, int32_t m, int32_t n_struct, uint64_t seed, int32_t variant, const ellp_opts* o) {
    if (!ctx || !o || m <= 0 || n_struct <= 0 || (m % 4) != 0 || variant < 0 || variant > 1)
        return set_err(ctx, ELLP_E_ARG, "generate_dense needs m % 4 == 0 and variant in {0,1}");
    CUDA_TRY(cudaSetDevice(ctx->device));
    DevLP lp{};
    lp.m = m;
    lp.n = n_struct + m;
    lp.nN = n_struct;
    lp.n_glob = lp.n;
    lp.col_lo = 0;
    lp.ld = m;
    int kc = std::min(kFtranMaxKc, std::max(64, (m + 63) / 64));
    const int KS = (m + kc - 1) / kc;
    const int64_t tcap = o->trace ? o->trace_cap : 0;
    const bool tableau = o->engine == ELLP_ENGINE_TABLEAU;
    const int blk = blk_slots(o, tableau, variant == 0 ? ELLP_PRIMAL : ELLP_DUAL);
    Arena probe;
    carve(probe, lp, KS, tcap, tableau, blk);
    if (int rc = ensure_arena(ctx, probe.off + 256)) return rc;
    Arena a;
    a.base = ctx->arena;
    carve(a, lp, KS, tcap, tableau, blk);
    ctx->has_binv = false;
    ctx->tab_from_binv = false;
    ctx->dj_live = false;
    ctx->devex_live = false;
    ctx->b_inf = -1.;
    CUDA_TRY(cudaMemsetAsync(lp.cB, 0, (size_t)((char*)lp.lam - (char*)lp.cB), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(lp.y, 0, sizeof(double) * lp.ld, ctx->stream));
    if (tableau && lp.coop) CUDA_TRY(cudaMemsetAsync(lp.coop, 0, sizeof(double) * 6 * 1024, ctx->stream));
    LAUNCH(k_gen_dense_cols, 148 * 16, 256, const_cast<double*>(lp.A), lp.ld, m, (int64_t)n_struct, (int64_t)0, (int64_t)lp.n, seed,
           variant == 0 ? 1.0 : -1.0, 1.0);
    LAUNCH(k_gen_dense_vectors, 148 * 2, 256, lp, (int64_t)n_struct, seed, (int)variant);
    ctx->lp = lp;
    ctx->KS = KS;
    ctx->kc = kc;
    ctx->trace_cap = tcap;
    ctx->solver = variant == 0 ? ELLP_PRIMAL : ELLP_DUAL;
    ctx->resident = true;
    ctx->lp_generation++;
    ctx->tableau = tableau;
    ctx->blk_kmax = blk;
    ctx->blk_fill = 0;
    ctx->sharded = false;
    ctx->peer_mode = false;
    ctx->binv_valid = false;
    ctx->a_resident = true;
    ctx->dual_obj0 = 0.;  // y = 0 and every bound is Lower(0): dual_obj(y, d) = 0
    if (tableau && variant == 1) {
        // slack basis B = -I (A x - s = b): T = B^-1 A_N = -A_N, generated straight into T; dj = d = c
        LAUNCH(k_gen_dense_cols, 148 * 16, 256, lp.T, lp.ld, m, (int64_t)n_struct, (int64_t)0, (int64_t)n_struct, seed, 1.0, -1.0);
        CUDA_TRY(cudaMemcpyAsync(lp.Bv0, lp.Bv, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToDevice, ctx->stream));
        LAUNCH(k_fill_const, (int)((lp.ld + 255) / 256), 256, lp.bscale, (int64_t)lp.ld, -1.0);
        LAUNCH(k_init_cB, (int)((lp.ld + 255) / 256), 256, lp);
        LAUNCH(k_dual_tab_init, (lp.nT + 255) / 256, 256, lp);
        ctx->dj_live = true;
        ctx->binv_valid = true;
        ctx->pivots_since_refactor = 0;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    return ELLP_OK;
}

/* copies the resident standard form + point to caller (host) buffers: used to obtain host copies of generated LPs */
int ellp_b200_download_std_form(ellp_b200_ctx* ctx, double* A, double* c, double* b, uint8_t* kind, double* lb, double* ub) {
    if (!ctx || !ctx->resident) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const DevLP& lp = ctx->lp;
    cudaStream_t s = ctx->stream;
    if (A && !ctx->a_resident) return set_err(ctx, ELLP_E_ARG, "A is not resident (condensed upload kept only the nonbasic columns)");
    if (A) CUDA_TRY(cudaMemcpy2DAsync(A, sizeof(double) * lp.m, lp.A, sizeof(double) * lp.ld, sizeof(double) * lp.m, ctx->peer_mode ? lp.nT : lp.n, cudaMemcpyDeviceToHost, s));
    if (c) CUDA_TRY(cudaMemcpyAsync(c, lp.c, sizeof(double) * lp.n_glob, cudaMemcpyDeviceToHost, s));
    if (b) CUDA_TRY(cudaMemcpyAsync(b, lp.b, sizeof(double) * lp.m, cudaMemcpyDeviceToHost, s));
    if (kind) CUDA_TRY(cudaMemcpyAsync(kind, lp.kind, (size_t)lp.n_glob, cudaMemcpyDeviceToHost, s));
    if (lb) CUDA_TRY(cudaMemcpyAsync(lb, lp.lb, sizeof(double) * lp.n_glob, cudaMemcpyDeviceToHost, s));
    if (ub) CUDA_TRY(cudaMemcpyAsync(ub, lp.ub, sizeof(double) * lp.n_glob, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return ELLP_OK;
}


// ---- column sharding: one rank per GPU ---------------------------------------------------------------------------
int ellp_b200_comm_unique_id(const char* nccl_path, void* out128) {
    std::string err;
    if (!out128 || !nccl::load(nccl_path, &err)) return ELLP_E_CUDA;
    nccl::UniqueId id;
    if (nccl::api.GetUniqueId(&id) != 0) return ELLP_E_CUDA;
    std::memcpy(out128, id.internal, 128);
    return ELLP_OK;
}

int ellp_b200_comm_init(ellp_b200_ctx* ctx, const char* nccl_path, const void* id128, int rank, int nranks) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    std::string err;
    if (!nccl::load(nccl_path, &err)) return set_err(ctx, ELLP_E_CUDA, err);
    nccl::UniqueId id;
    std::memcpy(id.internal, id128, 128);
    NCCL_TRY(nccl::api.CommInitRank(&ctx->nccl_comm, nranks, id, rank));
    ctx->rank = rank;
    ctx->nranks = nranks;
    return ELLP_OK;
}

static int sharded_prepare(ellp_b200_ctx* ctx, int32_t m, int32_t n_glob, const ellp_opts* o, DevLP* out) {
    if (!ctx->nccl_comm) return set_err(ctx, ELLP_E_ARG, "call ellp_b200_comm_init first");
    if (n_glob % ctx->nranks != 0) return set_err(ctx, ELLP_E_ARG, "column sharding needs n divisible by the number of ranks");
    if (m % 4 != 0) return set_err(ctx, ELLP_E_ARG, "column sharding needs m % 4 == 0");
    DevLP lp{};
    lp.m = m;
    lp.n_glob = n_glob;
    lp.n = n_glob / ctx->nranks;
    lp.col_lo = ctx->rank * lp.n;
    lp.nN = lp.n;
    lp.ld = m;
    const int64_t tcap = o->trace ? o->trace_cap : 0;
    const int blk = blk_slots(o, true);
    Arena probe;
    carve(probe, lp, 1, tcap, true, blk, true, ctx->nranks);
    if (int rc = ensure_arena(ctx, probe.off + 256)) return rc;
    Arena a;
    a.base = ctx->arena;
    carve(a, lp, 1, tcap, true, blk, true, ctx->nranks, &ctx->sendcol, &ctx->d_sides);
    ctx->blk_kmax = blk;
    ctx->blk_fill = 0;
    CUDA_TRY(cudaMemsetAsync(lp.cB, 0, (size_t)((char*)lp.lam - (char*)lp.cB), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(lp.y, 0, sizeof(double) * lp.ld, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(lp.xchg, 0, sizeof(double) * (64 + 4 * (size_t)ctx->nranks), ctx->stream));
    ctx->KS = 1;
    ctx->kc = 64;
    ctx->trace_cap = tcap;
    ctx->solver = ELLP_PRIMAL;
    ctx->tableau = true;
    ctx->sharded = true;
    ctx->dual_obj0 = 0.;
    *out = lp;
    return ELLP_OK;
}

static int sharded_finish_init(ellp_b200_ctx* ctx) {
    // reduced-cost row of the local columns (the starting basis is the identity, so T = A)
    DevLP& lp = ctx->lp;
    LAUNCH(k_init_cB, (int)((lp.ld + 255) / 256), 256, lp);
    LAUNCH(k_gemv_t<EPI_REDCOST>, gemv_grid(lp.n), 256, lp.T, lp.ld, (const int32_t*)nullptr, lp.n, lp.cB, lp.dj, lp.c + lp.col_lo,
           (const uint8_t*)nullptr, (double*)nullptr, (PivotState*)nullptr, 0);
    ctx->resident = true;
    ctx->lp_generation++;
    ctx->binv_valid = true;
    ctx->pivots_since_refactor = 0;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    return ELLP_OK;
}

int ellp_b200_sharded_generate_dense(ellp_b200_ctx* ctx, int32_t m, int32_t n_struct, uint64_t seed, const ellp_opts* o) {
    return ellp_b200_sharded_generate_dense_ex(ctx, m, n_struct, seed, 0, o);
}

int ellp_b200_sharded_generate_dense_ex(ellp_b200_ctx* ctx, int32_t m, int32_t n_struct, uint64_t seed, int32_t variant, const ellp_opts* o) {
    if (!ctx || !o || m <= 0 || n_struct <= 0 || variant < 0 || variant > 1) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    DevLP lp{};
    const int solver = variant == 0 ? ELLP_PRIMAL : ELLP_DUAL;
    if (variant == 1 && !(blk_slots(o, true, solver) > 1 && ctx->peer_exchange))
        return set_err(ctx, ELLP_E_ARG, "the sharded dual needs the peer-memory engine (block_k > 1)");
    if (blk_slots(o, true, solver) > 1 && ctx->peer_exchange) {
        // peer layout: rank g generates the structural columns (= nonbasic positions) [g nN/G, (g+1) nN/G) straight into T
        // (dual variant: slack basis -I, so T = B^-1 A_N = -A_N)
        if (int rc = peer_prepare(ctx, m, n_struct + m, o, &lp, solver)) return rc;
        LAUNCH(k_gen_dense_cols, 148 * 16, 256, lp.T, lp.ld, m, (int64_t)n_struct, (int64_t)lp.pos_lo, (int64_t)(lp.pos_lo + lp.nT), seed, 1.0,
               variant == 0 ? 1.0 : -1.0);
        LAUNCH(k_gen_dense_vectors, 148 * 2, 256, lp, (int64_t)n_struct, seed, (int)variant);
        LAUNCH(k_fill_const, (int)((lp.ld + 255) / 256), 256, lp.bscale, (int64_t)lp.ld, variant == 0 ? 1.0 : -1.0);
        ctx->lp = lp;
        return peer_finish_init(ctx);
    }
    ctx->peer_mode = false;
    if (int rc = sharded_prepare(ctx, m, n_struct + m, o, &lp)) return rc;
    LAUNCH(k_gen_dense_cols, 148 * 16, 256, const_cast<double*>(lp.A), lp.ld, m, (int64_t)n_struct, (int64_t)lp.col_lo,
           (int64_t)(lp.col_lo + lp.n), seed, 1.0, 1.0);
    LAUNCH(k_gen_dense_vectors, 148 * 2, 256, lp, (int64_t)n_struct, seed, 0);
    ctx->lp = lp;
    return sharded_finish_init(ctx);
}

int ellp_b200_sharded_upload(ellp_b200_ctx* ctx, const ellp_std_form* sf, const ellp_point* pt, const ellp_opts* o) {
    if (!ctx || !sf || !pt || !o) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int m = sf->m, ng = sf->n;
    if (pt->nB != m || pt->nN != ng - m) return set_err(ctx, ELLP_E_ARG, "sharded upload: B / N lengths do not match the standard form");
    DevLP lp{};
    ctx->peer_mode = false;
    if (int rc = sharded_prepare(ctx, m, ng, o, &lp)) return rc;
    cudaStream_t s = ctx->stream;
    // sf->A is THIS RANK's column block [col_lo, col_lo + n) (m x n, lda = m); the vectors are global
    CUDA_TRY(cudaMemcpyAsync((void*)lp.A, sf->A, sizeof(double) * (size_t)m * lp.n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.c, sf->c, sizeof(double) * ng, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.b, sf->b, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.lb, sf->lb, sizeof(double) * ng, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.ub, sf->ub, sizeof(double) * ng, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.kind, sf->kind, (size_t)ng, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(lp.x, pt->x, sizeof(double) * ng, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(lp.Bv, pt->B, sizeof(int32_t) * m, cudaMemcpyHostToDevice, s));
    std::vector<uint8_t> sides((size_t)ng, kColBasic);
    for (int j = 0; j < pt->nN; ++j) sides[pt->N[j]] = pt->N_side[j];
    CUDA_TRY(cudaMemcpyAsync(ctx->d_sides, sides.data(), (size_t)ng, cudaMemcpyHostToDevice, s));
    LAUNCH(k_shard_init, (lp.n + 255) / 256, 256, lp, ctx->d_sides);
    ctx->lp = lp;
    // the in-place tableau needs an identity starting basis: every rank checks the basis columns it owns
    CUDA_TRY(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), s));
    LAUNCH(k_check_identity_shard, m, 128, lp, ctx->d_flag);
    int mismatch = 0;
    CUDA_TRY(cudaMemcpyAsync(&mismatch, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (mismatch) return set_err(ctx, ELLP_E_ARG, "sharded tableau: the starting basis must be the identity (slack basis)");
    return sharded_finish_init(ctx);
}

int ellp_b200_sharded_upload_nonbasic(ellp_b200_ctx* ctx, const ellp_std_form* sf, const ellp_point* pt, const ellp_opts* o) {
    return ellp_b200_sharded_upload_nonbasic_ex(ctx, sf, pt, ELLP_PRIMAL, nullptr, o);
}

int ellp_b200_sharded_upload_nonbasic_ex(ellp_b200_ctx* ctx, const ellp_std_form* sf, const ellp_point* pt, int solver, const double* basis_diag,
                                         const ellp_opts* o) {
    if (!ctx || !sf || !pt || !o || (solver != ELLP_PRIMAL && solver != ELLP_DUAL)) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int m = sf->m, ng = sf->n;
    if (pt->nB != m || pt->nN != ng - m) return set_err(ctx, ELLP_E_ARG, "sharded upload: B / N lengths do not match the standard form");
    if (solver == ELLP_DUAL && (!pt->y || !pt->d)) return set_err(ctx, ELLP_E_ARG, "dual solve needs y and d");
    if (basis_diag)
        for (int i = 0; i < m; ++i)
            if (basis_diag[i] == 0. || basis_diag[i] != basis_diag[i]) return set_err(ctx, ELLP_E_ELLP, dev_err_message(kErrSingular));
    DevLP lp{};
    if (int rc = peer_prepare(ctx, m, ng, o, &lp, solver)) return rc;
    cudaStream_t s = ctx->stream;
    if (basis_diag) CUDA_TRY(cudaMemcpyAsync(lp.bscale, basis_diag, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, s));
    else LAUNCH(k_fill_const, (int)((lp.ld + 255) / 256), 256, lp.bscale, (int64_t)lp.ld, 1.0);
    if (solver == ELLP_DUAL) {
        CUDA_TRY(cudaMemcpyAsync(lp.y, pt->y, sizeof(double) * m, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(lp.d, pt->d, sizeof(double) * ng, cudaMemcpyHostToDevice, s));
        ctx->dual_obj0 = host_dual_obj(sf, pt->y, pt->d);
    }
    // sf->A = the columns of the nonbasic positions [pos_lo, pos_lo + nT) in pt->N order (m x nT, lda = m); vectors are global
    CUDA_TRY(cudaMemcpyAsync(lp.T, sf->A, sizeof(double) * (size_t)m * lp.nT, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.c, sf->c, sizeof(double) * ng, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.b, sf->b, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.lb, sf->lb, sizeof(double) * ng, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.ub, sf->ub, sizeof(double) * ng, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync((void*)lp.kind, sf->kind, (size_t)ng, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(lp.x, pt->x, sizeof(double) * ng, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(lp.Bv, pt->B, sizeof(int32_t) * m, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(lp.Nv, pt->N, sizeof(int32_t) * lp.nN, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(lp.Ns, pt->N_side, (size_t)lp.nN, cudaMemcpyHostToDevice, s));
    ctx->lp = lp;
    CUDA_TRY(cudaStreamSynchronize(s));  // host buffers are only borrowed for the duration of the call
    return peer_finish_init(ctx, basis_diag != nullptr);
}

// ---- K6: batches of independent small LPs ------------------------------------------------------------------------
static void batch_free(ellp_b200_ctx* ctx) {
    auto& B = ctx->batch;
    void* ptrs[] = {B.A, B.c, B.b, B.lb, B.ub, B.x, B.obj, B.kind, B.Ns, B.B, B.N, B.status, B.iters, B.err, B.trace_len, B.trace};
    for (void* p : ptrs) if (p) cudaFree(p);
    B = ellp_b200_ctx::Batch();
}

static int batch_alloc(ellp_b200_ctx* ctx, int nlp, int m, int n0, int trace_cap) {
    {   // same shape as the resident batch (the latency path solves one small LP after another): keep the buffers
        auto& R = ctx->batch;
        if (R.A && R.nlp == nlp && R.m == m && R.n0 == n0 && R.trace_cap == trace_cap) return ELLP_OK;
    }
    batch_free(ctx);
    auto& B = ctx->batch;
    const int nc = n0 + m;
    const size_t L = (size_t)nlp, nN = (size_t)n0;  // nc - m = n0 nonbasic positions
    // a failed allocation must not leave a half-built batch behind (the next call with the same shape would take the
    // "keep the buffers" return above and the kernel would write through null pointers): roll back, report, keep nlp = 0
    auto take = [&](void* pp, size_t bytes) {
        const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(pp), bytes);
        if (e != cudaSuccess) {
            ctx->err = std::string("cudaMalloc (batch): ") + cudaGetErrorString(e);
            cudaGetLastError();
            batch_free(ctx);
            return false;
        }
        return true;
    };
    if (!take(&B.A, sizeof(double) * L * m * n0) || !take(&B.c, sizeof(double) * L * n0) || !take(&B.b, sizeof(double) * L * m) ||
        !take(&B.lb, sizeof(double) * L * n0) || !take(&B.ub, sizeof(double) * L * n0) || !take(&B.kind, L * n0) ||
        !take(&B.x, sizeof(double) * L * nc) || !take(&B.obj, sizeof(double) * L) || !take(&B.B, sizeof(int32_t) * L * m) ||
        !take(&B.N, sizeof(int32_t) * L * nN) || !take(&B.Ns, L * nN) || !take(&B.status, sizeof(int32_t) * L) ||
        !take(&B.iters, sizeof(int32_t) * 2 * L) || !take(&B.err, sizeof(int32_t) * L) || !take(&B.trace_len, sizeof(int32_t) * L) ||
        (trace_cap > 0 && !take(&B.trace, sizeof(ellp_trace_rec) * L * trace_cap)))
        return ELLP_E_CUDA;
    // the shape is recorded only once every buffer exists
    B.nlp = nlp; B.m = m; B.n0 = n0; B.nc = nc; B.ld = (m % 2 == 0) ? m + 1 : m; B.trace_cap = trace_cap;
    return ELLP_OK;
}

int ellp_b200_batch_generate(ellp_b200_ctx* ctx, int32_t nlp, int32_t m, int32_t n_struct, uint64_t seed, int64_t first_lp,
                             int32_t trace_cap) {
    if (!ctx || nlp <= 0 || m <= 0 || n_struct <= 0) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (int rc = batch_alloc(ctx, nlp, m, n_struct + m, trace_cap)) return rc;
    auto& B = ctx->batch;
    LAUNCH(k_gen_batch, 148 * 16, 256, B.A, B.c, B.b, B.kind, B.lb, B.ub, nlp, m, n_struct, seed, first_lp);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    return ELLP_OK;
}

int ellp_b200_batch_upload(ellp_b200_ctx* ctx, const ellp_batch* bt, int32_t trace_cap) {
    if (!ctx || !bt || bt->nlp <= 0 || bt->m <= 0 || bt->n < bt->m) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (int rc = batch_alloc(ctx, bt->nlp, bt->m, bt->n, trace_cap)) return rc;
    auto& B = ctx->batch;
    const size_t L = (size_t)bt->nlp, m = (size_t)bt->m, n = (size_t)bt->n;
    cudaStream_t s = ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(B.A, bt->A, sizeof(double) * L * m * n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(B.c, bt->c, sizeof(double) * L * n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(B.b, bt->b, sizeof(double) * L * m, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(B.lb, bt->lb, sizeof(double) * L * n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(B.ub, bt->ub, sizeof(double) * L * n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(B.kind, bt->kind, L * n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return ELLP_OK;
}

// the 64 x 192 standard form of BASELINE.json configs[3] has a compile-time specialisation (fully unrolled rank-1 sweep)
using BatchKernel = void (*)(BatchArgs);
static BatchKernel batch_kernel_for(int m, int n0, int ld) {
    if (m == 64 && n0 == 192 && ld == 65) return k_batch_primal<64, 192>;
    return k_batch_primal<0, 0>;
}

int ellp_b200_batch_run(ellp_b200_ctx* ctx, const ellp_opts* o, ellp_batch_result* res) {
    if (!ctx || !o || !res) return ELLP_E_ARG;
    auto& B = ctx->batch;
    if (B.nlp <= 0) return set_err(ctx, ELLP_E_ARG, "no batch resident: call ellp_b200_batch_generate / ellp_b200_batch_upload first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t smem = batch_smem_bytes(B.m, B.n0, B.ld);
    if (smem > 227 * 1024) return set_err(ctx, ELLP_E_ARG, "LP too large for the shared-memory kernel (needs about (m+1)*n*8 bytes <= ~215 KB)");
    const BatchKernel kern = batch_kernel_for(B.m, B.n0, B.ld);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BatchArgs a{};
    a.nlp = B.nlp; a.m = B.m; a.n0 = B.n0; a.nc = B.nc; a.ld = B.ld;
    a.mode = 0;
    a.tie_rule = o->tie_rule;
    a.trace_cap = B.trace_cap;
    a.max_iter = o->max_iter;
    a.A = B.A; a.c = B.c; a.b = B.b; a.kind = B.kind; a.lb = B.lb; a.ub = B.ub;
    a.x = B.x; a.B = B.B; a.N = B.N; a.Ns = B.Ns;
    a.status = B.status; a.obj = B.obj; a.iters = B.iters; a.err = B.err; a.trace = B.trace; a.trace_len = B.trace_len;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    int per_sm = 1;  // persistent CTAs: as many per SM as the shared-memory tableau allows (2 for 64 x 192)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBatchThreads, smem);
    const int grid = std::min(B.nlp, sms * std::max(1, per_sm));
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    LAUNCH_SMEM(kern, grid, kBatchThreads, smem, a);
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res->ms_device = ms;
    res->launches = 1;
    return ELLP_OK;
}

int ellp_b200_batch_download(ellp_b200_ctx* ctx, ellp_batch_result* res) {
    if (!ctx || !res) return ELLP_E_ARG;
    auto& B = ctx->batch;
    if (B.nlp <= 0) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t L = (size_t)B.nlp;
    cudaStream_t s = ctx->stream;
    if (res->status) CUDA_TRY(cudaMemcpyAsync(res->status, B.status, sizeof(int32_t) * L, cudaMemcpyDeviceToHost, s));
    if (res->obj) CUDA_TRY(cudaMemcpyAsync(res->obj, B.obj, sizeof(double) * L, cudaMemcpyDeviceToHost, s));
    if (res->iters) CUDA_TRY(cudaMemcpyAsync(res->iters, B.iters, sizeof(int32_t) * 2 * L, cudaMemcpyDeviceToHost, s));
    if (res->err) CUDA_TRY(cudaMemcpyAsync(res->err, B.err, sizeof(int32_t) * L, cudaMemcpyDeviceToHost, s));
    if (res->x) CUDA_TRY(cudaMemcpyAsync(res->x, B.x, sizeof(double) * L * B.nc, cudaMemcpyDeviceToHost, s));
    if (res->trace_len) CUDA_TRY(cudaMemcpyAsync(res->trace_len, B.trace_len, sizeof(int32_t) * L, cudaMemcpyDeviceToHost, s));
    if (res->trace && B.trace) CUDA_TRY(cudaMemcpyAsync(res->trace, B.trace, sizeof(ellp_trace_rec) * L * B.trace_cap, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (res->iters) {
        uint64_t p = 0;
        for (size_t k = 0; k < 2 * L; ++k) p += (uint64_t)res->iters[k];
        res->pivots = p;
    }
    return ELLP_OK;
}

int ellp_b200_batch_download_lp(ellp_b200_ctx* ctx, int32_t k, double* A, double* c, double* b) {
    if (!ctx) return ELLP_E_ARG;
    auto& B = ctx->batch;
    if (k < 0 || k >= B.nlp) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (A) CUDA_TRY(cudaMemcpy(A, B.A + (size_t)k * B.m * B.n0, sizeof(double) * B.m * B.n0, cudaMemcpyDeviceToHost));
    if (c) CUDA_TRY(cudaMemcpy(c, B.c + (size_t)k * B.n0, sizeof(double) * B.n0, cudaMemcpyDeviceToHost));
    if (b) CUDA_TRY(cudaMemcpy(b, B.b + (size_t)k * B.m, sizeof(double) * B.m, cudaMemcpyDeviceToHost));
    return ELLP_OK;
}

int ellp_b200_batch_download_all(ellp_b200_ctx* ctx, double* A, double* c, double* b) {
    if (!ctx) return ELLP_E_ARG;
    auto& B = ctx->batch;
    if (B.nlp <= 0) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t L = (size_t)B.nlp;
    if (A) CUDA_TRY(cudaMemcpy(A, B.A, sizeof(double) * L * B.m * B.n0, cudaMemcpyDeviceToHost));
    if (c) CUDA_TRY(cudaMemcpy(c, B.c, sizeof(double) * L * B.n0, cudaMemcpyDeviceToHost));
    if (b) CUDA_TRY(cudaMemcpy(b, B.b, sizeof(double) * L * B.m, cudaMemcpyDeviceToHost));
    return ELLP_OK;
}

// Host buffers in, results out.  Large batches are pipelined: the batch is cut into chunks; the H2D copy of chunk c + 1 (copy
// stream) overlaps the pivoting of chunk c (compute stream), and the D2H of chunk c's results follows its kernel on the copy
// stream -- the PCIe transfer (6.4 GB for 65536 LPs of 64 x 192) hides behind the kernel instead of preceding it.
int ellp_b200_primal_solve_batch(ellp_b200_ctx* ctx, const ellp_batch* bt, const ellp_opts* o, ellp_batch_result* res) {
    if (!ctx || !bt || !o || !res) return ELLP_E_ARG;
    if (bt->nlp <= 0 || bt->m <= 0 || bt->n < bt->m) return ELLP_E_ARG;
    const int tcap = res->trace ? res->trace_cap : 0;
    const size_t lp_bytes = sizeof(double) * (size_t)bt->m * bt->n;
    const int nchunks = (int)std::min<size_t>(16, std::max<size_t>(1, ((size_t)bt->nlp * lp_bytes) >> 28));  // ~256 MB of A per chunk
    if (nchunks <= 1 || ctx->batch_pipeline == 0) {
        if (int rc = ellp_b200_batch_upload(ctx, bt, tcap)) return rc;
        if (int rc = ellp_b200_batch_run(ctx, o, res)) return rc;
        return ellp_b200_batch_download(ctx, res);
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (int rc = batch_alloc(ctx, bt->nlp, bt->m, bt->n, tcap)) return rc;
    auto& B = ctx->batch;
    const size_t smem = batch_smem_bytes(B.m, B.n0, B.ld);
    if (smem > 227 * 1024) return set_err(ctx, ELLP_E_ARG, "LP too large for the shared-memory kernel (needs about (m+1)*n*8 bytes <= ~215 KB)");
    const BatchKernel kern = batch_kernel_for(B.m, B.n0, B.ld);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!ctx->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (ctx->chunk_ev.size() < 2 * (size_t)nchunks) {
        const size_t old = ctx->chunk_ev.size();
        ctx->chunk_ev.resize(2 * (size_t)nchunks);
        for (size_t i = old; i < ctx->chunk_ev.size(); ++i) CUDA_TRY(cudaEventCreateWithFlags(&ctx->chunk_ev[i], cudaEventDisableTiming));
    }
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBatchThreads, smem);
    const size_t m = (size_t)bt->m, n = (size_t)bt->n, nc = (size_t)B.nc;
    const int per = (bt->nlp + nchunks - 1) / nchunks;
    if (!ctx->d2h_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    cudaStream_t cs = ctx->copy_stream, ks = ctx->stream, ds = ctx->d2h_stream;
    CUDA_TRY(cudaEventRecord(ctx->ev0, ks));
    CUDA_TRY(cudaStreamWaitEvent(cs, ctx->ev0, 0));  // the copies start after whatever was queued on the compute stream
    // three streams: H2D of chunk c + 1 (cs) and D2H of chunk c - 1 (ds) run under the kernel of chunk c (ks); the chunks use
    // disjoint regions of the device buffers, so the only dependencies are H2D(c) -> kernel(c) -> D2H(c)
    // the copies read / write the CALLER's buffers asynchronously: on any failure the three streams are drained before returning, so
    // that no DMA is still in flight when the caller gets its buffers back
    const int used = (bt->nlp + per - 1) / per;
    auto pipeline = [&]() -> int {
        auto h2d = [&](int c) -> int {
            const size_t l0 = (size_t)c * per, cnt = std::min<size_t>(per, (size_t)bt->nlp - l0);
            CUDA_TRY(cudaMemcpyAsync(B.A + l0 * m * n, bt->A + l0 * m * n, sizeof(double) * cnt * m * n, cudaMemcpyHostToDevice, cs));
            CUDA_TRY(cudaMemcpyAsync(B.c + l0 * n, bt->c + l0 * n, sizeof(double) * cnt * n, cudaMemcpyHostToDevice, cs));
            CUDA_TRY(cudaMemcpyAsync(B.b + l0 * m, bt->b + l0 * m, sizeof(double) * cnt * m, cudaMemcpyHostToDevice, cs));
            CUDA_TRY(cudaMemcpyAsync(B.lb + l0 * n, bt->lb + l0 * n, sizeof(double) * cnt * n, cudaMemcpyHostToDevice, cs));
            CUDA_TRY(cudaMemcpyAsync(B.ub + l0 * n, bt->ub + l0 * n, sizeof(double) * cnt * n, cudaMemcpyHostToDevice, cs));
            CUDA_TRY(cudaMemcpyAsync(B.kind + l0 * n, bt->kind + l0 * n, cnt * n, cudaMemcpyHostToDevice, cs));
            CUDA_TRY(cudaEventRecord(ctx->chunk_ev[2 * c], cs));
            return ELLP_OK;
        };
        if (int rc = h2d(0)) return rc;
        for (int c = 0; c < used; ++c) {
            const size_t l0 = (size_t)c * per, cnt = std::min<size_t>(per, (size_t)bt->nlp - l0);
            CUDA_TRY(cudaStreamWaitEvent(ks, ctx->chunk_ev[2 * c], 0));
            BatchArgs a{};
            a.nlp = (int)cnt; a.m = B.m; a.n0 = B.n0; a.nc = B.nc; a.ld = B.ld;
            a.mode = 0;
            a.tie_rule = o->tie_rule;
            a.trace_cap = B.trace_cap;
            a.max_iter = o->max_iter;
            a.A = B.A + l0 * m * n; a.c = B.c + l0 * n; a.b = B.b + l0 * m; a.kind = B.kind + l0 * n; a.lb = B.lb + l0 * n; a.ub = B.ub + l0 * n;
            a.x = B.x + l0 * nc; a.B = B.B + l0 * m; a.N = B.N + l0 * n; a.Ns = B.Ns + l0 * n;
            a.status = B.status + l0; a.obj = B.obj + l0; a.iters = B.iters + 2 * l0; a.err = B.err + l0;
            a.trace = B.trace ? B.trace + l0 * B.trace_cap : nullptr; a.trace_len = B.trace_len + l0;
            const int grid = (int)std::min<size_t>(cnt, (size_t)sms * std::max(1, per_sm));
            kern<<<grid, kBatchThreads, smem, ks>>>(a);
            ctx->launches++;
            CUDA_TRY(cudaEventRecord(ctx->chunk_ev[2 * c + 1], ks));
            if (c + 1 < used) { if (int rc = h2d(c + 1)) return rc; }
            CUDA_TRY(cudaStreamWaitEvent(ds, ctx->chunk_ev[2 * c + 1], 0));
            if (res->status) CUDA_TRY(cudaMemcpyAsync(res->status + l0, B.status + l0, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, ds));
            if (res->obj) CUDA_TRY(cudaMemcpyAsync(res->obj + l0, B.obj + l0, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ds));
            if (res->iters) CUDA_TRY(cudaMemcpyAsync(res->iters + 2 * l0, B.iters + 2 * l0, sizeof(int32_t) * 2 * cnt, cudaMemcpyDeviceToHost, ds));
            if (res->err) CUDA_TRY(cudaMemcpyAsync(res->err + l0, B.err + l0, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, ds));
            if (res->x) CUDA_TRY(cudaMemcpyAsync(res->x + l0 * nc, B.x + l0 * nc, sizeof(double) * cnt * nc, cudaMemcpyDeviceToHost, ds));
            if (res->trace_len) CUDA_TRY(cudaMemcpyAsync(res->trace_len + l0, B.trace_len + l0, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, ds));
            if (res->trace && B.trace) CUDA_TRY(cudaMemcpyAsync(res->trace + l0 * B.trace_cap, B.trace + l0 * B.trace_cap, sizeof(ellp_trace_rec) * cnt * B.trace_cap, cudaMemcpyDeviceToHost, ds));
        }
        return ELLP_OK;
    };
    if (int rc = pipeline()) {
        cudaStreamSynchronize(ks); cudaStreamSynchronize(cs); cudaStreamSynchronize(ds);
        return rc;
    }
    CUDA_TRY(cudaEventRecord(ctx->ev1, ks));
    CUDA_TRY(cudaStreamSynchronize(ks));
    CUDA_TRY(cudaStreamSynchronize(cs));
    CUDA_TRY(cudaStreamSynchronize(ds));
    CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res->ms_device = ms;
    res->launches = (uint64_t)used;
    if (res->iters) {
        uint64_t pv = 0;
        for (size_t k = 0; k < 2 * (size_t)bt->nlp; ++k) pv += (uint64_t)res->iters[k];
        res->pivots = pv;
    }
    return ELLP_OK;
}

int ellp_b200_run(ellp_b200_ctx* ctx, const ellp_opts* o, ellp_result* res) {
    if (!ctx || !o || !res) return ELLP_E_ARG;
    if (!ctx->resident) return set_err(ctx, ELLP_E_ARG, "no LP resident: call ellp_b200_upload first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    DevLP& lp = ctx->lp;
    const uint64_t launches0 = ctx->launches;
    std::memset(res, 0, sizeof(*res));
    PivotState& h = *ctx->h_st;
    std::memset(&h, 0, sizeof(h));
    h.status = kRunning;
    h.max_iter = o->max_iter;
    h.phase_tag = o->phase_tag;
    h.trace_cap = (o->trace && lp.trace) ? std::min<int64_t>(o->trace_cap, ctx->trace_cap) : 0;
    h.obj = ctx->dual_obj0;
    h.lmin_bits = 0x7ff0000000000000ll;
    if (o->max_iter == 0) h.status = ELLP_MAXITER;  // primal :163 (iter=1 > 0), dual :191 (0 >= 0)
    if (int rc = write_state(ctx)) return rc;
    const bool profile = o->profile != 0;
    if (profile && ctx->ev.size() < 8192) {
        const size_t old = ctx->ev.size();
        ctx->ev.resize(8192);
        for (size_t i = old; i < ctx->ev.size(); ++i) CUDA_TRY(cudaEventCreate(&ctx->ev[i]));
    }
    size_t ev_used = 0;
    if (!ctx->binv_valid) {
        if (int rc = refactor(ctx, &res->refactors)) return rc;
    }
    LAUNCH(k_init_cB, (int)((lp.ld + 255) / 256), 256, lp);
    const bool dse = ctx->solver == ELLP_DUAL && !ctx->tableau && o->pricing == ELLP_PRICE_STEEPEST_EDGE;
    if (dse) launch_row_norms(ctx);
    if (ctx->solver == ELLP_PRIMAL) LAUNCH(k_obj_dot, 1, 1024, lp.c, lp.x, lp.n_glob, ctx->d_st);
    // blocked tableau engine: slots were allocated at upload time; the caller may lower block_k per run
    int blk = (ctx->tableau && ctx->blk_kmax > 0 && o->block_k > 1) ? std::min(o->block_k, ctx->blk_kmax) : 0;
    const bool dual_tab = ctx->tableau && ctx->solver == ELLP_DUAL;
    // primal Harris ratio test (opt-in): implemented by the pick kernel of the kernel-per-phase paths (revised engine, rank-1 and
    // blocked tableau engines); the fused cooperative kernel keeps the reference's fold
    const bool primal_harris = ctx->solver == ELLP_PRIMAL && o->ratio == ELLP_RATIO_HARRIS;
    if (primal_harris && o->pricing == ELLP_PRICE_DEVEX) return set_err(ctx, ELLP_E_ARG, "primal: ELLP_PRICE_DEVEX (fused kernel) and ELLP_RATIO_HARRIS (kernel-per-phase paths) cannot be combined yet");
    if (primal_harris && (ctx->peer_mode || ctx->sharded)) return set_err(ctx, ELLP_E_ARG, "ELLP_RATIO_HARRIS for the primal is not available on the sharded engines");
    if (dual_tab) {  // the dual runs only on the blocked condensed tableau (single GPU or peer layout)
        if (blk == 0) blk = ctx->blk_kmax;
        if (blk <= 1 || !lp.condensed || (ctx->sharded && !ctx->peer_mode))
            return set_err(ctx, ELLP_E_ARG, "the dual tableau engine needs the condensed blocked layout (block_k > 1; sharded: the peer engine)");
        if ((o->pricing != ELLP_PRICE_REFERENCE && o->pricing != ELLP_PRICE_DEVEX) || o->ratio != ELLP_RATIO_REFERENCE)
            return set_err(ctx, ELLP_E_ARG, "the dual tableau engine prices with the reference rule or ELLP_PRICE_DEVEX; exact steepest edge / the Harris ratio test need ELLP_ENGINE_REVISED");
        if (o->pricing == ELLP_PRICE_DEVEX && !ctx->devex_live) {  // reference framework = the current basis
            LAUNCH(k_fill_const, (int)((lp.ld + 255) / 256), 256, lp.w, (int64_t)lp.ld, 1.0);
            ctx->devex_live = true;
        }
    } else if (o->pricing == ELLP_PRICE_DEVEX) {
        const bool fused_primal = ctx->tableau && blk > 1 && lp.condensed && (ctx->peer_mode || (!ctx->sharded && ctx->coop_pivots));
        if (!fused_primal) return set_err(ctx, ELLP_E_ARG, "ELLP_PRICE_DEVEX is implemented by the blocked tableau engines (ELLP_ENGINE_TABLEAU, block_k > 1)");
        if (!ctx->devex_live) {  // reference framework = the current nonbasic set
            LAUNCH(k_fill_const, (lp.nT + 255) / 256, 256, lp.wN, (int64_t)lp.nT, 1.0);
            ctx->devex_live = true;
        }
    }
    ctx->blk_fill = 0;
    if (blk > 0) { if (int rc = flush_attrs(ctx)) return rc; }
    int check_every = o->check_every > 0 ? o->check_every : 8;  // iterations enqueued per host read-back (finished solves make them no-ops)
    // default: netlib-sized LPs rebuild B^-1 (revised engine) / the tableau from A through B^-1 (tableau engines that kept room
    // for it) every 100 pivots, which bounds the drift of the updated matrix; large LPs refactor only on request
    const bool can_rebuild = !ctx->tableau || (ctx->has_binv && ctx->a_resident && lp.condensed && !ctx->peer_mode && !ctx->sharded);
    // (the dual and Devex-priced solves on the tableau every 25: on degenerate LPs -- netlib ADLITTLE's dual phase 1 -- a hundred accumulated updates
    // are enough for a pivot on rounding noise; the reference itself refactors at EVERY pivot)
    int refactor_every = o->refactor_every > 0 ? o->refactor_every : ((lp.m <= 512 && can_rebuild) ? ((dual_tab || (ctx->tableau && o->pricing == ELLP_PRICE_DEVEX)) ? 25 : 100) : 0);
    // a tableau can only be rebuilt from a resident constraint matrix: the condensed fast upload keeps no A, the peer layout
    // aliases A with its slice of T, the NCCL-sharded layout transforms A in place
    if (ctx->tableau && (!ctx->a_resident || ctx->peer_mode || ctx->sharded)) refactor_every = 0;
    // residual-triggered rebuild (tableau engines that can rebuild and do not already rebuild periodically)
    int residual_every = 0;
    uint64_t pivots_at_check = ctx->pivots_since_refactor;
    if (ctx->tableau && can_rebuild && lp.xpart && refactor_every == 0) {
        residual_every = ctx->residual_every >= 0 ? ctx->residual_every : (lp.m > 512 ? 1024 : 0);
        if (residual_every > 0 && ctx->b_inf < 0.) {
            std::vector<double> hb((size_t)lp.m);
            CUDA_TRY(cudaMemcpyAsync(hb.data(), lp.b, sizeof(double) * (size_t)lp.m, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            double mx = 0.;
            for (double v : hb) mx = std::max(mx, std::fabs(v));
            ctx->b_inf = mx;
        }
    }
    if (ctx->peer_mode && ctx->nranks > 1)  // stream-ordered barrier: no rank starts polling before every rank got here
        NCCL_TRY(nccl::api.AllReduce(lp.part, lp.part + 4, 1, nccl::kFloat64, nccl::kSum, ctx->nccl_comm, ctx->stream));
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc_loop = ELLP_OK;
    while (h.status == kRunning) {
        int batch = check_every;
        // never enqueue more iterations than the pivot budget still allows (they would be no-op launches)
        if (o->max_iter - h.pivots < (uint64_t)batch) batch = (int)std::max<uint64_t>(1, o->max_iter - h.pivots);
        if (refactor_every > 0) batch = (int)std::min<uint64_t>(batch, std::max<uint64_t>(1, refactor_every - ctx->pivots_since_refactor));
        if (ctx->peer_mode) {
            // peer engine: whole blocks of pivots per cooperative launch; every rank issues the same launches
            if (blk <= 1) { rc_loop = set_err(ctx, ELLP_E_ARG, "the peer-memory engine needs ellp_opts::block_k > 1"); break; }
            int left = std::max(batch, blk);
            if (o->max_iter - h.pivots < (uint64_t)left) left = (int)std::max<uint64_t>(1, o->max_iter - h.pivots);
            while (left > 0 && !rc_loop) {
                const int npiv = std::min(left, blk - ctx->blk_fill);
                rc_loop = launch_coop_pivots_peer(ctx, o, npiv, false);
                left -= npiv;
                if (!rc_loop && ctx->blk_fill >= blk) launch_flush(ctx, profile, &ev_used);
            }
            batch = 0;
        } else if (blk > 0 && !ctx->sharded && ((ctx->coop_pivots && !primal_harris) || dual_tab) && lp.condensed) {
            // cooperative path: whole blocks of pivots per launch, a flush after every full block
            int left = std::max(batch, blk);
            if (o->max_iter - h.pivots < (uint64_t)left) left = (int)std::max<uint64_t>(1, o->max_iter - h.pivots);
            if (ctx->peer_cap < lp.ld) {  // exchange buffer of the fused kernel (this GPU's own memory here)
                if (ctx->nranks == 1) rc_loop = peer_setup(ctx, lp.ld);
                else rc_loop = set_err(ctx, ELLP_E_ARG, "exchange buffer too small for a single-GPU LP on a multi-rank context");
            }
            while (left > 0 && !rc_loop) {
                const int npiv = std::min(left, blk - ctx->blk_fill);
                rc_loop = launch_coop_pivots_peer(ctx, o, npiv, true);
                left -= npiv;
                if (!rc_loop && ctx->blk_fill >= blk) launch_flush(ctx, profile, &ev_used);
            }
            batch = 0;
        }
        if (!ctx->tableau && !ctx->sharded && ctx->solver == ELLP_DUAL && ctx->small_path && !profile && lp.m <= kSmallMaxM &&
            o->pricing == ELLP_PRICE_REFERENCE && o->ratio == ELLP_RATIO_REFERENCE) {
            // netlib-sized dual: every iteration up to the next refactorisation / the pivot budget in ONE single-CTA launch
            uint64_t iters = std::min<uint64_t>(o->max_iter - h.pivots, 1u << 20);
            if (refactor_every > 0) iters = std::min<uint64_t>(iters, std::max<uint64_t>(1, refactor_every - ctx->pivots_since_refactor));
            LAUNCH(k_dual_small, 1, kSmallThreads, lp, ctx->kc, ctx->KS, (int)std::max<uint64_t>(1, iters), ctx->d_st);
            batch = 0;
        } else if (batch > 1 && batch == check_every && !ctx->tableau && !ctx->sharded && !profile && ctx->use_graphs) {
            // revised engine: replay `batch` iterations from a CUDA graph (kernels gate on PivotState::status, so iterations
            // after the end of the solve are no-ops exactly as with direct launches)
            const uint64_t key = ctx->lp_generation * 1000003ull + (uint64_t)(ctx->solver * 64 + o->pricing * 16 + o->ratio * 4 + o->tie_rule) * 131ull + (uint64_t)batch;
            if (!ctx->graph_exec || ctx->graph_key != key) {
                if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
                cudaGraph_t graph = nullptr;
                const uint64_t l0 = ctx->launches;
                CUDA_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
                for (int k = 0; k < batch; ++k) {
                    if (ctx->solver == ELLP_PRIMAL) launch_primal_iteration(ctx, o, false, &ev_used);
                    else launch_dual_iteration(ctx, o, false, &ev_used);
                }
                CUDA_TRY(cudaStreamEndCapture(ctx->stream, &graph));
                ctx->graph_launches = ctx->launches - l0;
                ctx->launches = l0;
                CUDA_TRY(cudaGraphInstantiate(&ctx->graph_exec, graph, 0));
                CUDA_TRY(cudaGraphDestroy(graph));
                ctx->graph_key = key;
                ctx->graph_iters = batch;
            }
            CUDA_TRY(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
            ctx->launches += ctx->graph_launches;
            batch = 0;
        }
        for (int k = 0; k < batch; ++k) {
            if (ctx->sharded) { if ((rc_loop = launch_sharded_iteration(ctx, o, profile, &ev_used, blk))) break; }
            else if (ctx->tableau) launch_tableau_primal_iteration(ctx, o, profile, &ev_used, blk);
            else if (ctx->solver == ELLP_PRIMAL) launch_primal_iteration(ctx, o, profile, &ev_used);
            else launch_dual_iteration(ctx, o, profile, &ev_used);
        }
        if (rc_loop) break;
        const uint64_t before = h.pivots;
        if ((rc_loop = read_state(ctx))) break;
        ctx->pivots_since_refactor += h.pivots - before;
        if (h.status != kRunning) break;
        if (residual_every > 0 && ctx->pivots_since_refactor - pivots_at_check >= (uint64_t)residual_every) {
            // long solve on a rebuildable tableau: |A x - b|_inf above tolerance => the accumulated updates have drifted, rebuild
            if (blk > 0) launch_flush(ctx, profile, &ev_used);
            double rsd = 0.;
            if ((rc_loop = primal_residual(ctx, &rsd))) break;
            ctx->last_residual = rsd;
            pivots_at_check = ctx->pivots_since_refactor;
            if (!(rsd <= ctx->residual_tol * (1.0 + ctx->b_inf))) {
                ctx->recompute_x = true;
                rc_loop = refactor(ctx, &res->refactors);
                ctx->recompute_x = false;
                if (rc_loop) break;
                pivots_at_check = 0;
                if (ctx->tableau && !ctx->a_resident) residual_every = 0;
            }
        }
        if (refactor_every > 0 && ctx->pivots_since_refactor >= (uint64_t)refactor_every) {
            if (blk > 0) launch_flush(ctx, profile, &ev_used);  // pending (U, V) slots belong to the tableau that is about to be replaced
            ctx->recompute_x = ctx->tableau;
            rc_loop = refactor(ctx, &res->refactors);
            ctx->recompute_x = false;
            if (rc_loop) break;
            if (dse) launch_row_norms(ctx);
            if (ctx->tableau && !ctx->a_resident) refactor_every = 0;  // the in-place rebuild consumed A: it cannot be repeated
        }
    }
    if (blk > 0) launch_flush(ctx, profile, &ev_used);  // leave a consistent tableau behind (the solve may be continued)
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (rc_loop) return rc_loop;
    CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res->ms_device = ms;
    if (profile) {
        for (size_t i = 0; i + 1 < ev_used; i += 2) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, ctx->ev[i], ctx->ev[i + 1]) == cudaSuccess) { res->ms_rank1 += t; res->n_rank1++; }
        }
    }
    res->status = h.status;
    res->iters = h.pivots;
    res->trace_len = h.trace_len;
    res->launches = ctx->launches - launches0;
    res->obj = h.obj;
    if (o->trace && h.trace_cap > 0) {
        const int64_t nrec = std::min<int64_t>(h.trace_len, h.trace_cap);
        if (nrec > 0) CUDA_TRY(cudaMemcpy(o->trace, lp.trace, sizeof(ellp_trace_rec) * (size_t)nrec, cudaMemcpyDeviceToHost));
    }
    if (h.err) {
        const int code = (h.err == kErrSingular) ? ELLP_E_ELLP : ELLP_E_PANIC;
        return set_err(ctx, code, dev_err_message(h.err));
    }
    return ELLP_OK;
}

int ellp_b200_download(ellp_b200_ctx* ctx, ellp_point* pt) {
    if (!ctx || !pt || !ctx->resident) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    DevLP& lp = ctx->lp;
    cudaStream_t s = ctx->stream;
    if (ctx->tableau && ctx->solver == ELLP_DUAL && ctx->dj_live && pt->y && pt->d) {
        if (int rc = dual_tab_export(ctx)) return rc;
    }
    if (ctx->peer_mode) {  // x, B and the N list (position order) are replicated
        CUDA_TRY(cudaMemcpyAsync(pt->x, lp.x, sizeof(double) * lp.n_glob, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(pt->B, lp.Bv, sizeof(int32_t) * lp.m, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(pt->N, lp.Nv, sizeof(int32_t) * lp.nN, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(pt->N_side, lp.Ns, (size_t)lp.nN, cudaMemcpyDeviceToHost, s));
        if (ctx->solver == ELLP_DUAL && pt->y && pt->d) {
            CUDA_TRY(cudaMemcpyAsync(pt->y, lp.y, sizeof(double) * lp.m, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaMemcpyAsync(pt->d, lp.d, sizeof(double) * lp.n_glob, cudaMemcpyDeviceToHost, s));
        }
        CUDA_TRY(cudaStreamSynchronize(s));
        return ELLP_OK;
    }
    if (ctx->sharded) {  // x and B are replicated; N is rebuilt from the gathered per-column status (ascending variable index)
        CUDA_TRY(cudaMemcpyAsync(pt->x, lp.x, sizeof(double) * lp.n_glob, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(pt->B, lp.Bv, sizeof(int32_t) * lp.m, cudaMemcpyDeviceToHost, s));
        NCCL_TRY(nccl::api.AllGather(lp.colstat, ctx->d_sides, (size_t)lp.n, nccl::kUint8, ctx->nccl_comm, s));
        std::vector<uint8_t> sides((size_t)lp.n_glob);
        CUDA_TRY(cudaMemcpyAsync(sides.data(), ctx->d_sides, (size_t)lp.n_glob, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        int k = 0;
        for (int j = 0; j < lp.n_glob && k < pt->nN; ++j)
            if (sides[j] != kColBasic) { pt->N[k] = j; pt->N_side[k] = sides[j]; ++k; }
        return ELLP_OK;
    }
    CUDA_TRY(cudaMemcpyAsync(pt->x, lp.x, sizeof(double) * lp.n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(pt->B, lp.Bv, sizeof(int32_t) * lp.m, cudaMemcpyDeviceToHost, s));
    if (lp.nN > 0) {
        CUDA_TRY(cudaMemcpyAsync(pt->N, lp.Nv, sizeof(int32_t) * lp.nN, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(pt->N_side, lp.Ns, (size_t)lp.nN, cudaMemcpyDeviceToHost, s));
    }
    if (ctx->solver == ELLP_DUAL && pt->y && pt->d) {
        CUDA_TRY(cudaMemcpyAsync(pt->y, lp.y, sizeof(double) * lp.m, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(pt->d, lp.d, sizeof(double) * lp.n, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return ELLP_OK;
}

static int solve_with_initial(ellp_b200_ctx* ctx, const ellp_std_form* sf, ellp_point* pt, const ellp_opts* o,
                              ellp_result* res, int solver) {
    if (!ctx || !sf || !pt || !o || !res) return ELLP_E_ARG;
    std::memset(res, 0, sizeof(*res));
    const int m = sf->m, n = sf->n;
    if (m == 0) {  // primal :118-122 / dual :132-136
        if (pt->nB != 0) return set_err(ctx, ELLP_E_PANIC, "assertion failed: B.is_empty()");
        res->status = ellp::host_solve_trivial(sf, pt, true);
        double obj = 0.;
        for (int i = 0; i < n; ++i) obj += sf->c[i] * pt->x[i];
        res->obj = obj;
        return ELLP_OK;
    }
    if (solver == ELLP_DUAL) {  // dual :139-151
        if (!pt->y || !pt->d) return set_err(ctx, ELLP_E_ARG, "dual solve needs y and d");
        for (int j = 0; j < pt->nN; ++j) {
            if (pt->N[j] < 0 || pt->N[j] >= n)  // d[N_i.index] would panic on the out-of-range index first
                return set_err(ctx, ELLP_E_PANIC, "index out of bounds: nonbasic variable index outside the standard form");
            const double d_i = pt->d[pt->N[j]];
            const int side = pt->N_side[j];
            const bool infeasible = side == ELLP_NB_LOWER ? d_i < -kEps : (side == ELLP_NB_UPPER ? d_i > kEps : std::fabs(d_i) > kEps);
            if (infeasible) return set_err(ctx, ELLP_E_PANIC, "initial point of dual phase 2 is dual infeasible");
        }
    }
    char buf[160];
    if (pt->nB != m) {  // primal :124-130 / dual :153-159
        std::snprintf(buf, sizeof buf, "invalid B, has %d elements but %d expected", pt->nB, m);
        return set_err(ctx, ELLP_E_ELLP, buf);
    }
    if (n < m) return set_err(ctx, ELLP_E_PANIC, "called `Option::unwrap()` on a `None` value");  // checked_sub
    if (pt->nN != n - m) {  // primal :134-140 / dual :163-169
        std::snprintf(buf, sizeof buf, "invalid N, has %d elements but %d expected", pt->nN, n - m);
        return set_err(ctx, ELLP_E_ELLP, buf);
    }
    if (n == m) {  // N empty: primal :149-151 / dual :175-177 (Optimal without any further check)
        res->status = ELLP_OPTIMAL;
        double obj = 0.;
        for (int i = 0; i < n; ++i) obj += sf->c[i] * pt->x[i];
        res->obj = solver == ELLP_PRIMAL ? obj : host_dual_obj(sf, pt->y, pt->d);
        return ELLP_OK;
    }
    if (solver == ELLP_DUAL && o->max_iter > 0 && pt->nB == m) {
        // Loop head of the dual (dual :188-246) on the caller's own point: no basic variable violates a bound => Optimal before any
        // LU / device work (the reference computes the LU first, but its only use would be the pivot that does not happen).  An
        // O(m) scan of host data; every dual phase 1 that starts feasible (netlib AFIRO) ends here without touching the GPU.
        bool any = false;
        for (int i = 0; i < m && !any; ++i) {
            const int v = pt->B[i];
            if (v < 0 || v >= n) { any = true; break; }  // let the upload report the out-of-range index
            const double x_i = pt->x[v];
            switch (sf->kind[v]) {
                case ELLP_LOWER: any = x_i < sf->lb[v] - kEps; break;
                case ELLP_UPPER: any = x_i > sf->ub[v] + kEps; break;
                case ELLP_TWOSIDED: any = (x_i > sf->ub[v] + kEps) || (x_i < sf->lb[v] - kEps); break;
                default: break;
            }
        }
        if (!any) {
            res->status = ELLP_OPTIMAL;
            res->obj = host_dual_obj(sf, pt->y, pt->d);
            return ELLP_OK;
        }
    }
    if (int rc = ellp_b200_upload(ctx, sf, pt, solver, o)) return rc;
    int rc = ellp_b200_run(ctx, o, res);
    if (rc) return rc;
    if ((rc = ellp_b200_download(ctx, pt))) return rc;
    if (solver == ELLP_PRIMAL) {  // StandardForm::obj (standard_form.rs:47-50)
        double obj = 0.;
        for (int i = 0; i < n; ++i) obj += sf->c[i] * pt->x[i];
        res->obj = obj;
    } else {
        res->obj = host_dual_obj(sf, pt->y, pt->d);
    }
    return ELLP_OK;
}

int ellp_b200_primal_solve_with_initial(ellp_b200_ctx* ctx, const ellp_std_form* sf, ellp_point* pt, const ellp_opts* o,
                                        ellp_result* res) {
    return solve_with_initial(ctx, sf, pt, o, res, ELLP_PRIMAL);
}

int ellp_b200_dual_solve_with_initial(ellp_b200_ctx* ctx, const ellp_std_form* sf, ellp_point* pt, const ellp_opts* o,
                                      ellp_result* res) {
    return solve_with_initial(ctx, sf, pt, o, res, ELLP_DUAL);
}

// ---- kernel-level entry points -------------------------------------------------------------------
int ellp_b200_dev_alloc(ellp_b200_ctx* ctx, uint64_t bytes, void** dptr) {
    if (!ctx || !dptr) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMalloc(dptr, bytes));
    return ELLP_OK;
}
int ellp_b200_dev_free(ellp_b200_ctx* ctx, void* dptr) {
    if (!ctx) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaFree(dptr));
    return ELLP_OK;
}
int ellp_b200_h2d(ellp_b200_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
    if (!ctx) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ELLP_OK;
}
int ellp_b200_d2h(ellp_b200_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
    if (!ctx) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ELLP_OK;
}
int ellp_b200_sync(ellp_b200_ctx* ctx) {
    if (!ctx) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    return ELLP_OK;
}
int ellp_b200_dev_fill_uniform(ellp_b200_ctx* ctx, double* dptr, uint64_t count, uint64_t seed, uint64_t offset, double lo,
                               double hi) {
    if (!ctx) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    LAUNCH(k_fill_uniform, 148 * 8, 256, dptr, count, seed, offset, lo, hi);
    CUDA_TRY(cudaGetLastError());
    return ELLP_OK;
}

__global__ void k_gather_row_plain(const double* __restrict__ E, int64_t ld, int64_t C, int64_t r, const double* __restrict__ alpha,
                                   double* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < C) out[j] = E[j * ld + r] / alpha[r];
}

int ellp_b200_rank1_update_dev(ellp_b200_ctx* ctx, double* E, int64_t R, int64_t C, int64_t ld, const double* alpha, int64_t r,
                               int32_t reps, float* ms_avg) {
    if (!ctx || !E || !alpha || R <= 0 || C <= 0 || ld < R || r < 0 || r >= R || reps < 1) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    double* prow = nullptr;
    double* apad = nullptr;
    const int64_t Rp = (int64_t)align_up((size_t)R, 2);
    CUDA_TRY(cudaMalloc(&prow, sizeof(double) * C));
    CUDA_TRY(cudaMalloc(&apad, sizeof(double) * Rp));
    CUDA_TRY(cudaMemsetAsync(apad, 0, sizeof(double) * Rp, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(apad, alpha, sizeof(double) * R, cudaMemcpyDeviceToDevice, ctx->stream));
    LAUNCH(k_gather_row_plain, (unsigned)((C + 255) / 256), 256, E, ld, C, r, alpha, prow);
    const bool vec = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(E) & 15) == 0);
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int k = 0; k < reps; ++k) {
        if (vec) {
            launch_rank1(ctx, E, ld, (int)R, (int)C, apad, prow, nullptr, (int)r);
        } else {
            dim3 grid((unsigned)((R + 255) / 256), (unsigned)std::min<int64_t>(C, 1024));
            LAUNCH(k_rank1_scalar, grid, 256, E, ld, (int)R, (int)C, apad, prow, (int)r);
        }
    }
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (ms_avg) *ms_avg = ms / (float)reps;
    cudaFree(prow);
    cudaFree(apad);
    return ELLP_OK;
}

int ellp_b200_rank1_update(ellp_b200_ctx* ctx, double* E, int64_t R, int64_t C, int64_t ld, const double* alpha, int64_t r) {
    if (!ctx || !E || !alpha || R <= 0 || C <= 0 || ld < R) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    double* dE = nullptr;
    double* da = nullptr;
    CUDA_TRY(cudaMalloc(&dE, sizeof(double) * ld * C));
    CUDA_TRY(cudaMalloc(&da, sizeof(double) * R));
    CUDA_TRY(cudaMemcpy(dE, E, sizeof(double) * ld * C, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(da, alpha, sizeof(double) * R, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaDeviceSynchronize());  // the copies above ran on the legacy stream; ctx->stream is non-blocking and would not wait for them
    int rc = ellp_b200_rank1_update_dev(ctx, dE, R, C, ld, da, r, 1, nullptr);
    if (rc == ELLP_OK) CUDA_TRY(cudaMemcpy(E, dE, sizeof(double) * ld * C, cudaMemcpyDeviceToHost));
    cudaFree(dE);
    cudaFree(da);
    return rc;
}

// K3b on caller-supplied DEVICE data: E (R x C, ld) -= U (R x k, ld) * V (k x C, row stride ldv); reps timed launches
int ellp_b200_rankk_update_dev(ellp_b200_ctx* ctx, double* E, int64_t R, int64_t C, int64_t ld, const double* U, const double* V,
                               int64_t ldv, int32_t k, int32_t reps, float* ms_avg) {
    if (!ctx || !E || !U || !V || R <= 0 || C <= 0 || ld < R || (ld % 2) != 0 || ldv < C || k < 1 || k > kBlkMax || reps < 1 ||
        (reinterpret_cast<uintptr_t>(E) & 15) != 0)
        return set_err(ctx, ELLP_E_ARG, "rankk_update needs ld even, ld >= R, ldv >= C, 1 <= k <= 64 and a 16-byte aligned E");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (int rc = flush_attrs(ctx)) return rc;
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int t = 0; t < reps; ++t) launch_rankk(ctx, E, ld, (int)ld, (int)C, U, V, ldv, (int)k);  // rows up to ld: padding rows of U are zero
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (ms_avg) *ms_avg = ms / (float)reps;
    return ELLP_OK;
}

// same on host buffers (E: R x C with leading dimension ld, U: R x k with leading dimension R, V: k x C row-major)
int ellp_b200_rankk_update(ellp_b200_ctx* ctx, double* E, int64_t R, int64_t C, int64_t ld, const double* U, const double* V, int32_t k) {
    if (!ctx || !E || !U || !V || R <= 0 || C <= 0 || ld < R || k < 1 || k > kBlkMax) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t ldp = (int64_t)align_up((size_t)R, 4);
    double *dE = nullptr, *dU = nullptr, *dV = nullptr;
    CUDA_TRY(cudaMalloc(&dE, sizeof(double) * ldp * C));
    CUDA_TRY(cudaMalloc(&dU, sizeof(double) * ldp * k));
    const int64_t ldvp = (int64_t)align_up((size_t)C, 128);  // rows of V padded to whole 64-column tiles (k_blk_flush3)
    CUDA_TRY(cudaMalloc(&dV, sizeof(double) * ldvp * k));
    CUDA_TRY(cudaMemset(dE, 0, sizeof(double) * ldp * C));
    CUDA_TRY(cudaMemset(dU, 0, sizeof(double) * ldp * k));
    CUDA_TRY(cudaMemset(dV, 0, sizeof(double) * ldvp * k));
    CUDA_TRY(cudaMemcpy2D(dE, sizeof(double) * ldp, E, sizeof(double) * ld, sizeof(double) * R, C, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy2D(dU, sizeof(double) * ldp, U, sizeof(double) * R, sizeof(double) * R, k, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy2D(dV, sizeof(double) * ldvp, V, sizeof(double) * C, sizeof(double) * C, k, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaDeviceSynchronize());  // the copies above ran on the legacy stream; ctx->stream is non-blocking and would not wait for them
    int rc = ellp_b200_rankk_update_dev(ctx, dE, R, C, ldp, dU, dV, ldvp, k, 1, nullptr);
    if (rc == ELLP_OK) CUDA_TRY(cudaMemcpy2D(E, sizeof(double) * ld, dE, sizeof(double) * ldp, sizeof(double) * R, C, cudaMemcpyDeviceToHost));
    cudaFree(dE); cudaFree(dU); cudaFree(dV);
    return rc;
}

int ellp_b200_gemv_t(ellp_b200_ctx* ctx, const double* M, int64_t R, int64_t C, int64_t ld, const int32_t* cols, int64_t ncols,
                     const double* v, double* y) {
    if (!ctx || !M || !v || !y || R <= 0 || C <= 0 || ld < R || ncols <= 0) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t ldp = (int64_t)align_up((size_t)R, 4);
    double *dM = nullptr, *dv = nullptr, *dy = nullptr;
    int32_t* dc = nullptr;
    CUDA_TRY(cudaMalloc(&dM, sizeof(double) * ldp * C));
    CUDA_TRY(cudaMalloc(&dv, sizeof(double) * ldp));
    CUDA_TRY(cudaMalloc(&dy, sizeof(double) * ncols));
    CUDA_TRY(cudaMemset(dM, 0, sizeof(double) * ldp * C));
    CUDA_TRY(cudaMemset(dv, 0, sizeof(double) * ldp));
    CUDA_TRY(cudaMemcpy2D(dM, sizeof(double) * ldp, M, sizeof(double) * ld, sizeof(double) * R, C, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dv, v, sizeof(double) * R, cudaMemcpyHostToDevice));
    if (cols) {
        CUDA_TRY(cudaMalloc(&dc, sizeof(int32_t) * ncols));
        CUDA_TRY(cudaMemcpy(dc, cols, sizeof(int32_t) * ncols, cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaDeviceSynchronize());  // the copies above ran on the legacy stream; ctx->stream is non-blocking and would not wait for them
    LAUNCH(k_gemv_t<EPI_PLAIN>, gemv_grid((int)ncols), 256, dM, ldp, dc, (int)ncols, dv, dy, (const double*)nullptr,
           (const uint8_t*)nullptr, (double*)nullptr, (PivotState*)nullptr, 0);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(y, dy, sizeof(double) * ncols, cudaMemcpyDeviceToHost));
    cudaFree(dM); cudaFree(dv); cudaFree(dy);
    if (dc) cudaFree(dc);
    return ELLP_OK;
}

int ellp_b200_gemv_n(ellp_b200_ctx* ctx, const double* M, int64_t R, int64_t C, int64_t ld, const double* v, double* y) {
    if (!ctx || !M || !v || !y || R <= 0 || C <= 0 || ld < R) return ELLP_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t ldp = (int64_t)align_up((size_t)R, 4);
    int kc = std::min(kFtranMaxKc, std::max(64, (int)((C + 63) / 64)));
    const int KS = (int)((C + kc - 1) / kc);
    double *dM = nullptr, *dv = nullptr, *dy = nullptr, *dp = nullptr;
    CUDA_TRY(cudaMalloc(&dM, sizeof(double) * ldp * C));
    CUDA_TRY(cudaMalloc(&dv, sizeof(double) * C));
    CUDA_TRY(cudaMalloc(&dy, sizeof(double) * R));
    CUDA_TRY(cudaMalloc(&dp, sizeof(double) * ldp * KS));
    CUDA_TRY(cudaMemset(dM, 0, sizeof(double) * ldp * C));
    CUDA_TRY(cudaMemcpy2D(dM, sizeof(double) * ldp, M, sizeof(double) * ld, sizeof(double) * R, C, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dv, v, sizeof(double) * C, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaDeviceSynchronize());  // the copies above ran on the legacy stream; ctx->stream is non-blocking and would not wait for them
    dim3 grid((unsigned)((ldp + 255) / 256), (unsigned)KS);
    LAUNCH(k_gemv_n_partial, grid, 128, dM, ldp, (int)R, (int)C, dv, dp, kc);
    LAUNCH(k_sum_partials, (unsigned)((R + 255) / 256), 256, dp, ldp, (int)R, KS, dy);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(y, dy, sizeof(double) * R, cudaMemcpyDeviceToHost));
    cudaFree(dM); cudaFree(dv); cudaFree(dy); cudaFree(dp);
    return ELLP_OK;
}

int ellp_b200_refactor_bench(ellp_b200_ctx* ctx, int32_t m, uint64_t seed, int32_t mode, int32_t reps, float* ms_avg) {
    // times the refactorisation (K4) of a dense random m x m basis that never leaves HBM
    if (!ctx || m <= 0 || (m % 4) != 0 || reps < 1) return ELLP_E_ARG;
    ellp_opts o;
    ellp_b200_default_opts(&o);
    o.engine = ELLP_ENGINE_REVISED;
    if (int rc = ellp_b200_generate_dense_ex(ctx, m, m, seed, 0, &o)) return rc;
    std::vector<int32_t> B((size_t)m);
    for (int i = 0; i < m; ++i) B[i] = i;  // basis = the m dense structural columns
    CUDA_TRY(cudaMemcpy(ctx->lp.Bv, B.data(), sizeof(int32_t) * m, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaDeviceSynchronize());  // the copies above ran on the legacy stream; ctx->stream is non-blocking and would not wait for them
    std::memset(ctx->h_st, 0, sizeof(PivotState));
    if (int rc = write_state(ctx)) return rc;
    const int saved = ctx->refactor_mode;
    ctx->refactor_mode = mode;
    int rc = refactor(ctx, nullptr);  // warm-up
    float total = 0.f;
    for (int k = 0; k < reps && rc == ELLP_OK; ++k) {
        cudaEventRecord(ctx->ev0, ctx->stream);
        rc = refactor(ctx, nullptr);
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        total += ms;
    }
    ctx->refactor_mode = saved;
    if (ms_avg) *ms_avg = total / (float)reps;
    return rc;
}

int ellp_b200_invert(ellp_b200_ctx* ctx, const double* Bmat, int64_t m, double* Binv) {
    if (!ctx || !Bmat || !Binv || m <= 0) return ELLP_E_ARG;
    // run the resident-LP refactorisation on an LP whose A is Bmat and whose basis is 0..m-1
    std::vector<double> zeros((size_t)m, 0.0);
    std::vector<uint8_t> kind((size_t)m, ELLP_LOWER);
    std::vector<int32_t> B((size_t)m);
    for (int64_t i = 0; i < m; ++i) B[i] = (int32_t)i;
    ellp_std_form sf{(int32_t)m, (int32_t)m, Bmat, zeros.data(), zeros.data(), kind.data(), zeros.data(), zeros.data()};
    std::vector<double> x((size_t)m, 0.0);
    ellp_point pt{x.data(), B.data(), nullptr, nullptr, nullptr, nullptr, (int32_t)m, 0};
    ellp_opts o;
    ellp_b200_default_opts(&o);
    if (int rc = ellp_b200_upload(ctx, &sf, &pt, ELLP_PRIMAL, &o)) return rc;
    std::memset(ctx->h_st, 0, sizeof(PivotState));
    if (int rc = write_state(ctx)) return rc;
    if (int rc = refactor(ctx, nullptr)) return rc;
    const DevLP& lp = ctx->lp;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // the copy below runs on the legacy stream, which does not wait for ctx->stream
    CUDA_TRY(cudaMemcpy2D(Binv, sizeof(double) * m, lp.Binv, sizeof(double) * lp.ld, sizeof(double) * m, m, cudaMemcpyDeviceToHost));
    return ELLP_OK;
}

}  // extern "C"
