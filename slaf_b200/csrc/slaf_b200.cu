// slaf_b200.cu -- C-ABI implementation (see include/slaf_b200.h for the contract and the
// reference interfaces each entry point replaces).
//
// Stream model.  Each context owns a side stream `s_in`, a helper thread and eight pipeline slots.
//   submit : [s_in]    wait(caller stream) -> H2D of the COO columns (host input, piece by piece, straight from
//                      the caller's pinned / registered memory) -> K1 CSR build, which also writes {n_cells, error
//                      flag} into mapped pinned memory -> event `front`
//            [helper]  (slafb_prepare_shuffle) waits `front`, builds the CPython-exact permutation
//   collect: [host]    waits `front` (and the helper); uploads the permutation
//            [stream]  wait(front) -> K2 tokenize -> K2g general path (or K6 rank-value) -> event `done`
// so the H2D copy and CSR build of the next batches overlap the tokenization of the current one, there is no
// device-to-host copy in either stream, and the only host<->device round trip per batch is the wait on `front`.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/slaf_b200.h"
#include "kernels.cuh"
#include "py_shuffle.h"
#include "pack_rows.h"

using namespace slafb;

#ifndef SLAFB_K1_LDG_DEFAULT
#define SLAFB_K1_LDG_DEFAULT false
#endif

namespace {

constexpr int kSlots = 8;
constexpr int kHelpers = 2;   // host threads preparing CPython-exact shuffles (a 32k-cell shuffle costs about as much host time as the GPU needs for the batch)

struct Slot {
    bool in_flight = false;
    uint64_t freed_at = 0;             // release order (pick_slot)
    // device scratch
    uint32_t *seg_start = nullptr, *seg_cell = nullptr;
    int32_t *perm_dev = nullptr;
    uint32_t *work_list = nullptr;
    uint32_t *counters = nullptr;   // [0] n_cells, [1] err, [2] work count, [3] work head, [4] fast-kernel row ticket
    // staged copies of host input (lazy)
    void *in_cell = nullptr, *in_gene = nullptr, *in_value = nullptr;
    // pinned host
    uint32_t *h_counters = nullptr; // [0] n_cells, [1] err, [2] cells deferred to the general kernel (mapped pinned: kernels write it)
    uint32_t *h_counters_dev = nullptr;  // device-side address of h_counters
    bool all_general = false;
    bool host_segs = false;            // the segment table came from the host (submit_segments / submit_csi): the cell count is known without the GPU
    bool packed = false;               // the rows arrived in the packed wire format (submit_packed): pk_bytes / blk_off describe them
    uint8_t *pk_bytes = nullptr;       // device copy of the packed blocks (lazy, grow-only)
    size_t pk_cap = 0;
    uint32_t *blk_off = nullptr;       // device [max_cells + 2] block offsets, 16-byte units (lazy)
    uint32_t *h_blk = nullptr;         // pinned host copy the table is assembled in (lazy)
    // permutation prepared ahead of collect() by the helper thread (slafb_prepare_shuffle)
    std::atomic<int> prep_state{0};   // 0 none, 1 queued / running, 2 ready
    int64_t prep_seed = 0;
    int64_t prep_cells = -1;
    int32_t *h_perm = nullptr;
    char *h_small = nullptr;           // pinned scratch small pageable host pieces are staged through (three column regions)
    size_t small_rows = 0;             // rows per column region of h_small
    uint32_t *h_seg = nullptr;         // pinned copy of a caller-supplied segment table (submit_segments): [C + 1] starts, then [C] cell ids
    uint32_t *h_draws = nullptr;    // scratch of the host shuffle (plain malloc)
    // batch description (device pointers actually used by the kernels)
    const void *cell = nullptr, *gene = nullptr, *value = nullptr;
    int64_t n_rows = 0;
    int32_t cell_dtype = 0, gene_dtype = 0, value_dtype = 0;
    cudaEvent_t ev_front = nullptr, ev_done = nullptr, ev_in = nullptr;
    cudaEvent_t ev_loaded = nullptr;   // all three columns are on the device (ev_front: cell column + CSR build only)
    cudaEvent_t ev_perm = nullptr;     // the helper thread's upload of the prepared permutation
    bool perm_uploaded = false;
    int32_t *h_perm_dev = nullptr;     // device-side address of h_perm (mapped pinned memory)
    // timing events
    cudaEvent_t t_h2d0 = nullptr, t_h2d1 = nullptr, t_k1a = nullptr, t_k1b = nullptr, t_k2a = nullptr, t_k2b = nullptr, t_k2c = nullptr;
    bool timed_front = false, timed_back = false, timed_h2d = false;
    bool sample = true;               // this batch carries the per-kernel timing events
    bool ran_back = false;
    bool used_f32_warp = false;        // the float32 warp-per-cell kernel ran this batch (its deferral rate steers the choice of kernel)
    int64_t back_cells = 0;
};

}  // namespace

struct slafb_ctx {
    int device = 0;
    int n_sm = 148;
    int64_t max_rows = 0, max_cells = 0;
    cudaStream_t s_in = nullptr;
    cudaStream_t s_aux = nullptr;      // small device-to-host reads that must not queue behind the big H2D copies
    cudaStream_t s_perm = nullptr;     // the helper thread's permutation uploads
    Slot slot[kSlots];
    uint64_t free_seq = 0;
    unsigned long long *tile_state = nullptr;
    uint32_t *ticket = nullptr;
    uint32_t ticket_base = 0, epoch = 0;
    double *log1p_lut = nullptr;
    int log1pf_fdlibm = 0;            // the host's log1pf is the fdlibm float algorithm (glibc <= 2.39): the device uses its restatement
    uint32_t *vstar_table = nullptr;  // binned window on uint16 storage: bin>=1 threshold per cell maximum
    int vstar_nbins = -1;
    uint32_t *sort_scratch = nullptr;
    unsigned long long *rank_scratch = nullptr;   // rank-value extension: composites of very large cells (lazy)
    unsigned long long *stats_packed = nullptr;   // gene statistics: packed (sum << 24 | count) reductions of one launch (lazy, grow-only)
    size_t stats_genes = 0;
    unsigned long long *dup_table = nullptr;   // contiguity check (lazy, grow-only)
    size_t dup_slots = 0;
    // window facade scratch
    uint32_t *kept_count = nullptr;
    long long *total_dev = nullptr;
    long long *h_total = nullptr;
    slafb_timing acc{};
    char err[512] = {0};
    // helper thread: builds CPython-exact permutations as soon as a batch's cell count is known
    std::thread worker[kHelpers];      // shuffle helpers: the batches' permutations are independent, each job is one slot
    std::mutex wm;
    std::condition_variable wcv, wdone;
    std::deque<int> wjobs;
    bool wstop = false;
    bool wstarted = false;
    // per-kernel CUDA-event timing is sampled: every `timing_every`-th batch carries the seven events
    // (each event record is a stream operation; on every batch they cost a few percent of a step)
    int timing_every = 1;
    uint64_t batch_seq = 0;
    bool front_on_caller = false;
    cudaEvent_t ev_k1 = nullptr;       // end of the latest CSR build: K1 launches share the context's span ticket / tile states, so a K1
    cudaStream_t k1_stream = nullptr;  // on another stream than the previous one (SLAFB_OPT_FRONT_ON_CALLER) waits for it first
    bool k1_any = false;
    bool f32_warp_off = false;        // float32 values: the warp-per-cell kernel deferred too many cells of a batch (more distinct values per cell
                                      // than its set holds): this context stays with the CTA-per-cell kernel
    bool last_fast_f32_warp = false;
    int64_t recent_slow = 1 << 20;    // decayed maximum of the cells recently deferred to the general kernel (starts pessimistic)
};

// A steady pipeline over one context: batches go in (submit + shuffle preparation), the oldest batch in flight comes out into
// the next entry of a ring of caller-owned output tensors.  One call per step; nothing is allocated per step.
struct slafb_pipe {
    slafb_ctx *ctx = nullptr;
    slafb_params p{};
    std::vector<slafb_out> ring;
    int depth = 1;
    struct Item { int32_t slot; int64_t seed; int64_t seq; };
    std::deque<Item> q;
    int64_t next_seq = 0;
    int32_t next_ring = 0;
};

namespace {

thread_local char g_err[512] = "";

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int fail(slafb_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s", buf);
    snprintf(g_err, sizeof g_err, "%s", buf);
    return code;
}

#define CU(ctx, call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(ctx, SLAFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

size_t dtype_size(int dt) { return dt == SLAFB_U16 ? 2 : 4; }

int check_coo(slafb_ctx *ctx, const slafb_coo *in) {
    if (!in) return fail(ctx, SLAFB_ERR_INVALID, "coo is NULL");
    if (in->n_rows < 0) return fail(ctx, SLAFB_ERR_INVALID, "n_rows < 0");
    if (in->n_rows > ctx->max_rows) return fail(ctx, SLAFB_ERR_CAPACITY, "n_rows %lld exceeds context max_rows %lld", (long long)in->n_rows, (long long)ctx->max_rows);
    if (in->cell_dtype != SLAFB_U32 && in->cell_dtype != SLAFB_I32) return fail(ctx, SLAFB_ERR_INVALID, "cell dtype must be uint32 or int32");
    if (in->gene_dtype != SLAFB_U16 && in->gene_dtype != SLAFB_I32 && in->gene_dtype != SLAFB_U32) return fail(ctx, SLAFB_ERR_INVALID, "gene dtype must be uint16, int32 or uint32");
    if (in->value_dtype != SLAFB_U16 && in->value_dtype != SLAFB_F32) return fail(ctx, SLAFB_ERR_INVALID, "value dtype must be uint16 or float32");
    if (in->n_rows > 0) {
        if (!in->cell || !in->gene || !in->value) return fail(ctx, SLAFB_ERR_INVALID, "NULL column pointer");
        if (in->location == SLAFB_LOC_DEVICE &&
            (((uintptr_t)in->cell | (uintptr_t)in->gene | (uintptr_t)in->value) & 15u))
            return fail(ctx, SLAFB_ERR_INVALID, "device column pointers must be 16-byte aligned");
    }
    return SLAFB_OK;
}

int check_params(slafb_ctx *ctx, const slafb_params *p) {
    if (!p) return fail(ctx, SLAFB_ERR_INVALID, "params is NULL");
    if (p->mode < 0 || p->mode > 3) return fail(ctx, SLAFB_ERR_INVALID, "unknown mode %d", p->mode);
    if (p->max_genes < 1) return fail(ctx, SLAFB_ERR_INVALID, "max_genes must be >= 1, got %d", p->max_genes);
    if (p->max_items < 1) return fail(ctx, SLAFB_ERR_INVALID, "max_items must be >= 1, got %d", p->max_items);
    if (p->mode >= SLAFB_SCGPT_RAW && p->n_bins < 1) return fail(ctx, SLAFB_ERR_INVALID, "n_bins must be >= 1");
    if (p->window_n_bins < 0) return fail(ctx, SLAFB_ERR_INVALID, "window_n_bins must be >= 0");
    if (p->order != SLAFB_ORDER_ROW && p->order != SLAFB_ORDER_RANK) return fail(ctx, SLAFB_ERR_INVALID, "unknown order %d", p->order);
    if ((p->order == SLAFB_ORDER_RANK || p->norm) && p->mode != SLAFB_GENEFORMER)
        return fail(ctx, SLAFB_ERR_INVALID, "rank order / per-gene normalisation are Geneformer-mode extensions");
    if (p->norm && p->n_norm < 1) return fail(ctx, SLAFB_ERR_INVALID, "norm given but n_norm < 1");
    if (p->mode == SLAFB_GENEFORMER_PCT && !(p->min_percentile >= 0.0 && p->min_percentile <= 100.0))
        return fail(ctx, SLAFB_ERR_INVALID, "min_percentile must be within [0, 100]");
    return SLAFB_OK;
}

// accumulate the event timings of a slot's previous use (events are complete: ev_done was waited on)
void harvest(slafb_ctx *ctx, Slot &s) {
    float ms;
    if (s.timed_h2d && cudaEventElapsedTime(&ms, s.t_h2d0, s.t_h2d1) == cudaSuccess) {
        float k1 = 0.f;                       // the CSR build sits between the cell column and the other two
        if (s.timed_front) cudaEventElapsedTime(&k1, s.t_k1a, s.t_k1b);
        ctx->acc.h2d_ms += ms - k1;
    }
    if (s.timed_front && cudaEventElapsedTime(&ms, s.t_k1a, s.t_k1b) == cudaSuccess) ctx->acc.csr_ms += ms;
    if (s.timed_front || s.timed_h2d) ctx->acc.timed_batches++;     // (host-table routes run no CSR build: the copy events alone mark them)
    if (s.timed_back) {
        if (cudaEventElapsedTime(&ms, s.t_k2a, s.t_k2b) == cudaSuccess) ctx->acc.tokenize_ms += ms;
        if (cudaEventElapsedTime(&ms, s.t_k2b, s.t_k2c) == cudaSuccess) ctx->acc.slowpath_ms += ms;
    }
    if (s.ran_back) {
        const int64_t slow = s.all_general ? (int64_t)s.h_counters[0] : (int64_t)s.h_counters[2];
        ctx->acc.slow_cells += slow;
        // how much work the general kernel found lately (it is launched after every fast kernel, usually for nothing)
        ctx->recent_slow = slow > ctx->recent_slow ? slow : (ctx->recent_slow * 7) / 8;
        if (s.used_f32_warp && s.back_cells >= 64 && slow * 8 > s.back_cells) ctx->f32_warp_off = true;
    }
    s.timed_front = s.timed_back = s.timed_h2d = s.ran_back = s.used_f32_warp = false;
}

// Which log1pf does the host's libm ship -- the one polars calls on float32 columns (kernels.cuh: log1pf_window)?  A sweep over
// 2^20 non-negative floats (every 2,040th bit pattern) against the fdlibm restatement and against the correctly rounded value:
// 1 = the host matches the fdlibm algorithm everywhere (glibc <= 2.39), 0 = it matches the correctly rounded value (glibc >= 2.40)
// or neither (an unknown libm gets the correctly rounded evaluation).  Evaluated once per process.
int host_log1pf_kind() {
    static const int kind = [] {
        bool is_fdlibm = true, is_cr = true;
        for (uint32_t b = 0; b < 0x7f800000u; b += 2040u) {
            const float x = slafb_u2f(b), h = log1pf(x);
            const float f = log1pf_fdlibm(x), c = (float)log1p((double)x);
            is_fdlibm &= memcmp(&h, &f, 4) == 0;
            is_cr &= memcmp(&h, &c, 4) == 0;
        }
        return (is_fdlibm && !is_cr) ? 1 : 0;
    }();
    return kind;
}

int ensure_staging(slafb_ctx *ctx, Slot &s) {
    if (s.in_cell) return SLAFB_OK;
    const size_t n = (size_t)ctx->max_rows + 64;
    CU(ctx, cudaMalloc(&s.in_cell, n * 4));
    CU(ctx, cudaMalloc(&s.in_gene, n * 4));
    CU(ctx, cudaMalloc(&s.in_value, n * 4));
    return SLAFB_OK;
}

int ensure_sort_scratch(slafb_ctx *ctx) {
    if (ctx->sort_scratch) return SLAFB_OK;
    CU(ctx, cudaMalloc(&ctx->sort_scratch, ((size_t)ctx->max_rows + 64) * 4));
    return SLAFB_OK;
}

// ---- kernel dispatch over (gene dtype, value dtype, mode) -------------------------------------
// K2 fast path.  Returns 1 when the configuration does not fit the fast kernel (then every cell
// goes through the general kernel).
template <typename GeneT, typename ValT, int MODE, bool PACKED = false>
int launch_fast(slafb_ctx *ctx, const TokArgs &a, cudaStream_t st, bool *launched, const uint8_t *packed = nullptr, const uint32_t *blk_off = nullptr) {
    *launched = false;
    constexpr bool SCGPT = (MODE == MODE_SCGPT_RAW || MODE == MODE_SCGPT_BINNED);
    const uint32_t limit = SCGPT ? a.max_genes : (a.L - 1u);
    FastArgs fa;
    fa.t = a;
    fa.gcap = (limit + 15u + 7u) & ~7u;
    // value stage: whole cells up to ~4k rows for uint16; float32 rows are twice as wide, so the stage holds 2,560 and the
    // (rarer) longer cells read their tail from global memory -- more CTAs per SM matter more than staging every row
    uint32_t vmin = sizeof(ValT) == 2 ? 4104u : 2568u;
    if (const char *env = getenv("SLAFB_VCAP")) {   // tuning knob: rows of values staged per cell
        const long v = atol(env);
        if (v >= 256 && v <= 16384) vmin = (uint32_t)(v + 7) & ~7u;
    }
    fa.vcap = fa.gcap > vmin ? fa.gcap : vmin;
    if (MODE == MODE_GF_PCT) fa.gcap = fa.vcap;          // percentile mode: any row may survive the filter, so every row's gene is staged too
    fa.vstar_table = ctx->vstar_table;
    fa.ticket = a.work_counters + 2;   // slot counters[4], zeroed by K1 together with the work-list counters
    fa.packed = packed;
    fa.blk_off = blk_off;
    // packed route: two raw-block stages (header + two byte streams of up to vcap - 8 rows + room for 64 exceptions) and ONE
    // decoded gene / value stage; otherwise two gene / value stages the bulk copies fill directly
    fa.rawcap = PACKED ? (uint32_t)((16u + 2u * (size_t)(fa.vcap - 8u) + 512u + 127u) & ~127u) : 0u;
    const size_t smem = PACKED ? 2 * (size_t)fa.rawcap + (size_t)fa.gcap * sizeof(GeneT) + (size_t)fa.vcap * sizeof(ValT)
                               : 2 * ((size_t)fa.gcap * sizeof(GeneT) + (size_t)fa.vcap * sizeof(ValT)) + (sizeof(ValT) == 4 ? (size_t)kHashSlots * 4 : 0);
    if (smem > 200 * 1024) return SLAFB_OK;   // very wide rows: general kernel only
    auto kern = tokenize_cells_kernel<GeneT, ValT, MODE, PACKED>;
    static std::atomic<bool> attr_set[64];   // zero-initialised; setting the attribute twice is harmless
    if (!attr_set[ctx->device & 63].load(std::memory_order_acquire)) {
        CU(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[ctx->device & 63].store(true, std::memory_order_release);
    }
    if (MODE == MODE_SCGPT_BINNED && sizeof(ValT) == 2 && ctx->vstar_nbins != a.w_bins) {
        vstar_table_kernel<<<256, 256, 0, st>>>(ctx->log1p_lut, a.w_bins, ctx->vstar_table);
        CU(ctx, cudaGetLastError());
        ctx->vstar_nbins = a.w_bins;
        ctx->acc.kernel_launches++;
    }
    size_t per_sm = (220 * 1024) / (smem + 1024);
    size_t per_sm_max = 2048 / kFastThreads;
    if (const char *env = getenv("SLAFB_FAST_CTAS_PER_SM")) {   // tuning knob
        const long v = atol(env);
        if (v >= 1 && v <= 32) per_sm_max = (size_t)v;
    }
    if (per_sm_max > (size_t)fast_ctas_per_sm<ValT, MODE>()) per_sm_max = (size_t)fast_ctas_per_sm<ValT, MODE>();   // what the registers allow
    if (per_sm > per_sm_max) per_sm = per_sm_max;
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t)(ctx->n_sm * per_sm);
    if (a.n_cells < grid) grid = a.n_cells;
    kern<<<grid, kFastThreads, smem, st>>>(fa);
    CU(ctx, cudaGetLastError());
    ctx->acc.kernel_launches++;
    *launched = true;
    return SLAFB_OK;
}

template <typename GeneT, typename ValT, int MODE, int SINK>
int launch_general(slafb_ctx *ctx, const TokArgs &a, const WindowSink &w, cudaStream_t st) {
    auto kern = tokenize_general_kernel<GeneT, ValT, MODE, SINK>;
    const size_t dyn = (sizeof(ValT) == 2) ? (size_t)kBitmapWords * 4 : (size_t)kSortSmemKeys * 4;
    static std::atomic<bool> attr_set[64];  // per instantiation, per device (zero-initialised; setting the attribute twice is harmless)
    if (!attr_set[ctx->device & 63].load(std::memory_order_acquire)) {
        CU(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        attr_set[ctx->device & 63].store(true, std::memory_order_release);
    }
    // persistent CTAs over a ticket: as many per SM as shared memory allows (static TokSmem ~17 KB + bitmap / sort keys)
    size_t per_sm = (220 * 1024) / (dyn + 18 * 1024);
    if (per_sm > 2048 / kTokThreads) per_sm = 2048 / kTokThreads;
    if (per_sm < 1) per_sm = 1;
    // behind a fast kernel the work list is normally empty: when the recent batches deferred (almost) nothing, one CTA
    // per SM is plenty and the launch itself is cheaper; any count is still processed correctly, only by fewer CTAs
    if (!a.all_cells && SINK == 0 && ctx->recent_slow < (int64_t)ctx->n_sm) per_sm = 1;
    uint32_t grid = (uint32_t)(ctx->n_sm * per_sm);
    if (a.n_cells < grid) grid = a.n_cells;
    if (grid == 0) grid = 1;
    kern<<<grid, kTokThreads, dyn, st>>>(a, w);
    CU(ctx, cudaGetLastError());
    ctx->acc.kernel_launches++;
    return SLAFB_OK;
}

template <typename GeneT, typename ValT>
int launch_rank(slafb_ctx *ctx, const RankArgs &ra0, cudaStream_t st) {
    RankArgs ra = ra0;
    // K6b first (bucket distribution instead of a sorting network; kernels.cuh): it tokenizes every cell it can and leaves the rest
    // -- cells with heavy ties or more rows than its shared-memory array -- on the work list, which the sorting kernel then drains.
    // SLAFB_RANK_SORT=1 keeps the sorting kernel for every cell (A/B runs).
    static const bool sort_only = [] { const char *e = getenv("SLAFB_RANK_SORT"); return e && atoi(e) != 0; }();
    ra.from_list = 0;
    if (!sort_only) {
        // E = 16 rows per thread over all cells, then E = 32 over what it left (cells of 4,097..8,192 rows; the others stay listed)
        auto b16 = tokenize_rank_bucket_kernel<GeneT, ValT, 16>;
        auto b32 = tokenize_rank_bucket_kernel<GeneT, ValT, 32>;
        const size_t dyn16 = rank_bucket_smem<16>(), dyn32 = rank_bucket_smem<32>();
        static std::atomic<int> b_per_sm_cached[64], b32_per_sm_cached[64];
        int p16 = b_per_sm_cached[ctx->device & 63].load(std::memory_order_acquire), p32 = b32_per_sm_cached[ctx->device & 63].load(std::memory_order_acquire);
        if (p16 == 0 || p32 == 0) {
            CU(ctx, cudaFuncSetAttribute(b16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn16));
            CU(ctx, cudaFuncSetAttribute(b32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn32));
            CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p16, b16, kTokThreads, dyn16));
            CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p32, b32, kTokThreads, dyn32));
            if (p16 < 1) p16 = 1;
            if (p32 < 1) p32 = 1;
            b32_per_sm_cached[ctx->device & 63].store(p32, std::memory_order_release);
            b_per_sm_cached[ctx->device & 63].store(p16, std::memory_order_release);
        }
        uint32_t g16 = (uint32_t)(ctx->n_sm * p16);
        if (ra.t.n_cells < g16) g16 = ra.t.n_cells;
        if (g16 == 0) g16 = 1;
        b16<<<g16, kTokThreads, dyn16, st>>>(ra);
        CU(ctx, cudaGetLastError());
        ra.from_list = 1;
        uint32_t g32 = (uint32_t)(ctx->n_sm * p32);
        if (ra.t.n_cells < g32) g32 = ra.t.n_cells;
        if (g32 == 0) g32 = 1;
        b32<<<g32, kTokThreads, dyn32, st>>>(ra);
        CU(ctx, cudaGetLastError());
        ctx->acc.kernel_launches += 2;
    }
    auto kern = tokenize_rank_kernel<GeneT, ValT>;
    const size_t dyn = (size_t)kRankSmemSlots * sizeof(unsigned long long);
    static std::atomic<bool> attr_set[64];   // zero-initialised; setting the attribute twice is harmless
    if (!attr_set[ctx->device & 63].load(std::memory_order_acquire)) {
        CU(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        attr_set[ctx->device & 63].store(true, std::memory_order_release);
    }
    static std::atomic<int> per_sm_cached[64];       // persistent CTAs over a ticket: as many as are resident (registers: the sort keeps 16 keys per thread)
    int per_sm = per_sm_cached[ctx->device & 63].load(std::memory_order_acquire);
    if (per_sm == 0) {
        CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTokThreads, dyn));
        if (per_sm < 1) per_sm = 1;
        per_sm_cached[ctx->device & 63].store(per_sm, std::memory_order_release);
    }
    uint32_t grid = (uint32_t)(ctx->n_sm * per_sm);
    if (ra.t.n_cells < grid) grid = ra.t.n_cells;
    if (grid == 0) grid = 1;
    kern<<<grid, kTokThreads, dyn, st>>>(ra);
    CU(ctx, cudaGetLastError());
    ctx->acc.kernel_launches++;
    return SLAFB_OK;
}

// percentile mode on uint16 values: one warp per cell (kernels.cuh, K2w).  SLAFB_PCT_VSTAGE=0|1 and SLAFB_PCT_VCAP=<rows> select
// the variant and the stage size for A/B runs (defaults: the measured best).
template <typename GeneT, bool VSTAGE>
int launch_pct_warp_v(slafb_ctx *ctx, const TokArgs &a, cudaStream_t st, uint32_t vcap, bool *launched) {
    FastArgs fa;
    memset(&fa, 0, sizeof fa);
    fa.t = a;
    fa.vcap = fa.gcap = vcap;
    fa.ticket = a.work_counters + 2;
    const size_t smem = (size_t)kPctWarps * pct_warp_bytes<GeneT, VSTAGE>(vcap);
    if (smem > 200 * 1024) return SLAFB_OK;         // very wide rows: the CTA kernel / general kernel take them
    auto kern = tokenize_pct_warp_kernel<GeneT, VSTAGE>;
    static std::atomic<int> per_sm_cached[64];
    static std::atomic<size_t> smem_cached[64];
    int per_sm = per_sm_cached[ctx->device & 63].load(std::memory_order_acquire);
    if (per_sm == 0 || smem_cached[ctx->device & 63].load() != smem) {
        CU(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPctWarps * 32, smem));
        if (per_sm < 1) per_sm = 1;
        smem_cached[ctx->device & 63].store(smem);
        per_sm_cached[ctx->device & 63].store(per_sm, std::memory_order_release);
    }
    uint32_t grid = (uint32_t)(ctx->n_sm * per_sm);
    const uint32_t need = (a.n_cells + kPctWarps - 1) / kPctWarps;
    if (need < grid) grid = need;
    if (grid == 0) grid = 1;
    kern<<<grid, kPctWarps * 32, smem, st>>>(fa);
    CU(ctx, cudaGetLastError());
    ctx->acc.kernel_launches++;
    *launched = true;
    return SLAFB_OK;
}

template <typename GeneT>
int launch_pct_warp(slafb_ctx *ctx, const TokArgs &a, cudaStream_t st, bool *launched) {
    *launched = false;
    static const bool vstage = [] { const char *e = getenv("SLAFB_PCT_VSTAGE"); return e ? atoi(e) != 0 : true; }();
    static const uint32_t vcap_env = [] { const char *e = getenv("SLAFB_PCT_VCAP"); return e ? (uint32_t)atoi(e) : 0u; }();
    const uint32_t limit = a.L - 1u;
    uint32_t vcap = vcap_env ? vcap_env : 3072u;
    if (limit + 24u > vcap) vcap = limit + 24u;     // the stage holds every gene that can be emitted ...
    vcap = (vcap + 255u) & ~255u;                   // ... and a whole number of 256-row steps of the warp
    return vstage ? launch_pct_warp_v<GeneT, true>(ctx, a, st, vcap, launched) : launch_pct_warp_v<GeneT, false>(ctx, a, st, vcap, launched);
}

// float32 values, plain Geneformer and both scGPT modes: one warp per cell (kernels.cuh, K2f).  SLAFB_F32_CTA=1 keeps the
// CTA-per-cell kernel (A/B runs); SLAFB_F32_VCAP=<rows> sets the value stage.
template <typename GeneT, int MODE>
int launch_f32_warp(slafb_ctx *ctx, const TokArgs &a, cudaStream_t st, bool *launched) {
    *launched = false;
    static const bool use_cta = [] { const char *e = getenv("SLAFB_F32_CTA"); return e && atoi(e) != 0; }();
    static const uint32_t vcap_env = [] { const char *e = getenv("SLAFB_F32_VCAP"); return e ? (uint32_t)atoi(e) : 0u; }();
    if (use_cta || ctx->f32_warp_off) return SLAFB_OK;
    constexpr bool SCGPT = (MODE == MODE_SCGPT_RAW || MODE == MODE_SCGPT_BINNED);
    const uint32_t limit = SCGPT ? a.max_genes : (a.L - 1u);
    FastArgs fa;
    memset(&fa, 0, sizeof fa);
    fa.t = a;
    fa.gcap = (limit + 15u + 7u) & ~7u;             // the genes of every row that can be emitted
    uint32_t vcap = vcap_env ? vcap_env : 2304u;
    if (fa.gcap > vcap) vcap = fa.gcap;             // the emitted rows' values are staged as well ...
    fa.vcap = (vcap + 255u) & ~255u;                // ... in a whole number of 256-row steps of the warp
    if (MODE == MODE_GF_PCT) fa.gcap = fa.vcap;     // percentile mode: any row may survive the filter, the gene stage is as long as the value stage
    fa.ticket = a.work_counters + 2;
    const size_t smem = (size_t)kPctWarps * f32_warp_bytes<GeneT>(fa.vcap, fa.gcap);
    if (smem > 200 * 1024) return SLAFB_OK;         // very wide rows: the CTA kernel / general kernel take them
    auto kern = tokenize_f32_warp_kernel<GeneT, MODE>;
    static std::atomic<int> per_sm_cached[64];
    static std::atomic<size_t> smem_cached[64];
    int per_sm = per_sm_cached[ctx->device & 63].load(std::memory_order_acquire);
    if (per_sm == 0 || smem_cached[ctx->device & 63].load() != smem) {
        CU(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPctWarps * 32, smem));
        if (per_sm < 1) per_sm = 1;
        smem_cached[ctx->device & 63].store(smem);
        per_sm_cached[ctx->device & 63].store(per_sm, std::memory_order_release);
    }
    uint32_t grid = (uint32_t)(ctx->n_sm * per_sm);
    const uint32_t need = (a.n_cells + kPctWarps - 1) / kPctWarps;
    if (need < grid) grid = need;
    if (grid == 0) grid = 1;
    kern<<<grid, kPctWarps * 32, smem, st>>>(fa);
    CU(ctx, cudaGetLastError());
    ctx->acc.kernel_launches++;
    ctx->last_fast_f32_warp = true;
    *launched = true;
    return SLAFB_OK;
}

template <typename GeneT, typename ValT>
int dispatch_fast(slafb_ctx *ctx, int mode, const TokArgs &a, cudaStream_t st, bool *launched) {
    ctx->last_fast_f32_warp = false;
    if constexpr (sizeof(ValT) == 4) {
        int rc = SLAFB_OK;
        *launched = false;
        if (mode == MODE_GF) rc = launch_f32_warp<GeneT, MODE_GF>(ctx, a, st, launched);
        else if (mode == MODE_SCGPT_RAW) rc = launch_f32_warp<GeneT, MODE_SCGPT_RAW>(ctx, a, st, launched);
        else if (mode == MODE_SCGPT_BINNED) rc = launch_f32_warp<GeneT, MODE_SCGPT_BINNED>(ctx, a, st, launched);
        else if (mode == MODE_GF_PCT) rc = launch_f32_warp<GeneT, MODE_GF_PCT>(ctx, a, st, launched);
        if (rc || *launched) return rc;
    }
    switch (mode) {
    case MODE_GF: return launch_fast<GeneT, ValT, MODE_GF>(ctx, a, st, launched);
    case MODE_SCGPT_RAW: return launch_fast<GeneT, ValT, MODE_SCGPT_RAW>(ctx, a, st, launched);
    case MODE_SCGPT_BINNED: return launch_fast<GeneT, ValT, MODE_SCGPT_BINNED>(ctx, a, st, launched);
    case MODE_GF_PCT:
        // uint16 values: one warp per cell (K2w), or the CTA-per-cell variant with SLAFB_PCT_CTA=1 (A/B); float32 values need a sort
        // per cell (general kernel)
        if constexpr (sizeof(ValT) == 2) {
            static const bool use_cta = [] { const char *e = getenv("SLAFB_PCT_CTA"); return e && atoi(e) != 0; }();
            if (!use_cta) {
                int rc = launch_pct_warp<GeneT>(ctx, a, st, launched);
                if (rc || *launched) return rc;
            }
            return launch_fast<GeneT, ValT, MODE_GF_PCT>(ctx, a, st, launched);
        } else { *launched = false; return SLAFB_OK; }
    default: *launched = false; return SLAFB_OK;
    }
}
template <typename GeneT>
int dispatch_fast_packed(slafb_ctx *ctx, int mode, const TokArgs &a, cudaStream_t st, bool *launched, const uint8_t *packed, const uint32_t *blk_off) {
    switch (mode) {
    case MODE_GF: return launch_fast<GeneT, uint16_t, MODE_GF, true>(ctx, a, st, launched, packed, blk_off);
    case MODE_GF_PCT: return launch_fast<GeneT, uint16_t, MODE_GF_PCT, true>(ctx, a, st, launched, packed, blk_off);
    case MODE_SCGPT_RAW: return launch_fast<GeneT, uint16_t, MODE_SCGPT_RAW, true>(ctx, a, st, launched, packed, blk_off);
    default: return launch_fast<GeneT, uint16_t, MODE_SCGPT_BINNED, true>(ctx, a, st, launched, packed, blk_off);
    }
}
template <typename GeneT, typename ValT, int SINK>
int dispatch_general(slafb_ctx *ctx, int mode, const TokArgs &a, const WindowSink &w, cudaStream_t st) {
    switch (mode) {
    case MODE_GF: return launch_general<GeneT, ValT, MODE_GF, SINK>(ctx, a, w, st);
    case MODE_GF_PCT: return launch_general<GeneT, ValT, MODE_GF_PCT, SINK>(ctx, a, w, st);
    case MODE_SCGPT_RAW: return launch_general<GeneT, ValT, MODE_SCGPT_RAW, SINK>(ctx, a, w, st);
    default: return launch_general<GeneT, ValT, MODE_SCGPT_BINNED, SINK>(ctx, a, w, st);
    }
}

#define DISPATCH_TYPES(gene_dt, val_dt, CALL)                                                    \
    do {                                                                                         \
        if ((gene_dt) == SLAFB_U16) {                                                            \
            if ((val_dt) == SLAFB_U16) { using GeneT = uint16_t; using ValT = uint16_t; CALL; } \
            else { using GeneT = uint16_t; using ValT = float; CALL; }                           \
        } else {                                                                                 \
            if ((val_dt) == SLAFB_U16) { using GeneT = int32_t; using ValT = uint16_t; CALL; }  \
            else { using GeneT = int32_t; using ValT = float; CALL; }                            \
        }                                                                                        \
    } while (0)

void fill_tok_args(slafb_ctx *ctx, Slot &s, const slafb_params *p, TokArgs &a, uint32_t n_cells, bool use_perm) {
    memset(&a, 0, sizeof a);
    a.gene = s.gene;
    a.value = s.value;
    a.n_rows = (uint32_t)s.n_rows;
    a.seg_start = s.seg_start;
    a.seg_cell = s.seg_cell;
    a.perm = use_perm ? s.perm_dev : nullptr;
    a.n_cells = n_cells;
    a.cell_signed = (s.cell_dtype == SLAFB_I32);
    a.max_items = (uint32_t)p->max_items;
    a.max_genes = (uint32_t)p->max_genes;
    a.L = (p->mode >= SLAFB_SCGPT_RAW) ? (uint32_t)p->max_genes + 2u : (uint32_t)p->max_genes;
    a.n_bins = p->n_bins;
    a.w_bins = p->window_n_bins > 0 ? p->window_n_bins : p->n_bins;
    a.vocab = p->vocab_size;
    a.min_pct = p->min_percentile;
    a.bin_size = (float)(1.0 / (double)(p->n_bins > 0 ? p->n_bins : 1));  // tokenizers.py:508 then float32 at :531
    a.log1p_lut = ctx->log1p_lut;
    a.work_list = s.work_list;
    a.work_counters = s.counters + 2;
    a.sort_scratch = ctx->sort_scratch;
    a.all_cells = 0;
    a.host_slow = nullptr;
}

// The slot for the next batch: a free one whose previous tokenization has already finished (reusing it costs no wait) and,
// among those, one that already owns staging columns (no allocation).  When every slot with staging columns is still busy on the
// GPU and at least three of them exist, the one released longest ago is taken and submit waits for it: the host is ahead of the
// copy engine then anyway, and a cudaMalloc of another slot's columns in the middle of a run costs 20-30 ms (it waits for the
// queued copies to drain).  Otherwise a free slot without columns, or the free slot released longest ago.  A feeder that keeps
// one or two batches in flight therefore touches three slots (and their 12 B/row of staging); only a pipeline that holds more
// uncollected batches than that gets more.
int pick_slot(slafb_ctx *ctx, bool needs_staging = false) {
    int ready_alloc = -1, ready = -1, oldest = -1, oldest_alloc = -1, n_alloc = 0;
    for (int i = 0; i < kSlots; i++) {
        Slot &s = ctx->slot[i];
        if (s.in_cell) n_alloc++;
        if (s.in_flight) continue;
        if (oldest < 0 || s.freed_at < ctx->slot[oldest].freed_at) oldest = i;
        if (s.in_cell && (oldest_alloc < 0 || s.freed_at < ctx->slot[oldest_alloc].freed_at)) oldest_alloc = i;
        bool done = true;
        if (s.ev_done && cudaEventQuery(s.ev_done) != cudaSuccess) { done = false; cudaGetLastError(); }
        if (!done) continue;
        if (s.in_cell) { if (ready_alloc < 0 || s.freed_at < ctx->slot[ready_alloc].freed_at) ready_alloc = i; }
        else if (ready < 0 || s.freed_at < ctx->slot[ready].freed_at) ready = i;
    }
    if (ready_alloc >= 0) return ready_alloc;
    if (needs_staging && n_alloc >= 3 && oldest_alloc >= 0) return oldest_alloc;
    return ready >= 0 ? ready : oldest;
}

// Small host pieces in pageable memory (the host feeder's partial cells: a few thousand rows) are copied into the slot's
// pinned scratch first.  A pageable cudaMemcpyAsync makes the calling thread wait until the copy stream has drained, which
// would tie the feeder to the copy engine exactly when it tries to run ahead of it; big pageable pieces are left to the
// driver.  `src` receives the pointers to copy from, per piece: {cell, gene, value}.
struct PieceSrc { const void *col[3]; };
constexpr size_t kSmallPieceRows = 1u << 16;      // per piece
constexpr size_t kSmallTotalRows = 1u << 21;      // per batch (8 MiB per column region at most)

bool is_pageable_host(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

int stage_small_pieces(slafb_ctx *ctx, Slot &s, const slafb_coo *pieces, int n_pieces, bool with_cell, std::vector<PieceSrc> &src) {
    src.resize((size_t)n_pieces);
    std::vector<char> take((size_t)n_pieces, 0);
    size_t rows = 0;
    for (int i = 0; i < n_pieces; i++) {
        const slafb_coo &p = pieces[i];
        src[i] = PieceSrc{{p.cell, p.gene, p.value}};
        const size_t n = (size_t)p.n_rows;
        if (p.location != SLAFB_LOC_HOST || n == 0 || n > kSmallPieceRows || rows + n > kSmallTotalRows) continue;
        if (!(is_pageable_host(p.gene) || is_pageable_host(p.value) || (with_cell && is_pageable_host(p.cell)))) continue;
        take[i] = 1;
        rows += n;
    }
    if (!rows) return SLAFB_OK;
    if (rows > s.small_rows) {
        if (s.h_small) CU(ctx, cudaFreeHost(s.h_small));
        s.h_small = nullptr; s.small_rows = 0;
        size_t cap = 1u << 18;       // 3 MiB: the pending cells of a 16-scanner feeder fit without ever growing
        while (cap < rows) cap <<= 1;
        CU(ctx, cudaHostAlloc(&s.h_small, cap * 12, cudaHostAllocPortable));
        s.small_rows = cap;
    }
    const size_t width[3] = {4, dtype_size(pieces[0].gene_dtype), dtype_size(pieces[0].value_dtype)};
    size_t off = 0;
    for (int i = 0; i < n_pieces; i++) {
        if (!take[i]) continue;
        const size_t n = (size_t)pieces[i].n_rows;
        for (int c = with_cell ? 0 : 1; c < 3; c++) {
            char *dst = s.h_small + (size_t)c * s.small_rows * 4 + off * width[c];
            memcpy(dst, src[i].col[c], n * width[c]);
            src[i].col[c] = dst;
        }
        off += n;
    }
    return SLAFB_OK;
}

// front half: (H2D) + K1 + count read-back, on the side stream.  The batch is the concatenation of
// `n_pieces` row blocks (the host feeder's partial cells + fragment chunks, datasets.py:880-888):
// host pieces are copied back to back into the slot's staging columns straight from where they
// live (pinned / registered memory = true async DMA, no host-side assembly); a single device piece
// is used in place.
int front(slafb_ctx *ctx, Slot &s, const slafb_coo *pieces, int n_pieces, cudaStream_t caller) {
    if (!pieces || n_pieces < 1) return fail(ctx, SLAFB_ERR_INVALID, "no input pieces");
    int64_t total = 0;
    for (int i = 0; i < n_pieces; i++) {
        int rc = check_coo(ctx, &pieces[i]);
        if (rc) return rc;
        if (pieces[i].cell_dtype != pieces[0].cell_dtype || pieces[i].gene_dtype != pieces[0].gene_dtype ||
            pieces[i].value_dtype != pieces[0].value_dtype || pieces[i].location != pieces[0].location)
            return fail(ctx, SLAFB_ERR_INVALID, "all pieces of a batch must share column dtypes and location");
        total += pieces[i].n_rows;
    }
    if (total > ctx->max_rows) return fail(ctx, SLAFB_ERR_CAPACITY, "n_rows %lld exceeds context max_rows %lld", (long long)total, (long long)ctx->max_rows);
    const slafb_coo *in = &pieces[0];
    // The side stream exists to overlap H2D copies with tokenization.  A device-resident batch has no
    // copy to hide; SLAFB_OPT_FRONT_ON_CALLER keeps its CSR build on the caller's stream, so every kernel
    // runs alone and its CUDA-event time is that kernel's own (bench.py's isolated-kernel pass).
    cudaStream_t sf = (ctx->front_on_caller && in->location == SLAFB_LOC_DEVICE && n_pieces == 1) ? caller : ctx->s_in;
    if (s.ev_done) {
        // the slot's scratch may still be read by the previous batch's tokenization
        CU(ctx, cudaEventSynchronize(s.ev_done));
        harvest(ctx, s);
    }
    s.sample = ctx->timing_every > 0 && ((ctx->timing_every == 1) || (ctx->batch_seq % (uint64_t)ctx->timing_every == 0));
    ctx->batch_seq++;
    s.host_segs = false;
    s.packed = false;
    s.n_rows = total;
    s.cell_dtype = in->cell_dtype;
    s.gene_dtype = in->gene_dtype;
    s.value_dtype = in->value_dtype;
    s.h_counters[0] = 0;
    s.h_counters[1] = 0;
    if (total == 0) {
        CU(ctx, cudaEventRecord(s.ev_front, sf));
        CU(ctx, cudaEventRecord(s.ev_loaded, sf));
        return SLAFB_OK;
    }
    CU(ctx, cudaEventRecord(s.ev_in, caller));
    CU(ctx, cudaStreamWaitEvent(sf, s.ev_in, 0));
    bool staged = false;
    std::vector<PieceSrc> src;
    if (in->location == SLAFB_LOC_HOST || n_pieces > 1) {
        int rc = ensure_staging(ctx, s);
        if (rc) return rc;
        rc = stage_small_pieces(ctx, s, pieces, n_pieces, true, src);
        if (rc) return rc;
        const cudaMemcpyKind kind = in->location == SLAFB_LOC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
        if (s.sample) CU(ctx, cudaEventRecord(s.t_h2d0, sf));
        // The cell column goes first and the CSR build runs right behind it: the segment table (all the host
        // needs to decide what the NEXT batch looks like) is ready after half of the bytes, while the gene and
        // value columns are still in flight.  Column by column, so each staging column is written front to back.
        size_t off = 0;
        for (int i = 0; i < n_pieces; i++) {
            const size_t n = (size_t)pieces[i].n_rows;
            if (n) CU(ctx, cudaMemcpyAsync(static_cast<char *>(s.in_cell) + off * 4, src[i].col[0], n * 4, kind, sf));
            off += n;
        }
        staged = true;
        s.cell = s.in_cell; s.gene = s.in_gene; s.value = s.in_value;
    } else {
        s.cell = in->cell; s.gene = in->gene; s.value = in->value;
    }
    CsrArgs k{};
    k.cell = static_cast<const uint32_t *>(s.cell);
    k.n_rows = (uint32_t)total;
    k.max_cells = (uint32_t)ctx->max_cells;
    k.seg_start = s.seg_start;
    k.seg_cell = s.seg_cell;
    k.tile_state = ctx->tile_state;
    k.ticket = ctx->ticket;
    k.ticket_base = ctx->ticket_base;
    ctx->epoch = (ctx->epoch + 1u) & 0x3fffffffu;
    if (ctx->epoch == 0) {  // 2^30 launches: restart the epoch space
        CU(ctx, cudaMemsetAsync(ctx->tile_state, 0, (size_t)ctx->n_sm * 8 * sizeof(unsigned long long), sf));
        ctx->epoch = 1;
    }
    k.epoch = ctx->epoch;
    k.n_cells_out = s.counters;
    k.host_counts = s.h_counters_dev;
    k.work_counters = s.counters + 2;
    // one contiguous span of rows per CTA (a whole number of 2048-row stages).  Two variants of the scan: rows through the
    // bulk-copy engine (4 CTAs of 288 threads per SM) or through direct 128-bit loads (8 CTAs of 256 threads per SM)
    static const bool use_ldg = [] { const char *e = getenv("SLAFB_K1_LDG"); return e ? atoi(e) != 0 : SLAFB_K1_LDG_DEFAULT; }();
    uint32_t per_sm = use_ldg ? 8 : 4;
    if (const char *env = getenv("SLAFB_CSR_CTAS_PER_SM")) {   // tuning knob
        const long v = atol(env);
        if (v >= 1 && v <= 8) per_sm = (uint32_t)v;
    }
    uint32_t grid = (uint32_t)ctx->n_sm * per_sm;
    const uint64_t tiles = ((uint64_t)total + kCsrStageRows - 1) / kCsrStageRows;
    if (tiles < grid) grid = (uint32_t)tiles;
    k.rows_per_cta = (uint32_t)(((tiles + grid - 1) / grid) * kCsrStageRows);
    grid = (uint32_t)(((uint64_t)total + k.rows_per_cta - 1) / k.rows_per_cta);   // no empty spans
    // every CSR build of this context draws spans from one ticket and publishes into one tile-state array: two of them must
    // never run at once.  On one stream they are ordered anyway; when the stream changes, wait for the previous build.
    if (ctx->k1_any && ctx->k1_stream != sf) CU(ctx, cudaStreamWaitEvent(sf, ctx->ev_k1, 0));
    if (s.sample) CU(ctx, cudaEventRecord(s.t_k1a, sf));
    if (use_ldg) csr_build_ldg_kernel<<<grid, kCsrLdgBlock, 0, sf>>>(k);
    else csr_build_kernel<<<grid, kCsrBlock, 0, sf>>>(k);
    CU(ctx, cudaGetLastError());
    if (s.sample) { CU(ctx, cudaEventRecord(s.t_k1b, sf)); s.timed_front = true; }
    if (ctx->front_on_caller || ctx->k1_stream != sf) {          // (the default configuration keeps every build on s_in: no event needed)
        CU(ctx, cudaEventRecord(ctx->ev_k1, sf));
        ctx->k1_stream = sf;
        ctx->k1_any = true;
    }
    ctx->ticket_base += grid;
    ctx->acc.kernel_launches++;
    CU(ctx, cudaEventRecord(s.ev_front, sf));
    if (staged) {
        const cudaMemcpyKind kind = in->location == SLAFB_LOC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
        const size_t gsz = dtype_size(in->gene_dtype), vsz = dtype_size(in->value_dtype);
        size_t off = 0;
        for (int i = 0; i < n_pieces; i++) {
            const size_t n = (size_t)pieces[i].n_rows;
            if (n) CU(ctx, cudaMemcpyAsync(static_cast<char *>(s.in_gene) + off * gsz, src[i].col[1], n * gsz, kind, sf));
            off += n;
        }
        off = 0;
        for (int i = 0; i < n_pieces; i++) {
            const size_t n = (size_t)pieces[i].n_rows;
            if (n) CU(ctx, cudaMemcpyAsync(static_cast<char *>(s.in_value) + off * vsz, src[i].col[2], n * vsz, kind, sf));
            off += n;
        }
        if (s.sample) { CU(ctx, cudaEventRecord(s.t_h2d1, sf)); s.timed_h2d = true; }
    }
    CU(ctx, cudaEventRecord(s.ev_loaded, sf));
    return SLAFB_OK;
}

void shuffle_worker(slafb_ctx *ctx) {
    cudaSetDevice(ctx->device);
    for (;;) {
        int i;
        {
            std::unique_lock<std::mutex> lk(ctx->wm);
            ctx->wcv.wait(lk, [&] { return ctx->wstop || !ctx->wjobs.empty(); });
            if (ctx->wstop) return;
            i = ctx->wjobs.front();
            ctx->wjobs.pop_front();
        }
        Slot &s = ctx->slot[i];
        int64_t C = -1;
        s.perm_uploaded = false;
        // wait for the cell count without occupying a core: the helpers have most of a pipeline depth of slack, and on an
        // 8-GPU host (4 vCPUs per rank on this pool) a spinning cudaEventSynchronize per helper competes with the feeder threads
        cudaError_t ready = cudaSuccess;
        while (!s.host_segs && (ready = cudaEventQuery(s.ev_front)) == cudaErrorNotReady) {   // host-supplied segments: the count is already there
            cudaGetLastError();
            std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
        if (ready == cudaSuccess && !s.h_counters[1]) {
            C = s.h_counters[0];
            py_shuffle_perm(s.prep_seed, C, s.h_perm, s.h_draws);
            // upload it right away on its own stream with a tiny kernel (not the copy engine, which is busy with whole
            // batches of columns): collect() then only waits for an event that is long complete
            if (C > 0) {
                perm_upload_kernel<<<(unsigned)((C + 255) / 256), 256, 0, ctx->s_perm>>>(s.h_perm_dev, s.perm_dev, (uint32_t)C);
                if (cudaGetLastError() == cudaSuccess && cudaEventRecord(s.ev_perm, ctx->s_perm) == cudaSuccess) s.perm_uploaded = true;
            }
        }
        {
            std::lock_guard<std::mutex> lk(ctx->wm);
            s.prep_cells = C;
            s.prep_state.store(2);
        }
        ctx->wdone.notify_all();
    }
}

// the helper may still be writing the slot's permutation buffers: wait for it (and drop its result)
void settle_prepared(slafb_ctx *ctx, Slot &s) {
    if (s.prep_state.load() == 0) return;
    std::unique_lock<std::mutex> lk(ctx->wm);
    ctx->wdone.wait(lk, [&] { return s.prep_state.load() == 2; });
}

// waits for the cell count; builds + uploads the permutation.  Returns n_cells through *n.
int middle(slafb_ctx *ctx, Slot &s, const slafb_params *p, int64_t *n, cudaStream_t st, bool *use_perm) {
    if (!s.host_segs) {               // (a host-supplied segment table: the count is known, nothing to wait for)
        const double tw0 = now_ms();
        CU(ctx, cudaEventSynchronize(s.ev_front));
        ctx->acc.host_wait_ms += now_ms() - tw0;
    }
    if (s.h_counters[1]) {
        *n = s.h_counters[0];
        return fail(ctx, SLAFB_ERR_CAPACITY, "batch has %u cells, context max_cells is %lld", s.h_counters[0], (long long)ctx->max_cells);
    }
    const int64_t C = s.h_counters[0];
    *n = C;
    *use_perm = false;
    if (C == 0) return SLAFB_OK;
    if (p->flags & SLAFB_FLAG_CHECK_CONTIGUOUS) {
        size_t slots = 1024;
        while (slots < (size_t)C * 4) slots <<= 1;
        if (slots > ctx->dup_slots) {
            if (ctx->dup_table) CU(ctx, cudaFree(ctx->dup_table));
            ctx->dup_table = nullptr; ctx->dup_slots = 0;
            CU(ctx, cudaMalloc(&ctx->dup_table, slots * sizeof(unsigned long long)));
            ctx->dup_slots = slots;
        }
        CU(ctx, cudaStreamWaitEvent(st, s.ev_front, 0));
        CU(ctx, cudaMemsetAsync(ctx->dup_table, 0, slots * sizeof(unsigned long long), st));
        CU(ctx, cudaMemsetAsync(s.counters + 6, 0, sizeof(uint32_t), st));
        dup_check_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(s.seg_cell, (uint32_t)C, ctx->dup_table, (uint32_t)(slots - 1), s.counters + 6);
        CU(ctx, cudaGetLastError());
        ctx->acc.kernel_launches++;
        CU(ctx, cudaMemcpyAsync(s.h_counters + 3, s.counters + 6, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaStreamSynchronize(st));
        if (s.h_counters[3])
            return fail(ctx, SLAFB_ERR_NONCONTIGUOUS, "a cell id occurs in two separated row runs; splice partial cells on the host first");
    }
    settle_prepared(ctx, s);
    const bool prepared = s.prep_state.load() == 2 && p->shuffle == SLAFB_SHUFFLE_SEED && s.prep_seed == p->seed && s.prep_cells == C;
    s.prep_state.store(0);
    if (s.perm_uploaded && !prepared) {
        // a permutation was prepared (and is being uploaded) for another seed / order: let that copy finish before the
        // permutation buffers are rewritten
        CU(ctx, cudaEventSynchronize(s.ev_perm));
        s.perm_uploaded = false;
    }
    if (p->shuffle == SLAFB_SHUFFLE_SEED) {
        if (!prepared) {
            const double ts0 = now_ms();
            py_shuffle_perm(p->seed, C, s.h_perm, s.h_draws);
            ctx->acc.host_shuffle_ms += now_ms() - ts0;
        }
        *use_perm = true;
    } else if (p->shuffle == SLAFB_SHUFFLE_PERM) {
        // a permutation of all segments, or a SELECTION: fewer entries than segments emit only those
        // cells (the host feeder drops incomplete boundary cells this way, datasets.py:891-928)
        if (p->n_perm < 0 || p->n_perm > C || (p->n_perm > 0 && !p->perm))
            return fail(ctx, SLAFB_ERR_INVALID, "perm length %lld exceeds the number of cells %lld", (long long)p->n_perm, (long long)C);
        for (int64_t i = 0; i < p->n_perm; i++) {
            if (p->perm[i] < 0 || p->perm[i] >= C) return fail(ctx, SLAFB_ERR_INVALID, "perm[%lld] out of range", (long long)i);
            s.h_perm[i] = p->perm[i];
        }
        *use_perm = true;
        *n = p->n_perm;
    }
    CU(ctx, cudaStreamWaitEvent(st, s.ev_loaded, 0));      // gene / value columns (ev_front covered the cell column only)
    if (prepared && s.perm_uploaded) {
        CU(ctx, cudaStreamWaitEvent(st, s.ev_perm, 0));    // the helper already uploaded this permutation
        s.perm_uploaded = false;
        return SLAFB_OK;
    }
    if (*use_perm && *n > 0) {
        perm_upload_kernel<<<(unsigned)((*n + 255) / 256), 256, 0, st>>>(s.h_perm_dev, s.perm_dev, (uint32_t)*n);
        CU(ctx, cudaGetLastError());
        ctx->acc.kernel_launches++;
    }
    return SLAFB_OK;
}

int back(slafb_ctx *ctx, Slot &s, const slafb_params *p, const slafb_out *out, int64_t *n_cells, cudaStream_t st) {
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!out || !n_cells) return fail(ctx, SLAFB_ERR_INVALID, "out / n_cells is NULL");
    bool use_perm = false;
    rc = middle(ctx, s, p, n_cells, st, &use_perm);
    if (rc) return rc;
    const int64_t C = *n_cells;
    if (C == 0) { CU(ctx, cudaEventRecord(s.ev_done, st)); return SLAFB_OK; }
    if (C > out->capacity_cells) return fail(ctx, SLAFB_ERR_CAPACITY, "batch has %lld cells, output capacity is %lld", (long long)C, (long long)out->capacity_cells);
    const bool scgpt = p->mode >= SLAFB_SCGPT_RAW;
    if (!out->input_ids || !out->attention_mask || !out->cell_ids || (scgpt && !out->values))
        return fail(ctx, SLAFB_ERR_INVALID, "NULL output tensor");
    if (((uintptr_t)out->input_ids | (uintptr_t)out->attention_mask | (scgpt ? (uintptr_t)out->values : 0)) & 15u)
        return fail(ctx, SLAFB_ERR_INVALID, "output tensors must be 16-byte aligned");
    if (s.value_dtype == SLAFB_F32 || p->mode == SLAFB_GENEFORMER_PCT) {
        rc = ensure_sort_scratch(ctx);
        if (rc) return rc;
    }
    TokArgs a;
    fill_tok_args(ctx, s, p, a, (uint32_t)C, use_perm);
    a.ids = reinterpret_cast<long long *>(out->input_ids);
    a.mask = out->attention_mask;
    a.vals = reinterpret_cast<long long *>(out->values);
    a.cell_ids = reinterpret_cast<long long *>(out->cell_ids);
    a.host_slow = s.h_counters_dev + 2;
    WindowSink none{};
    if (s.packed && p->mode == SLAFB_GENEFORMER && (p->order == SLAFB_ORDER_RANK || p->norm))
        return fail(ctx, SLAFB_ERR_INVALID, "the rank-value extension takes unpacked rows");
    if (p->mode == SLAFB_GENEFORMER && (p->order == SLAFB_ORDER_RANK || p->norm)) {
        // extension path: every cell through the rank-value kernel
        if (!ctx->rank_scratch) CU(ctx, cudaMalloc(&ctx->rank_scratch, ((size_t)ctx->max_rows + 64) * sizeof(unsigned long long)));
        RankArgs ra;
        ra.t = a;
        ra.norm = p->norm;
        ra.n_norm = (uint32_t)(p->n_norm > 0xffffffffll ? 0xffffffffll : p->n_norm);
        ra.rank_order = p->order == SLAFB_ORDER_RANK;
        ra.scratch64 = ctx->rank_scratch;
        ra.from_list = 0;
        CU(ctx, cudaMemsetAsync(s.counters + 2, 0, 4 * sizeof(uint32_t), st));     // work-list count, head (sorting kernel), row ticket, head (second bucket launch)
        if (s.sample) CU(ctx, cudaEventRecord(s.t_k2a, st));
        DISPATCH_TYPES(s.gene_dtype, s.value_dtype, rc = (launch_rank<GeneT, ValT>(ctx, ra, st)));
        if (rc) return rc;
        if (s.sample) { CU(ctx, cudaEventRecord(s.t_k2b, st)); CU(ctx, cudaEventRecord(s.t_k2c, st)); }
        s.h_counters[2] = 0;
        s.all_general = false;
        s.timed_back = s.sample;
        s.ran_back = true;
        CU(ctx, cudaEventRecord(s.ev_done, st));
        ctx->acc.batches++;
        ctx->acc.cells += C;
        ctx->acc.rows += s.n_rows;
        return SLAFB_OK;
    }
    if (s.sample) CU(ctx, cudaEventRecord(s.t_k2a, st));
    bool fast_launched = false;
    if (s.packed) {
        // packed rows: the tokenizing kernel decodes every cell's block in its shared-memory stage; the cells it defers are
        // unpacked into the slot's scratch columns (a.gene / a.value) for the general kernel
        if (s.gene_dtype == SLAFB_U16) rc = dispatch_fast_packed<uint16_t>(ctx, p->mode, a, st, &fast_launched, s.pk_bytes, s.blk_off);
        else rc = dispatch_fast_packed<int32_t>(ctx, p->mode, a, st, &fast_launched, s.pk_bytes, s.blk_off);
        if (rc) return rc;
        if (!fast_launched) {                        // (rows too wide for the fast kernel) every cell is unpacked: all of them are "deferred"
            return fail(ctx, SLAFB_ERR_INVALID, "max_genes too large for the packed route");
        }
        const unsigned ugrid = (unsigned)ctx->n_sm * 2u;
        if (s.gene_dtype == SLAFB_U16)
            unpack_deferred_kernel<uint16_t><<<ugrid, kTokThreads, 0, st>>>(s.pk_bytes, s.blk_off, s.seg_start, a.perm, s.work_list, s.counters + 2,
                                                                           static_cast<uint16_t *>(s.in_cell), static_cast<uint16_t *>(s.in_value));
        else
            unpack_deferred_kernel<int32_t><<<ugrid, kTokThreads, 0, st>>>(s.pk_bytes, s.blk_off, s.seg_start, a.perm, s.work_list, s.counters + 2,
                                                                          static_cast<int32_t *>(s.in_cell), static_cast<uint16_t *>(s.in_value));
        CU(ctx, cudaGetLastError());
        ctx->acc.kernel_launches++;
    } else {
        DISPATCH_TYPES(s.gene_dtype, s.value_dtype, rc = (dispatch_fast<GeneT, ValT>(ctx, p->mode, a, st, &fast_launched)));
        if (rc) return rc;
        s.used_f32_warp = ctx->last_fast_f32_warp;
        s.back_cells = C;
    }
    if (!fast_launched) a.all_cells = 1;
    if (s.sample) CU(ctx, cudaEventRecord(s.t_k2b, st));
    DISPATCH_TYPES(s.gene_dtype, s.value_dtype, rc = (dispatch_general<GeneT, ValT, 0>(ctx, p->mode, a, none, st)));
    if (rc) return rc;
    if (s.sample) CU(ctx, cudaEventRecord(s.t_k2c, st));
    s.all_general = !fast_launched;
    s.timed_back = s.sample;
    s.ran_back = true;
    CU(ctx, cudaEventRecord(s.ev_done, st));
    ctx->acc.batches++;
    ctx->acc.cells += C;
    ctx->acc.rows += s.n_rows;
    return SLAFB_OK;
}

// submit() when the segment table is known on the host: validates the pieces, takes a slot and leaves the table's pinned
// block (s.h_seg: [C + 1] starts, then [C] ids) to be filled by `fill`, then enqueues table + gene / value columns.
template <typename Fill>
int submit_host_segments(slafb_ctx *ctx, const slafb_coo *pieces, int32_t n_pieces, int64_t n_cells_max, void *stream, int32_t *slot,
                         int64_t *n_cells_out, Fill fill) {
    if (!ctx || !slot || !pieces || n_pieces < 1) return fail(ctx, SLAFB_ERR_INVALID, "bad ctx / slot / pieces");
    CU(ctx, cudaSetDevice(ctx->device));
    const int si = pick_slot(ctx, pieces[0].location == SLAFB_LOC_HOST || n_pieces > 1);
    if (si < 0) return fail(ctx, SLAFB_ERR_BUSY, "all %d pipeline slots hold batches that were not collected yet", kSlots);
    Slot &s = ctx->slot[si];
    const double t0 = now_ms();
    int64_t total = 0;
    for (int i = 0; i < n_pieces; i++) {
        const slafb_coo &p = pieces[i];
        if (p.n_rows < 0 || (p.n_rows > 0 && (!p.gene || !p.value))) return fail(ctx, SLAFB_ERR_INVALID, "NULL gene / value column");
        if (p.gene_dtype != pieces[0].gene_dtype || p.value_dtype != pieces[0].value_dtype || p.location != pieces[0].location ||
            p.cell_dtype != pieces[0].cell_dtype)
            return fail(ctx, SLAFB_ERR_INVALID, "all pieces of a batch must share column dtypes and location");
        total += p.n_rows;
    }
    const slafb_coo *in = &pieces[0];
    if (in->gene_dtype != SLAFB_U16 && in->gene_dtype != SLAFB_I32 && in->gene_dtype != SLAFB_U32) return fail(ctx, SLAFB_ERR_INVALID, "gene dtype must be uint16, int32 or uint32");
    if (in->value_dtype != SLAFB_U16 && in->value_dtype != SLAFB_F32) return fail(ctx, SLAFB_ERR_INVALID, "value dtype must be uint16 or float32");
    if (in->cell_dtype != SLAFB_U32 && in->cell_dtype != SLAFB_I32) return fail(ctx, SLAFB_ERR_INVALID, "cell dtype must be uint32 or int32");
    if (total > ctx->max_rows) return fail(ctx, SLAFB_ERR_CAPACITY, "n_rows %lld exceeds context max_rows %lld", (long long)total, (long long)ctx->max_rows);
    if (n_cells_max < 0 || n_cells_max > ctx->max_cells) return fail(ctx, SLAFB_ERR_CAPACITY, "n_cells %lld exceeds context max_cells %lld", (long long)n_cells_max, (long long)ctx->max_cells);
    static const bool trace = [] { const char *e = getenv("SLAFB_SUBMIT_TRACE"); return e && atoi(e) != 0; }();   // stderr: where a submit's host time goes
    if (s.ev_done) {
        CU(ctx, cudaEventSynchronize(s.ev_done));      // the slot's tables (and h_seg) may still be in use by its previous batch
        harvest(ctx, s);
    }
    const double t1 = now_ms();
    int64_t n_cells = 0;
    int rc = fill(s.h_seg, s.h_seg + (size_t)ctx->max_cells + 1, total, &n_cells);    // starts, ids
    if (rc) return rc;
    const double t2 = now_ms();
    // the table must describe the rows: non-decreasing offsets from 0 to the row count (an out-of-range entry would send the
    // tokenizing kernel's bulk copies out of bounds)
    if (n_cells > 0) {
        const uint32_t *st = s.h_seg;
        if (st[0] != 0 || (int64_t)st[n_cells] != total) return fail(ctx, SLAFB_ERR_INVALID, "segment table does not cover the rows");
        uint32_t bad = 0;
        for (int64_t i = 0; i < n_cells; i++) bad |= (st[i] > st[i + 1]) ? 1u : 0u;
        if (bad) return fail(ctx, SLAFB_ERR_INVALID, "segment table offsets are not monotone");
    } else if (total != 0) {
        return fail(ctx, SLAFB_ERR_INVALID, "rows without a segment table");
    }
    cudaStream_t sf = ctx->s_in;
    s.sample = ctx->timing_every > 0 && ((ctx->timing_every == 1) || (ctx->batch_seq % (uint64_t)ctx->timing_every == 0));
    ctx->batch_seq++;
    s.host_segs = true;
    s.packed = false;
    s.n_rows = total;
    s.cell_dtype = in->cell_dtype;
    s.gene_dtype = in->gene_dtype;
    s.value_dtype = in->value_dtype;
    s.h_counters[0] = (uint32_t)n_cells;
    s.h_counters[1] = 0;
    CU(ctx, cudaEventRecord(s.ev_in, static_cast<cudaStream_t>(stream)));
    CU(ctx, cudaStreamWaitEvent(sf, s.ev_in, 0));
    // the segment table comes from the host (8 bytes per cell): no cell column crosses the bus and no CSR build runs
    CU(ctx, cudaMemsetAsync(s.counters, 0, 8 * sizeof(uint32_t), sf));
    if (n_cells > 0) {
        CU(ctx, cudaMemcpyAsync(s.seg_start, s.h_seg, (size_t)(n_cells + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, sf));
        CU(ctx, cudaMemcpyAsync(s.seg_cell, s.h_seg + (size_t)ctx->max_cells + 1, (size_t)n_cells * sizeof(uint32_t), cudaMemcpyHostToDevice, sf));
    }
    CU(ctx, cudaEventRecord(s.ev_front, sf));
    if (total > 0) {
        if (in->location == SLAFB_LOC_HOST || n_pieces > 1) {
            rc = ensure_staging(ctx, s);
            if (rc) return rc;
            std::vector<PieceSrc> src;
            rc = stage_small_pieces(ctx, s, pieces, n_pieces, false, src);
            if (rc) return rc;
            const cudaMemcpyKind kind = in->location == SLAFB_LOC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
            const size_t gsz = dtype_size(in->gene_dtype), vsz = dtype_size(in->value_dtype);
            if (s.sample) CU(ctx, cudaEventRecord(s.t_h2d0, sf));
            size_t off = 0;
            for (int i = 0; i < n_pieces; i++) {
                const size_t n = (size_t)pieces[i].n_rows;
                if (n) CU(ctx, cudaMemcpyAsync(static_cast<char *>(s.in_gene) + off * gsz, src[i].col[1], n * gsz, kind, sf));
                off += n;
            }
            off = 0;
            for (int i = 0; i < n_pieces; i++) {
                const size_t n = (size_t)pieces[i].n_rows;
                if (n) CU(ctx, cudaMemcpyAsync(static_cast<char *>(s.in_value) + off * vsz, src[i].col[2], n * vsz, kind, sf));
                off += n;
            }
            if (s.sample) { CU(ctx, cudaEventRecord(s.t_h2d1, sf)); s.timed_h2d = true; }
            s.cell = nullptr; s.gene = s.in_gene; s.value = s.in_value;
        } else {
            if (((uintptr_t)in->gene | (uintptr_t)in->value) & 15u) return fail(ctx, SLAFB_ERR_INVALID, "device column pointers must be 16-byte aligned");
            s.cell = nullptr; s.gene = in->gene; s.value = in->value;
        }
    }
    CU(ctx, cudaEventRecord(s.ev_loaded, sf));
    ctx->acc.host_submit_ms += now_ms() - t0;
    if (trace) fprintf(stderr, "[slafb submit] slot %d: wait for the slot %.3f ms, segment table %.3f ms (%lld cells), enqueue %.3f ms (%d pieces, %lld rows)\n",
                       si, t1 - t0, t2 - t1, (long long)n_cells, now_ms() - t2, n_pieces, (long long)total);
    s.in_flight = true;
    *slot = si;
    if (n_cells_out) *n_cells_out = n_cells;
    return SLAFB_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int slafb_version(void) { return SLAFB_ABI_VERSION; }

const char *slafb_last_error(const slafb_ctx *ctx) { return ctx ? ctx->err : g_err; }

int slafb_shuffle_perm(int64_t seed, int64_t n, int32_t *perm) {
    if (n < 0 || (n > 0 && !perm) || n > 0x7fffffff) return fail(nullptr, SLAFB_ERR_INVALID, "bad shuffle arguments");
    py_shuffle_perm(seed, n, perm);
    return SLAFB_OK;
}

int slafb_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return fail(nullptr, SLAFB_ERR_INVALID, "ptr is NULL");
    CU(nullptr, cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return SLAFB_OK;
}
int slafb_host_register(void *ptr, size_t bytes) {
    if (!ptr || !bytes) return fail(nullptr, SLAFB_ERR_INVALID, "nothing to register");
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        // read-only mappings (a memory-mapped Arrow file) cannot be registered for writing; the copy engine only reads them
        cudaGetLastError();
        e = cudaHostRegister(ptr, bytes, cudaHostRegisterReadOnly);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, SLAFB_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e)); }
    return SLAFB_OK;
}
int slafb_host_unregister(void *ptr) {
    if (ptr) CU(nullptr, cudaHostUnregister(ptr));
    return SLAFB_OK;
}
int slafb_host_free(void *ptr) {
    if (ptr) CU(nullptr, cudaFreeHost(ptr));
    return SLAFB_OK;
}

int slafb_create(slafb_ctx **out, int device, int64_t max_rows, int64_t max_cells) {
    if (!out) return fail(nullptr, SLAFB_ERR_INVALID, "ctx out pointer is NULL");
    *out = nullptr;
    if (max_rows < 1 || max_rows > 0x7ffffff0ll) return fail(nullptr, SLAFB_ERR_INVALID, "max_rows must be in [1, 2^31)");
    if (max_cells < 1 || max_cells > max_rows) max_cells = max_rows;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(nullptr, SLAFB_ERR_NO_DEVICE, "no CUDA device visible: libslaf_b200 has no CPU fallback");
    if (device < 0 || device >= n_dev) return fail(nullptr, SLAFB_ERR_INVALID, "device %d out of range (%d visible)", device, n_dev);
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, SLAFB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    CU(nullptr, cudaSetDevice(device));
    slafb_ctx *ctx = new (std::nothrow) slafb_ctx();
    if (!ctx) return fail(nullptr, SLAFB_ERR_INVALID, "out of host memory");
    ctx->device = device;
    ctx->n_sm = prop.multiProcessorCount;
    ctx->max_rows = max_rows;
    ctx->max_cells = max_cells;
#define CUC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { int rc__ = fail(nullptr, SLAFB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); slafb_destroy(ctx); return rc__; } } while (0)
    CUC(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&ctx->s_perm, cudaStreamNonBlocking));
    CUC(cudaEventCreateWithFlags(&ctx->ev_k1, cudaEventDisableTiming));
    const size_t n_tiles_max = (size_t)ctx->n_sm * 8;
    CUC(cudaMalloc(&ctx->tile_state, n_tiles_max * sizeof(unsigned long long)));
    CUC(cudaMemset(ctx->tile_state, 0, n_tiles_max * sizeof(unsigned long long)));
    CUC(cudaMalloc(&ctx->vstar_table, 65536 * sizeof(uint32_t)));
    CUC(cudaMalloc(&ctx->ticket, sizeof(uint32_t)));
    CUC(cudaMemset(ctx->ticket, 0, sizeof(uint32_t)));
    // log1p table for uint16 values through the HOST libm (glibc log1p, the function Rust std /
    // polars calls), so the binned window is bit-exact without trusting the device's log1p.
    {
        double *h = (double *)malloc(65536 * sizeof(double));
        if (!h) { slafb_destroy(ctx); return fail(nullptr, SLAFB_ERR_INVALID, "out of host memory"); }
        for (int i = 0; i < 65536; i++) h[i] = log1p((double)i);
        cudaError_t e1 = cudaMalloc(&ctx->log1p_lut, 65536 * sizeof(double));
        cudaError_t e2 = e1 == cudaSuccess ? cudaMemcpy(ctx->log1p_lut, h, 65536 * sizeof(double), cudaMemcpyHostToDevice) : e1;
        free(h);
        CUC(e2);
    }
    // which log1pf the host's libm ships (host_log1pf_kind below): the device uses the evaluation that matches it
    {
        const int use_fdlibm = host_log1pf_kind();
        ctx->log1pf_fdlibm = use_fdlibm;
        CUC(cudaMemcpyToSymbol(c_log1pf_fdlibm, &use_fdlibm, sizeof(int)));
    }
    CUC(cudaMalloc(&ctx->kept_count, ((size_t)max_cells + 1) * sizeof(uint32_t)));
    CUC(cudaMalloc(&ctx->total_dev, sizeof(long long)));
    CUC(cudaHostAlloc(&ctx->h_total, sizeof(long long), cudaHostAllocDefault));
    for (int i = 0; i < kSlots; i++) {
        Slot &s = ctx->slot[i];
        CUC(cudaMalloc(&s.seg_start, ((size_t)max_cells + 2) * sizeof(uint32_t)));
        CUC(cudaMalloc(&s.seg_cell, ((size_t)max_cells + 1) * sizeof(uint32_t)));
        CUC(cudaMalloc(&s.perm_dev, ((size_t)max_cells + 1) * sizeof(int32_t)));
        CUC(cudaMalloc(&s.work_list, ((size_t)max_cells + 1) * sizeof(uint32_t)));
        CUC(cudaMalloc(&s.counters, 8 * sizeof(uint32_t)));
        CUC(cudaMemset(s.counters, 0, 8 * sizeof(uint32_t)));
        CUC(cudaHostAlloc(&s.h_counters, 4 * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable));
        memset(s.h_counters, 0, 4 * sizeof(uint32_t));
        CUC(cudaHostGetDevicePointer(reinterpret_cast<void **>(&s.h_counters_dev), s.h_counters, 0));
        CUC(cudaHostAlloc(&s.h_perm, ((size_t)max_cells + 1) * sizeof(int32_t), cudaHostAllocMapped | cudaHostAllocPortable));
        CUC(cudaHostGetDevicePointer(reinterpret_cast<void **>(&s.h_perm_dev), s.h_perm, 0));
        CUC(cudaHostAlloc(&s.h_seg, (2 * (size_t)max_cells + 2) * sizeof(uint32_t), cudaHostAllocPortable));
        s.h_draws = (uint32_t *)malloc(((size_t)max_cells + 1) * sizeof(uint32_t));
        if (!s.h_draws) { slafb_destroy(ctx); return fail(nullptr, SLAFB_ERR_INVALID, "out of host memory"); }
        cudaEvent_t *plain[] = {&s.ev_front, &s.ev_done, &s.ev_in, &s.ev_loaded, &s.ev_perm};
        for (auto e : plain) CUC(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        cudaEvent_t *timed[] = {&s.t_h2d0, &s.t_h2d1, &s.t_k1a, &s.t_k1b, &s.t_k2a, &s.t_k2b, &s.t_k2c};
        for (auto e : timed) CUC(cudaEventCreate(e));
        CUC(cudaEventRecord(s.ev_done, ctx->s_in));
    }
#undef CUC
    *out = ctx;
    return SLAFB_OK;
}

int slafb_destroy(slafb_ctx *ctx) {
    if (!ctx) return SLAFB_OK;
    cudaSetDevice(ctx->device);
    if (ctx->wstarted) {
        {
            std::lock_guard<std::mutex> lk(ctx->wm);
            ctx->wstop = true;
        }
        ctx->wcv.notify_all();
        for (auto &w : ctx->worker) if (w.joinable()) w.join();
    }
    if (ctx->s_in) cudaStreamSynchronize(ctx->s_in);
    for (int i = 0; i < kSlots; i++) {
        Slot &s = ctx->slot[i];
        if (s.ev_done) cudaEventSynchronize(s.ev_done);
        cudaFree(s.seg_start); cudaFree(s.seg_cell); cudaFree(s.perm_dev); cudaFree(s.work_list); cudaFree(s.counters);
        cudaFree(s.in_cell); cudaFree(s.in_gene); cudaFree(s.in_value);
        if (s.h_counters) cudaFreeHost(s.h_counters);
        if (s.h_perm) cudaFreeHost(s.h_perm);
        if (s.h_seg) cudaFreeHost(s.h_seg);
        if (s.h_small) cudaFreeHost(s.h_small);
        if (s.h_blk) cudaFreeHost(s.h_blk);
        cudaFree(s.pk_bytes); cudaFree(s.blk_off);
        free(s.h_draws);
        cudaEvent_t evs[] = {s.ev_front, s.ev_done, s.ev_in, s.ev_loaded, s.ev_perm, s.t_h2d0, s.t_h2d1, s.t_k1a, s.t_k1b, s.t_k2a, s.t_k2b, s.t_k2c};
        for (auto e : evs) if (e) cudaEventDestroy(e);
    }
    cudaFree(ctx->tile_state); cudaFree(ctx->ticket); cudaFree(ctx->log1p_lut); cudaFree(ctx->vstar_table); cudaFree(ctx->sort_scratch); cudaFree(ctx->rank_scratch); cudaFree(ctx->dup_table); cudaFree(ctx->stats_packed);
    cudaFree(ctx->kept_count); cudaFree(ctx->total_dev);
    if (ctx->h_total) cudaFreeHost(ctx->h_total);
    if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_aux) cudaStreamDestroy(ctx->s_aux);
    if (ctx->s_perm) cudaStreamDestroy(ctx->s_perm);
    delete ctx;
    return SLAFB_OK;
}

int slafb_submit_pieces(slafb_ctx *ctx, const slafb_coo *pieces, int32_t n_pieces, void *stream, int32_t *slot) {
    if (!ctx || !slot) return fail(ctx, SLAFB_ERR_INVALID, "ctx / slot is NULL");
    CU(ctx, cudaSetDevice(ctx->device));
    const int i = pick_slot(ctx, pieces && n_pieces >= 1 && (pieces[0].location == SLAFB_LOC_HOST || n_pieces > 1));
    if (i < 0) return fail(ctx, SLAFB_ERR_BUSY, "all %d pipeline slots hold batches that were not collected yet", kSlots);
    Slot &s = ctx->slot[i];
    const double t0 = now_ms();
    int rc = front(ctx, s, pieces, n_pieces, static_cast<cudaStream_t>(stream));
    ctx->acc.host_submit_ms += now_ms() - t0;
    if (rc) return rc;
    s.in_flight = true;
    *slot = i;
    return SLAFB_OK;
}

int slafb_submit_segments(slafb_ctx *ctx, const slafb_coo *pieces, int32_t n_pieces, const uint32_t *seg_start, const uint32_t *seg_cell,
                          int64_t n_cells, void *stream, int32_t *slot) {
    if (n_cells > 0 && (!seg_start || !seg_cell)) return fail(ctx, SLAFB_ERR_INVALID, "NULL segment table");
    // through the slot's own pinned block: the caller's arrays may be pageable (a pageable async copy stalls the copy stream
    // behind the driver's staging) and need not outlive this call
    return submit_host_segments(ctx, pieces, n_pieces, n_cells, stream, slot, nullptr,
                                [&](uint32_t *st, uint32_t *id, int64_t, int64_t *n_out) -> int {
                                    if (n_cells > 0) {
                                        memcpy(st, seg_start, (size_t)(n_cells + 1) * sizeof(uint32_t));
                                        memcpy(id, seg_cell, (size_t)n_cells * sizeof(uint32_t));
                                    }
                                    *n_out = n_cells;
                                    return SLAFB_OK;
                                });
}

int slafb_csi_segments(const slafb_coo *pieces, int32_t n_pieces, const int64_t *row_lo, const int64_t *csi, int64_t n_csi, int32_t check,
                       uint32_t *seg_start, uint32_t *seg_cell, int64_t capacity_cells, int64_t *n_cells) {
    // Host only (no CUDA call).  Segment table of the concatenated pieces from the index alone: piece i covers table rows
    // [row_lo[i], row_lo[i] + n_i); cell c owns rows [csi[c], csi[c + 1]).  Adjacent pieces that continue one cell give one
    // segment; cells without rows give none.
    if (!pieces || n_pieces < 1 || !row_lo || !csi || n_csi < 1 || !seg_start || !seg_cell || !n_cells)
        return fail(nullptr, SLAFB_ERR_INVALID, "bad pieces / index / table pointers");
    const int64_t n_total = n_csi - 1;                 // cells in the table
    uint32_t *st = seg_start, *id = seg_cell;
    int64_t C = 0, pos = 0;
    for (int i = 0; i < n_pieces; i++) {
        const int64_t n = pieces[i].n_rows;
        if (n < 0) return fail(nullptr, SLAFB_ERR_INVALID, "n_rows < 0");
        if (n == 0) continue;
        const int64_t lo = row_lo[i], hi = lo + n;
        if (lo < csi[0] || hi > csi[n_total]) return fail(nullptr, SLAFB_ERR_INVALID, "piece %d covers table rows [%lld, %lld) outside the cell start index", i, (long long)lo, (long long)hi);
        if (pos + n > 0xffffffffll) return fail(nullptr, SLAFB_ERR_CAPACITY, "more than 2^32 rows in one batch");
        int64_t a = 0, b = n_total;                    // largest c with csi[c] <= lo
        while (b - a > 1) { const int64_t m = (a + b) >> 1; if (csi[m] <= lo) a = m; else b = m; }
        const uint32_t *cells = (check && pieces[i].location == SLAFB_LOC_HOST) ? static_cast<const uint32_t *>(pieces[i].cell) : nullptr;
        for (int64_t c = a; c < n_total && csi[c] < hi; c++) {
            const int64_t first = csi[c] > lo ? csi[c] : lo;
            if (csi[c + 1] <= first) continue;         // no rows (here)
            const int64_t start = first - lo + pos;
            if (C > 0 && (int64_t)id[C - 1] == c && start == pos) continue;      // the previous piece's last cell continues: one segment
            if (C >= capacity_cells) return fail(nullptr, SLAFB_ERR_CAPACITY, "batch has more than %lld cells (table capacity)", (long long)capacity_cells);
            if (c > 0xffffffffll) return fail(nullptr, SLAFB_ERR_INVALID, "cell id does not fit 32 bits");
            st[C] = (uint32_t)start;
            id[C] = (uint32_t)c;
            // consistency probe against the rows themselves (every segment with check = 2, every 16th with check = 1):
            // the cell column must show exactly that id starting exactly there, or the index is stale / shifted
            if (cells && (check >= 2 || (C & 15) == 0)) {
                const int64_t r = first - lo;
                if (cells[r] != (uint32_t)c || (r > 0 && cells[r - 1] == (uint32_t)c))
                    return fail(nullptr, SLAFB_ERR_INVALID, "cell_start_index does not describe the expression rows (cell %lld does not start at table row %lld)", (long long)c, (long long)first);
            }
            C++;
        }
        if (cells && (C == 0 || cells[n - 1] != id[C - 1]))
            return fail(nullptr, SLAFB_ERR_INVALID, "cell_start_index does not describe the expression rows (last cell of piece %d differs)", i);
        pos += n;
    }
    st[C] = (uint32_t)pos;
    *n_cells = C;
    return SLAFB_OK;
}

int slafb_submit_csi(slafb_ctx *ctx, const slafb_coo *pieces, int32_t n_pieces, const int64_t *row_lo, const int64_t *csi, int64_t n_csi,
                     int32_t check, void *stream, int32_t *slot, int64_t *n_cells) {
    if (!ctx) return fail(ctx, SLAFB_ERR_INVALID, "ctx is NULL");
    return submit_host_segments(ctx, pieces, n_pieces, 0, stream, slot, n_cells,
                                [&](uint32_t *st, uint32_t *id, int64_t, int64_t *n_out) -> int {
                                    const int rc = slafb_csi_segments(pieces, n_pieces, row_lo, csi, n_csi, check, st, id, ctx->max_cells, n_out);
                                    if (rc) snprintf(ctx->err, sizeof ctx->err, "%s", g_err);
                                    return rc;
                                });
}

int64_t slafb_pack_bound(int64_t n_rows, int64_t n_cells) {
    // header + two padded byte streams per cell, plus room for one exception per 16 rows (count data needs a handful per cell)
    if (n_rows < 0 || n_cells < 0) return -1;
    return 2 * n_rows + 64 * n_cells + n_rows / 2 + 64;
}

int slafb_pack_rows(const void *gene, int32_t gene_dtype, const uint16_t *value, const uint32_t *row_start, int64_t n_cells, uint8_t *out,
                    int64_t out_capacity, uint32_t *blk_off, int64_t *out_bytes) {
    if (n_cells < 0 || !row_start || !blk_off || !out_bytes || (n_cells > 0 && (!gene || !value || !out)))
        return fail(nullptr, SLAFB_ERR_INVALID, "NULL argument");
    if (gene_dtype != SLAFB_U16 && gene_dtype != SLAFB_I32 && gene_dtype != SLAFB_U32) return fail(nullptr, SLAFB_ERR_INVALID, "gene dtype must be uint16, int32 or uint32");
    if (((uintptr_t)out & 15u) != 0) return fail(nullptr, SLAFB_ERR_INVALID, "the packed buffer must be 16-byte aligned");
    for (int64_t c = 0; c < n_cells; c++)
        if (row_start[c] > row_start[c + 1]) return fail(nullptr, SLAFB_ERR_INVALID, "row offsets are not monotone");
    size_t nb = 0;
    int rc;
    if (gene_dtype == SLAFB_U16)
        rc = slafb_pack::pack<uint16_t>(static_cast<const uint16_t *>(gene), value, row_start, (size_t)n_cells, out, (size_t)out_capacity, blk_off, &nb);
    else
        rc = slafb_pack::pack<int32_t>(static_cast<const int32_t *>(gene), value, row_start, (size_t)n_cells, out, (size_t)out_capacity, blk_off, &nb);
    *out_bytes = (int64_t)nb;
    if (rc == -1) return fail(nullptr, SLAFB_ERR_INVALID, "gene ids do not ascend strictly inside every cell: these rows cannot be packed");
    if (rc == -3) return fail(nullptr, SLAFB_ERR_CAPACITY, "packed buffer too small (%lld bytes)", (long long)out_capacity);
    return SLAFB_OK;
}

int slafb_unpack_rows(const uint8_t *packed, const uint32_t *blk_off, const uint32_t *row_start, int64_t n_cells, void *gene, int32_t gene_dtype,
                      uint16_t *value) {
    if (n_cells < 0 || !blk_off || !row_start || (n_cells > 0 && (!packed || !gene || !value))) return fail(nullptr, SLAFB_ERR_INVALID, "NULL argument");
    const int rc = gene_dtype == SLAFB_U16
                       ? slafb_pack::unpack<uint16_t>(packed, blk_off, (size_t)n_cells, row_start, static_cast<uint16_t *>(gene), value)
                       : slafb_pack::unpack<int32_t>(packed, blk_off, (size_t)n_cells, row_start, static_cast<int32_t *>(gene), value);
    if (rc) return fail(nullptr, SLAFB_ERR_INVALID, "malformed packed block");
    return SLAFB_OK;
}

int slafb_submit_packed(slafb_ctx *ctx, const slafb_packed *pieces, int32_t n_pieces, const uint32_t *seg_cell, int32_t cell_dtype, void *stream,
                        int32_t *slot) {
    if (!ctx || !slot || !pieces || n_pieces < 1) return fail(ctx, SLAFB_ERR_INVALID, "bad ctx / slot / pieces");
    CU(ctx, cudaSetDevice(ctx->device));
    if (cell_dtype != SLAFB_U32 && cell_dtype != SLAFB_I32) return fail(ctx, SLAFB_ERR_INVALID, "cell dtype must be uint32 or int32");
    int64_t C = 0, R = 0, B = 0;
    for (int i = 0; i < n_pieces; i++) {
        const slafb_packed &p = pieces[i];
        if (p.n_cells < 0 || p.n_bytes < 0 || (p.n_bytes & 15) || (p.n_cells > 0 && (!p.bytes || !p.blk_off || !p.row_start)))
            return fail(ctx, SLAFB_ERR_INVALID, "bad packed piece %d", i);
        if (p.gene_dtype != pieces[0].gene_dtype || p.location != pieces[0].location) return fail(ctx, SLAFB_ERR_INVALID, "all pieces must share gene dtype and location");
        if (p.n_cells > 0 && ((int64_t)p.blk_off[p.n_cells] << 4) != p.n_bytes) return fail(ctx, SLAFB_ERR_INVALID, "piece %d: block offsets do not cover the bytes", i);
        if (((uintptr_t)p.bytes & 15u) != 0) return fail(ctx, SLAFB_ERR_INVALID, "packed bytes must be 16-byte aligned");
        C += p.n_cells;
        R += p.n_cells ? p.row_start[p.n_cells] - p.row_start[0] : 0;
        B += p.n_bytes;
    }
    const int gdt = pieces[0].gene_dtype;
    if (gdt != SLAFB_U16 && gdt != SLAFB_I32) return fail(ctx, SLAFB_ERR_INVALID, "packed genes decode to uint16 or int32");
    if (C > ctx->max_cells) return fail(ctx, SLAFB_ERR_CAPACITY, "n_cells %lld exceeds context max_cells %lld", (long long)C, (long long)ctx->max_cells);
    if (R > ctx->max_rows) return fail(ctx, SLAFB_ERR_CAPACITY, "n_rows %lld exceeds context max_rows %lld", (long long)R, (long long)ctx->max_rows);
    if (C > 0 && !seg_cell) return fail(ctx, SLAFB_ERR_INVALID, "NULL cell ids");
    if ((B >> 4) > 0xffffffffll) return fail(ctx, SLAFB_ERR_CAPACITY, "more than 64 GiB of packed rows in one batch");
    const int si = pick_slot(ctx);
    if (si < 0) return fail(ctx, SLAFB_ERR_BUSY, "all %d pipeline slots hold batches that were not collected yet", kSlots);
    Slot &s = ctx->slot[si];
    const double t0 = now_ms();
    if (s.ev_done) {
        CU(ctx, cudaEventSynchronize(s.ev_done));
        harvest(ctx, s);
    }
    int rc = ensure_staging(ctx, s);                  // scratch columns the deferred cells are unpacked into
    if (rc) return rc;
    if (!s.blk_off) CU(ctx, cudaMalloc(&s.blk_off, ((size_t)ctx->max_cells + 2) * sizeof(uint32_t)));
    if (!s.h_blk) CU(ctx, cudaHostAlloc(&s.h_blk, ((size_t)ctx->max_cells + 2) * sizeof(uint32_t), cudaHostAllocPortable));
    if ((size_t)B + 16 > s.pk_cap) {
        if (s.pk_bytes) CU(ctx, cudaFree(s.pk_bytes));
        s.pk_bytes = nullptr; s.pk_cap = 0;
        const size_t cap = (size_t)B + (size_t)B / 8 + 4096;
        CU(ctx, cudaMalloc(&s.pk_bytes, cap));
        s.pk_cap = cap;
    }
    // the batch's tables in the slot's pinned blocks: row offsets, cell ids, block offsets (rebased piece by piece)
    uint32_t *st = s.h_seg, *id = s.h_seg + (size_t)ctx->max_cells + 1, *bo = s.h_blk;
    int64_t c0 = 0, r0 = 0, b0 = 0;
    for (int i = 0; i < n_pieces; i++) {
        const slafb_packed &p = pieces[i];
        for (int64_t c = 0; c < p.n_cells; c++) {
            if (p.row_start[c] > p.row_start[c + 1] || p.blk_off[c] > p.blk_off[c + 1]) return fail(ctx, SLAFB_ERR_INVALID, "piece %d: offsets are not monotone", i);
            st[c0 + c] = (uint32_t)(r0 + (p.row_start[c] - p.row_start[0]));
            bo[c0 + c] = (uint32_t)((b0 >> 4) + p.blk_off[c]);
        }
        if (p.n_cells) { r0 += p.row_start[p.n_cells] - p.row_start[0]; }
        c0 += p.n_cells;
        b0 += p.n_bytes;
    }
    st[C] = (uint32_t)R;
    bo[C] = (uint32_t)(B >> 4);
    if (C) memcpy(id, seg_cell, (size_t)C * sizeof(uint32_t));
    cudaStream_t sf = ctx->s_in;
    s.sample = ctx->timing_every > 0 && ((ctx->timing_every == 1) || (ctx->batch_seq % (uint64_t)ctx->timing_every == 0));
    ctx->batch_seq++;
    s.host_segs = true;
    s.packed = true;
    s.n_rows = R;
    s.cell_dtype = cell_dtype;
    s.gene_dtype = gdt;
    s.value_dtype = SLAFB_U16;
    s.h_counters[0] = (uint32_t)C;
    s.h_counters[1] = 0;
    s.cell = nullptr; s.gene = s.in_cell; s.value = s.in_value;
    CU(ctx, cudaEventRecord(s.ev_in, static_cast<cudaStream_t>(stream)));
    CU(ctx, cudaStreamWaitEvent(sf, s.ev_in, 0));
    CU(ctx, cudaMemsetAsync(s.counters, 0, 8 * sizeof(uint32_t), sf));
    if (C > 0) {
        CU(ctx, cudaMemcpyAsync(s.seg_start, st, (size_t)(C + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, sf));
        CU(ctx, cudaMemcpyAsync(s.seg_cell, id, (size_t)C * sizeof(uint32_t), cudaMemcpyHostToDevice, sf));
        CU(ctx, cudaMemcpyAsync(s.blk_off, bo, (size_t)(C + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, sf));
    }
    CU(ctx, cudaEventRecord(s.ev_front, sf));
    if (s.sample) CU(ctx, cudaEventRecord(s.t_h2d0, sf));
    const cudaMemcpyKind kind = pieces[0].location == SLAFB_LOC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    size_t off = 0;
    for (int i = 0; i < n_pieces; i++) {
        if (pieces[i].n_bytes) CU(ctx, cudaMemcpyAsync(s.pk_bytes + off, pieces[i].bytes, (size_t)pieces[i].n_bytes, kind, sf));
        off += (size_t)pieces[i].n_bytes;
    }
    if (s.sample) { CU(ctx, cudaEventRecord(s.t_h2d1, sf)); s.timed_h2d = true; }
    CU(ctx, cudaEventRecord(s.ev_loaded, sf));
    ctx->acc.host_submit_ms += now_ms() - t0;
    s.in_flight = true;
    *slot = si;
    return SLAFB_OK;
}

int slafb_wait_loaded(slafb_ctx *ctx, int32_t slot) {
    if (!ctx || slot < 0 || slot >= kSlots) return fail(ctx, SLAFB_ERR_INVALID, "bad ctx / slot");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaEventSynchronize(ctx->slot[slot].ev_loaded));
    return SLAFB_OK;
}

int slafb_submit(slafb_ctx *ctx, const slafb_coo *in, void *stream, int32_t *slot) {
    return slafb_submit_pieces(ctx, in, 1, stream, slot);
}

int slafb_prepare_shuffle(slafb_ctx *ctx, int32_t slot, int64_t seed) {
    if (!ctx || slot < 0 || slot >= kSlots) return fail(ctx, SLAFB_ERR_INVALID, "bad ctx / slot");
    Slot &s = ctx->slot[slot];
    if (!s.in_flight) return fail(ctx, SLAFB_ERR_INVALID, "slot %d has no submitted batch", slot);
    if (s.prep_state.load() != 0) return fail(ctx, SLAFB_ERR_BUSY, "slot %d already has a prepared shuffle", slot);
    {
        std::lock_guard<std::mutex> lk(ctx->wm);
        if (!ctx->wstarted) {
            for (auto &w : ctx->worker) w = std::thread(shuffle_worker, ctx);
            ctx->wstarted = true;
        }
        s.prep_seed = seed;
        s.prep_cells = -1;
        ctx->acc.kernel_launches++;      // the helper launches perm_upload_kernel for this batch
        s.prep_state.store(1);
        ctx->wjobs.push_back(slot);
    }
    ctx->wcv.notify_one();
    return SLAFB_OK;
}

int slafb_collect(slafb_ctx *ctx, int32_t slot, const slafb_params *p, const slafb_out *out, int64_t *n_cells, void *stream) {
    if (!ctx || slot < 0 || slot >= kSlots) return fail(ctx, SLAFB_ERR_INVALID, "bad ctx / slot");
    CU(ctx, cudaSetDevice(ctx->device));
    Slot &s = ctx->slot[slot];
    if (!s.in_flight) return fail(ctx, SLAFB_ERR_INVALID, "slot %d has no submitted batch", slot);
    const double t0 = now_ms();
    const int rc = back(ctx, s, p, out, n_cells, static_cast<cudaStream_t>(stream));
    ctx->acc.host_collect_ms += now_ms() - t0;
    // output tensors too small: the CSR of this slot is still valid, so the caller may
    // re-collect with larger tensors; every other outcome releases the slot
    const bool retryable = (rc == SLAFB_ERR_CAPACITY) && n_cells && (*n_cells <= ctx->max_cells);
    if (!retryable) {
        settle_prepared(ctx, s);        // the helper must be done with this slot's buffers before it is reused
        s.prep_state.store(0);
        s.in_flight = false;
        s.freed_at = ++ctx->free_seq;
    }
    return rc;
}

int slafb_pipe_create(slafb_ctx *ctx, const slafb_params *p, const slafb_out *ring, int32_t n_ring, int32_t depth, slafb_pipe **pipe) {
    if (!ctx || !p || !ring || !pipe) return fail(ctx, SLAFB_ERR_INVALID, "NULL argument");
    *pipe = nullptr;
    if (n_ring < 1 || n_ring > 1024) return fail(ctx, SLAFB_ERR_INVALID, "the output ring needs 1..1024 entries");
    if (depth < 0 || depth > kSlots - 1) return fail(ctx, SLAFB_ERR_INVALID, "depth must be in [0, %d]", kSlots - 1);
    if (p->shuffle == SLAFB_SHUFFLE_PERM) return fail(ctx, SLAFB_ERR_INVALID, "a pipe shuffles by seed (or not at all)");
    int rc = check_params(ctx, p);
    if (rc) return rc;
    slafb_pipe *pp = new (std::nothrow) slafb_pipe();
    if (!pp) return fail(ctx, SLAFB_ERR_INVALID, "out of host memory");
    pp->ctx = ctx;
    pp->p = *p;
    pp->ring.assign(ring, ring + n_ring);
    pp->depth = depth;
    *pipe = pp;
    return SLAFB_OK;
}

static int pipe_pop(slafb_pipe *pipe, void *stream, slafb_pipe_result *done) {
    slafb_ctx *ctx = pipe->ctx;
    const slafb_pipe::Item it = pipe->q.front();
    pipe->q.pop_front();
    slafb_params p = pipe->p;
    p.seed = it.seed;
    const int32_t ri = pipe->next_ring;
    int64_t n = 0;
    int rc = slafb_collect(ctx, it.slot, &p, &pipe->ring[ri], &n, stream);
    if (rc == SLAFB_ERR_CAPACITY && ctx->slot[it.slot].in_flight) {
        // the ring's tensors are too small for this batch; the pipe cannot grow them: drop the batch so the slot is not leaked
        char keep[sizeof ctx->err];
        memcpy(keep, ctx->err, sizeof keep);
        slafb_params e = pipe->p;
        e.shuffle = SLAFB_SHUFFLE_PERM; e.perm = nullptr; e.n_perm = 0;
        slafb_out none{};
        int64_t z = 0;
        slafb_collect(ctx, it.slot, &e, &none, &z, stream);
        memcpy(ctx->err, keep, sizeof keep);
    }
    if (rc) return rc;
    pipe->next_ring = (ri + 1) % (int32_t)pipe->ring.size();
    if (done) { done->ring_index = ri; done->n_cells = n; done->seq = it.seq; }
    return SLAFB_OK;
}

int slafb_pipe_push(slafb_pipe *pipe, const slafb_coo *pieces, int32_t n_pieces, const uint32_t *seg_start, const uint32_t *seg_cell,
                    int64_t n_seg_cells, int64_t seed, void *stream, slafb_pipe_result *done) {
    if (!pipe) return fail(nullptr, SLAFB_ERR_INVALID, "pipe is NULL");
    slafb_ctx *ctx = pipe->ctx;
    if (done) { done->ring_index = -1; done->n_cells = 0; done->seq = -1; }
    int32_t slot = -1;
    int rc = seg_start ? slafb_submit_segments(ctx, pieces, n_pieces, seg_start, seg_cell, n_seg_cells, stream, &slot)
                       : slafb_submit_pieces(ctx, pieces, n_pieces, stream, &slot);
    if (rc) return rc;
    if (pipe->p.shuffle == SLAFB_SHUFFLE_SEED) {
        rc = slafb_prepare_shuffle(ctx, slot, seed);       // the helper thread builds the CPython-exact permutation off this thread
        if (rc) return rc;
    }
    pipe->q.push_back({slot, seed, pipe->next_seq++});
    if ((int)pipe->q.size() > pipe->depth) return pipe_pop(pipe, stream, done);
    return SLAFB_OK;
}

int slafb_pipe_push_packed(slafb_pipe *pipe, const slafb_packed *pieces, int32_t n_pieces, const uint32_t *seg_cell, int32_t cell_dtype,
                           int64_t seed, void *stream, slafb_pipe_result *done) {
    if (!pipe) return fail(nullptr, SLAFB_ERR_INVALID, "pipe is NULL");
    slafb_ctx *ctx = pipe->ctx;
    if (done) { done->ring_index = -1; done->n_cells = 0; done->seq = -1; }
    int32_t slot = -1;
    int rc = slafb_submit_packed(ctx, pieces, n_pieces, seg_cell, cell_dtype, stream, &slot);
    if (rc) return rc;
    if (pipe->p.shuffle == SLAFB_SHUFFLE_SEED) {
        rc = slafb_prepare_shuffle(ctx, slot, seed);
        if (rc) return rc;
    }
    pipe->q.push_back({slot, seed, pipe->next_seq++});
    if ((int)pipe->q.size() > pipe->depth) return pipe_pop(pipe, stream, done);
    return SLAFB_OK;
}

int slafb_pipe_pop(slafb_pipe *pipe, void *stream, slafb_pipe_result *done) {
    if (!pipe) return fail(nullptr, SLAFB_ERR_INVALID, "pipe is NULL");
    if (done) { done->ring_index = -1; done->n_cells = 0; done->seq = -1; }
    if (pipe->q.empty()) return SLAFB_OK;
    return pipe_pop(pipe, stream, done);
}

int slafb_pipe_destroy(slafb_pipe *pipe) {
    if (!pipe) return SLAFB_OK;
    while (!pipe->q.empty()) {                            // give the slots of batches still in flight back
        const int32_t slot = pipe->q.front().slot;
        pipe->q.pop_front();
        slafb_params e = pipe->p;
        e.shuffle = SLAFB_SHUFFLE_PERM; e.perm = nullptr; e.n_perm = 0;
        slafb_out none{};
        int64_t z = 0;
        slafb_collect(pipe->ctx, slot, &e, &none, &z, nullptr);
    }
    delete pipe;
    return SLAFB_OK;
}

int slafb_segments(slafb_ctx *ctx, int32_t slot, int64_t *n_cells, uint32_t *seg_start, uint32_t *seg_cell, int64_t capacity_cells) {
    if (!ctx || slot < 0 || slot >= kSlots || !n_cells) return fail(ctx, SLAFB_ERR_INVALID, "bad ctx / slot / n_cells");
    CU(ctx, cudaSetDevice(ctx->device));
    Slot &s = ctx->slot[slot];
    if (!s.in_flight) return fail(ctx, SLAFB_ERR_INVALID, "slot %d has no submitted batch", slot);
    if (!s.host_segs) {
        const double tw0 = now_ms();
        CU(ctx, cudaEventSynchronize(s.ev_front));
        ctx->acc.host_wait_ms += now_ms() - tw0;
    }
    *n_cells = s.h_counters[0];
    if (s.h_counters[1]) return fail(ctx, SLAFB_ERR_CAPACITY, "batch has %u cells, context max_cells is %lld", s.h_counters[0], (long long)ctx->max_cells);
    const int64_t C = *n_cells;
    if (C > capacity_cells) return fail(ctx, SLAFB_ERR_CAPACITY, "batch has %lld cells, host table capacity is %lld", (long long)C, (long long)capacity_cells);
    if (C == 0) { if (seg_start) seg_start[0] = 0; return SLAFB_OK; }
    if (!seg_start || !seg_cell) return fail(ctx, SLAFB_ERR_INVALID, "NULL table pointer");
    if (s.host_segs) {                       // the table was built on the host: it is still in the slot's pinned block
        memcpy(seg_start, s.h_seg, (size_t)(C + 1) * sizeof(uint32_t));
        memcpy(seg_cell, s.h_seg + (size_t)ctx->max_cells + 1, (size_t)C * sizeof(uint32_t));
        return SLAFB_OK;
    }
    CU(ctx, cudaMemcpyAsync(seg_start, s.seg_start, (size_t)(C + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->s_aux));
    CU(ctx, cudaMemcpyAsync(seg_cell, s.seg_cell, (size_t)C * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->s_aux));
    CU(ctx, cudaStreamSynchronize(ctx->s_aux));
    return SLAFB_OK;
}

int slafb_tokenize(slafb_ctx *ctx, const slafb_coo *in, const slafb_params *p, const slafb_out *out, int64_t *n_cells, void *stream) {
    int32_t slot = -1;
    int rc = slafb_submit(ctx, in, stream, &slot);
    if (rc) return rc;
    return slafb_collect(ctx, slot, p, out, n_cells, stream);
}

int slafb_collect_csr(slafb_ctx *ctx, int32_t slot, const slafb_params *p, int64_t *crow, int32_t *col, void *val, int64_t *cell_ids,
                      int64_t capacity_cells, int64_t capacity_rows, int64_t *n_cells, int64_t *n_rows, void *stream) {
    if (!ctx || slot < 0 || slot >= kSlots || !p || !n_cells || !n_rows) return fail(ctx, SLAFB_ERR_INVALID, "bad ctx / slot / argument");
    CU(ctx, cudaSetDevice(ctx->device));
    Slot &s = ctx->slot[slot];
    if (!s.in_flight) return fail(ctx, SLAFB_ERR_INVALID, "slot %d has no submitted batch", slot);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bool use_perm = false;
    int rc = s.packed ? fail(ctx, SLAFB_ERR_INVALID, "packed batches are tokenized (slafb_collect), not returned as CSR") : middle(ctx, s, p, n_cells, st, &use_perm);
    const int64_t C = *n_cells;
    *n_rows = 0;
    if (rc == SLAFB_OK && C > 0) {
        if (C > capacity_cells) rc = fail(ctx, SLAFB_ERR_CAPACITY, "batch has %lld cells, output capacity is %lld", (long long)C, (long long)capacity_cells);
        else if (s.n_rows > capacity_rows) rc = fail(ctx, SLAFB_ERR_CAPACITY, "batch has %lld rows, output capacity is %lld", (long long)s.n_rows, (long long)capacity_rows);
        else if (!crow || !col || !val || !cell_ids) rc = fail(ctx, SLAFB_ERR_INVALID, "NULL output");
    }
    if (rc == SLAFB_OK && C > 0) {
        const int32_t *perm = use_perm ? s.perm_dev : nullptr;
        csr_lengths_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(s.seg_start, s.seg_cell, perm, (uint32_t)C, s.cell_dtype == SLAFB_I32,
                                                                        ctx->kept_count, reinterpret_cast<long long *>(cell_ids));
        scan_counts_kernel<<<1, 1024, 0, st>>>(ctx->kept_count, reinterpret_cast<long long *>(crow), (uint32_t)C, ctx->total_dev);
        const unsigned grid = (unsigned)ctx->n_sm * 8u;
        DISPATCH_TYPES(s.gene_dtype, s.value_dtype,
                       (csr_gather_kernel<GeneT, ValT><<<grid, 256, 0, st>>>(static_cast<const GeneT *>(s.gene), static_cast<const ValT *>(s.value),
                                                                           s.seg_start, perm, (uint32_t)C, reinterpret_cast<const long long *>(crow),
                                                                           col, static_cast<ValT *>(val))));
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(ctx, SLAFB_ERR_CUDA, "csr kernels failed: %s", cudaGetErrorString(e));
        ctx->acc.kernel_launches += 3;
        if (rc == SLAFB_OK) {
            cudaMemcpyAsync(ctx->h_total, ctx->total_dev, sizeof(long long), cudaMemcpyDeviceToHost, st);
            cudaEventRecord(s.ev_done, st);
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc = fail(ctx, SLAFB_ERR_CUDA, "csr collect failed: %s", cudaGetErrorString(e));
            *n_rows = *ctx->h_total;
            ctx->acc.batches++;
            ctx->acc.cells += C;
            ctx->acc.rows += s.n_rows;
        }
    } else if (rc == SLAFB_OK) {
        cudaEventRecord(s.ev_done, st);
    }
    settle_prepared(ctx, s);
    s.prep_state.store(0);
    s.in_flight = false;
    s.freed_at = ++ctx->free_seq;
    return rc;
}

int slafb_window(slafb_ctx *ctx, const slafb_coo *in, const slafb_params *p, int64_t *cell_ids, int64_t *offsets,
                 int64_t *gene_list, double *expr_list, int64_t capacity_cells, int64_t *n_cells, int64_t *n_items,
                 void *stream) {
    if (!ctx || !n_cells || !n_items) return fail(ctx, SLAFB_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = check_params(ctx, p);
    if (rc) return rc;
    const int wi = pick_slot(ctx);
    if (wi < 0) return fail(ctx, SLAFB_ERR_BUSY, "all %d pipeline slots hold batches that were not collected yet", kSlots);
    Slot &s = ctx->slot[wi];
    rc = front(ctx, s, in, 1, st);
    if (rc) return rc;
    bool use_perm = false;
    rc = middle(ctx, s, p, n_cells, st, &use_perm);
    if (rc) return rc;
    const int64_t C = *n_cells;
    *n_items = 0;
    if (C == 0) { CU(ctx, cudaEventRecord(s.ev_done, st)); return SLAFB_OK; }
    if (C > capacity_cells) return fail(ctx, SLAFB_ERR_CAPACITY, "batch has %lld cells, output capacity is %lld", (long long)C, (long long)capacity_cells);
    const bool scgpt = p->mode >= SLAFB_SCGPT_RAW;
    if (!cell_ids || !offsets || !gene_list || (scgpt && !expr_list)) return fail(ctx, SLAFB_ERR_INVALID, "NULL output");
    rc = ensure_sort_scratch(ctx);
    if (rc) return rc;
    TokArgs a;
    fill_tok_args(ctx, s, p, a, (uint32_t)C, use_perm);
    a.all_cells = 1;
    WindowSink w{};
    w.kept_count = ctx->kept_count;
    w.offsets = reinterpret_cast<const long long *>(offsets);
    w.gene_list = reinterpret_cast<long long *>(gene_list);
    w.expr_list = expr_list;
    // pass A: kept rows per cell
    CU(ctx, cudaMemsetAsync(s.counters + 3, 0, sizeof(uint32_t), st));
    DISPATCH_TYPES(s.gene_dtype, s.value_dtype, rc = (dispatch_general<GeneT, ValT, 1>(ctx, p->mode, a, w, st)));
    if (rc) return rc;
    scan_counts_kernel<<<1, 1024, 0, st>>>(ctx->kept_count, reinterpret_cast<long long *>(offsets), (uint32_t)C, ctx->total_dev);
    CU(ctx, cudaGetLastError());
    // pass B: flat lists at the scanned offsets
    CU(ctx, cudaMemsetAsync(s.counters + 3, 0, sizeof(uint32_t), st));
    DISPATCH_TYPES(s.gene_dtype, s.value_dtype, rc = (dispatch_general<GeneT, ValT, 2>(ctx, p->mode, a, w, st)));
    if (rc) return rc;
    cell_ids_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(s.seg_cell, use_perm ? s.perm_dev : nullptr, (uint32_t)C,
                                                                 s.cell_dtype == SLAFB_I32, reinterpret_cast<long long *>(cell_ids));
    CU(ctx, cudaGetLastError());
    ctx->acc.kernel_launches += 2;
    CU(ctx, cudaMemcpyAsync(ctx->h_total, ctx->total_dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaEventRecord(s.ev_done, st));
    CU(ctx, cudaStreamSynchronize(st));
    *n_items = *ctx->h_total;
    return SLAFB_OK;
}

int slafb_tokenize_lists(slafb_ctx *ctx, int32_t scgpt, const int64_t *offsets, const int64_t *genes, const double *exprs,
                         int32_t exprs_are_int, int64_t n_lists, int32_t max_genes, int32_t n_bins, int64_t vocab_size,
                         const slafb_out *out, void *stream) {
    if (!ctx || !out) return fail(ctx, SLAFB_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    if (max_genes < 1) return fail(ctx, SLAFB_ERR_INVALID, "max_genes must be >= 1, got %d", max_genes);
    if (n_lists < 0 || n_lists > out->capacity_cells) return fail(ctx, SLAFB_ERR_CAPACITY, "n_lists %lld exceeds output capacity %lld", (long long)n_lists, (long long)out->capacity_cells);
    if (n_lists == 0) return SLAFB_OK;
    if (!offsets || !genes || (scgpt && !exprs) || !out->input_ids || !out->attention_mask || (scgpt && !out->values))
        return fail(ctx, SLAFB_ERR_INVALID, "NULL tensor");
    if (scgpt && n_bins < 1) return fail(ctx, SLAFB_ERR_INVALID, "n_bins must be >= 1");
    ListArgs a{};
    a.offsets = reinterpret_cast<const long long *>(offsets);
    a.genes = reinterpret_cast<const long long *>(genes);
    a.exprs = exprs;
    a.exprs_are_int = exprs_are_int;
    a.scgpt = scgpt;
    a.n_lists = (uint32_t)n_lists;
    a.max_genes = (uint32_t)max_genes;
    a.L = scgpt ? (uint32_t)max_genes + 2u : (uint32_t)max_genes;
    a.n_bins = n_bins;
    a.vocab = vocab_size;
    a.bin_size = (float)(1.0 / (double)(n_bins > 0 ? n_bins : 1));
    a.ids = reinterpret_cast<long long *>(out->input_ids);
    a.mask = out->attention_mask;
    a.vals = reinterpret_cast<long long *>(out->values);
    tokenize_lists_kernel<<<(unsigned)n_lists, kTokThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    CU(ctx, cudaGetLastError());
    ctx->acc.kernel_launches++;
    return SLAFB_OK;
}

int slafb_gene_stats(slafb_ctx *ctx, const slafb_coo *in, int64_t n_genes, uint64_t *sum, uint64_t *cnt, void *stream) {
    if (!ctx || !sum || !cnt) return fail(ctx, SLAFB_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = check_coo(ctx, in);
    if (rc) return rc;
    if (in->location != SLAFB_LOC_DEVICE) return fail(ctx, SLAFB_ERR_INVALID, "gene_stats takes device-resident columns");
    if (n_genes < 1 || n_genes > 0x7fffffff) return fail(ctx, SLAFB_ERR_INVALID, "bad n_genes");
    if (in->n_rows == 0) return SLAFB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)ctx->n_sm * 8u;
    auto *usum = reinterpret_cast<unsigned long long *>(sum);
    auto *ucnt = reinterpret_cast<unsigned long long *>(cnt);
    if (in->value_dtype == SLAFB_U16) {
        // counts: one packed 64-bit reduction per nonzero row (kernels.cuh, K5), in launches of < 2^24 rows
        if ((size_t)n_genes > ctx->stats_genes) {
            if (ctx->stats_packed) CU(ctx, cudaFree(ctx->stats_packed));
            ctx->stats_packed = nullptr; ctx->stats_genes = 0;
            CU(ctx, cudaMalloc(&ctx->stats_packed, (size_t)n_genes * sizeof(unsigned long long)));
            CU(ctx, cudaMemsetAsync(ctx->stats_packed, 0, (size_t)n_genes * sizeof(unsigned long long), st));
            ctx->stats_genes = (size_t)n_genes;
        }
        const unsigned ublocks = (unsigned)((n_genes + 255) / 256);
        for (int64_t lo = 0; lo < in->n_rows; lo += kStatsRowsPerLaunch) {
            const uint32_t n = (uint32_t)(in->n_rows - lo < (int64_t)kStatsRowsPerLaunch ? in->n_rows - lo : (int64_t)kStatsRowsPerLaunch);
            const uint16_t *val = static_cast<const uint16_t *>(in->value) + lo;
            if (in->gene_dtype == SLAFB_U16)
                gene_stats_packed_kernel<uint16_t><<<grid, 256, 0, st>>>(static_cast<const uint16_t *>(in->gene) + lo, val, n, (uint32_t)n_genes, ctx->stats_packed);
            else
                gene_stats_packed_kernel<int32_t><<<grid, 256, 0, st>>>(static_cast<const int32_t *>(in->gene) + lo, val, n, (uint32_t)n_genes, ctx->stats_packed);
            gene_stats_unpack_kernel<<<ublocks, 256, 0, st>>>(ctx->stats_packed, (uint32_t)n_genes, usum, ucnt);
            CU(ctx, cudaGetLastError());
            ctx->acc.kernel_launches += 2;
        }
        return SLAFB_OK;
    }
    DISPATCH_TYPES(in->gene_dtype, in->value_dtype,
                   (gene_stats_kernel<GeneT, ValT><<<grid, 256, 0, st>>>(static_cast<const GeneT *>(in->gene), static_cast<const ValT *>(in->value),
                                                                       (uint32_t)in->n_rows, (uint32_t)n_genes, usum, ucnt)));
    CU(ctx, cudaGetLastError());
    ctx->acc.kernel_launches++;
    return SLAFB_OK;
}

int slafb_gene_hist(slafb_ctx *ctx, const slafb_coo *in, int64_t n_genes, int32_t n_bins, uint32_t *hist, void *stream) {
    if (!ctx || !hist) return fail(ctx, SLAFB_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = check_coo(ctx, in);
    if (rc) return rc;
    if (in->location != SLAFB_LOC_DEVICE) return fail(ctx, SLAFB_ERR_INVALID, "gene_hist takes device-resident columns");
    if (in->value_dtype != SLAFB_U16) return fail(ctx, SLAFB_ERR_INVALID, "gene_hist needs uint16 values (exact medians are defined on counts)");
    if (n_genes < 1 || n_genes > 0x7fffffff || n_bins < 2 || n_bins > 65536) return fail(ctx, SLAFB_ERR_INVALID, "bad n_genes / n_bins");
    if (in->n_rows == 0) return SLAFB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)ctx->n_sm * 8u;
    if (in->gene_dtype == SLAFB_U16)
        gene_hist_kernel<uint16_t><<<grid, 256, 0, st>>>(static_cast<const uint16_t *>(in->gene), static_cast<const uint16_t *>(in->value),
                                                         (uint32_t)in->n_rows, (uint32_t)n_genes, (uint32_t)n_bins, hist);
    else
        gene_hist_kernel<int32_t><<<grid, 256, 0, st>>>(static_cast<const int32_t *>(in->gene), static_cast<const uint16_t *>(in->value),
                                                        (uint32_t)in->n_rows, (uint32_t)n_genes, (uint32_t)n_bins, hist);
    CU(ctx, cudaGetLastError());
    ctx->acc.kernel_launches++;
    return SLAFB_OK;
}

int slafb_gene_median(slafb_ctx *ctx, const uint32_t *hist, int64_t n_genes, int32_t n_bins, float *median, uint8_t *exact, void *stream) {
    if (!ctx || !hist || !median || !exact) return fail(ctx, SLAFB_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    if (n_genes < 1 || n_genes > 0x3ffffff || n_bins < 2 || n_bins > 65536) return fail(ctx, SLAFB_ERR_INVALID, "bad n_genes / n_bins");
    const unsigned blocks = (unsigned)((n_genes * 32 + 255) / 256);
    gene_median_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(hist, (uint32_t)n_genes, (uint32_t)n_bins, median, exact);
    CU(ctx, cudaGetLastError());
    ctx->acc.kernel_launches++;
    return SLAFB_OK;
}

int slafb_debug_log1pf(int32_t kind, const float *x, float *y, int64_t n) {
    if (kind >= 0 && x && y)
        for (int64_t i = 0; i < n; i++) y[i] = kind ? log1pf_fdlibm(x[i]) : (float)log1p((double)x[i]);
    return host_log1pf_kind();
}

int slafb_debug_check_failures(uint32_t *info) {
#ifdef SLAFB_CHECKED
    unsigned int h[8] = {0};
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(h, slafb::g_check_fail, sizeof h) != cudaSuccess) {
        cudaGetLastError();
        return -2;
    }
    if (info) memcpy(info, h, sizeof h);
    return (int)(h[0] > 0x7fffffffu ? 0x7fffffffu : h[0]);
#else
    (void)info;
    return -1;      // not a checked build
#endif
}

int slafb_set_option(slafb_ctx *ctx, int32_t option, int64_t value) {
    if (!ctx) return fail(ctx, SLAFB_ERR_INVALID, "ctx is NULL");
    switch (option) {
    case SLAFB_OPT_TIMING_INTERVAL:
        if (value < 0 || value > 1000000) return fail(ctx, SLAFB_ERR_INVALID, "timing interval must be in [0, 1e6]");
        ctx->timing_every = (int)value;
        return SLAFB_OK;
    case SLAFB_OPT_FRONT_ON_CALLER:
        ctx->front_on_caller = value != 0;
        return SLAFB_OK;
    default:
        return fail(ctx, SLAFB_ERR_INVALID, "unknown option %d", option);
    }
}

int slafb_get_timing(slafb_ctx *ctx, slafb_timing *t) {
    if (!ctx || !t) return fail(ctx, SLAFB_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->s_in));
    for (int i = 0; i < kSlots; i++) {
        Slot &s = ctx->slot[i];
        if (s.in_flight) continue;
        CU(ctx, cudaEventSynchronize(s.ev_done));
        harvest(ctx, s);
    }
    *t = ctx->acc;
    memset(&ctx->acc, 0, sizeof ctx->acc);
    return SLAFB_OK;
}

}  // extern "C"
