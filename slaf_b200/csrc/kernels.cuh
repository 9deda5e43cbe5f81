// kernels.cuh -- sm_100a device code of the SLAF hot path.
//
//   K1  csr_build_kernel        COO cell column -> segment table (CSR offsets + cell ids), one pass:
//                               every warp scans one contiguous span of rows (128-bit loads, head
//                               detection with shuffles/ballots, heads kept in shared memory), one
//                               decoupled look-back per CTA gives the global offsets.
//                               Stands in for polars partition_by / group_by(maintain_order=True)
//                               (reference slaf/ml/samplers.py:93, slaf/ml/aggregators.py:111,190,208).
//   K2  tokenize_cells_kernel   persistent CTAs, one cell at a time, the NEXT cell's gene/value
//                               segments staged into shared memory by TMA bulk copies
//                               (cp.async.bulk + mbarrier, double buffered) while the current cell
//                               is emitted: dense-rank<=k filter decided from the value range,
//                               log1p-bin threshold from a per-context table, token pairs with
//                               CLS/SEP/PAD, attention mask and cell id written straight into the
//                               batch tensors with 128-bit stores.
//                               (aggregators.py:87-134,197-214 + tokenizers.py:407-479,666-722
//                                + datasets.py:1087-1098)
//   K2w tokenize_pct_warp_kernel  the min_percentile branch (aggregators.py:167-196) on uint16 values, one WARP per cell.
//   K2f tokenize_f32_warp_kernel  all four modes on float32 values, one warp per cell (exact set of the cell's distinct values).
//   K2g tokenize_general_kernel same emission, but with the exact k-th-distinct-value machinery
//                               (value bitmap for uint16, shared/global bitonic sort for float32);
//                               runs over the (normally empty) work list K2 / K2w / K2f leave behind.
//   K3  window_*                Window.apply facade: kept-count, offsets, flat list columns.
//   K4  tokenize_lists_kernel   SLAFTokenizer.tokenize facade on pre-windowed lists.
//   K5  gene_stats_kernel       per-gene nonzero sum / count (extension, BASELINE.json config 4).
//   K6  tokenize_rank_bucket_kernel (bucket distribution), tokenize_rank_kernel (sorting network, for what the first leaves)
//                               Geneformer rank-value encoding with per-gene normalisation / rank order (extension).
//   K7  gene_hist / gene_median per-gene value histogram and exact nonzero medians (extension).
//   K8  csr_lengths / csr_gather raw mode on the device: the batch as CSR in shuffled cell order
//                               (RawPrefetchBatch, reference slaf/ml/datasets.py:958-1008).
//
// All of it is HBM-bound integer/byte work: no tensor cores.  Loads are 128-bit, aligned to
// 8-row groups of the column arrays; stores are coalesced 8/16-byte vectors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace slafb {

// output stores of K2: SLAFB_STORE_MODE 0 = default (write-back), 1 = st.global.cs (streaming), 2 = st.global.wt
#ifndef SLAFB_STORE_MODE
#define SLAFB_STORE_MODE 0
#endif
__device__ __forceinline__ void store_pair(long long *p, long long x0, long long x1) {
#if SLAFB_STORE_MODE == 1
    __stcs(reinterpret_cast<longlong2 *>(p), make_longlong2(x0, x1));
#elif SLAFB_STORE_MODE == 2
    __stwt(reinterpret_cast<longlong2 *>(p), make_longlong2(x0, x1));
#else
    *reinterpret_cast<longlong2 *>(p) = make_longlong2(x0, x1);
#endif
}

// Checked build (-DSLAFB_CHECKED): every index the kernels compute for a shared-memory stage, a scratch table or an output
// tensor is compared with its bound before use; the first violation is recorded (code, value, CTA, thread) in a device word the
// host reads with slafb_debug_check_failures().  compute-sanitizer is closed on the pool these kernels are developed on
// (profiles/r2a_sanitizer_refused.txt), so the GPU test-suite is run once against this build instead (scripts/checked_build.sh).
#ifdef SLAFB_CHECKED
__device__ unsigned int g_check_fail[8];
#define SLAFB_CHECK(cond, code, val)                                                                              \
    do {                                                                                                          \
        if (!(cond)) {                                                                                            \
            if (atomicAdd(&g_check_fail[0], 1u) == 0u) {                                                          \
                g_check_fail[1] = (unsigned)(code); g_check_fail[2] = (unsigned)(val);                            \
                g_check_fail[3] = blockIdx.x; g_check_fail[4] = threadIdx.x;                                      \
            }                                                                                                     \
        }                                                                                                         \
    } while (0)
#else
#define SLAFB_CHECK(cond, code, val) do { } while (0)
#endif

constexpr int kTokThreads = 256;
constexpr int kGroup = 8;                           // rows per thread per chunk (one 128-bit load of u16)
constexpr int kChunkRows = kTokThreads * kGroup;    // 2048 rows staged per iteration
constexpr int kCsrListCap = 1024;                   // segment heads a CTA keeps in shared memory (else: replay)
#ifndef SLAFB_FAST_THREADS
#define SLAFB_FAST_THREADS 256
#endif
constexpr int kFastThreads = SLAFB_FAST_THREADS;    // K2: threads per persistent CTA (build-time tuning knob)
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kBitmapWords = 65536 / 32;
constexpr int kSortSmemKeys = 12288;                // general path: float keys sorted in shared memory up to here
constexpr int kHashBits = 11;                       // K2, float32 values: distinct-value set of 2048 slots in shared memory,
constexpr int kHashSlots = 1 << kHashBits;          // trusted up to kHashMaxDistinct entries (every thread may add one more
constexpr int kHashMaxDistinct = 1536;              // before it sees the limit: 1536 + 256 < 2048, so probing always ends)
constexpr int kPctBits = 8192;                      // K2, percentile mode on uint16 values: presence bitmap over [lo, hi] of the cell
constexpr int kPctWords = kPctBits / 32;            // (1 KiB of shared memory); a wider value range goes to the general kernel

enum : int { MODE_GF = 0, MODE_GF_PCT = 1, MODE_SCGPT_RAW = 2, MODE_SCGPT_BINNED = 3 };

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ uint32_t warp_min(uint32_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
__device__ __forceinline__ uint32_t warp_max(uint32_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

// Block-wide exclusive scan over kTokThreads threads.  `ws` = 16 words of shared scratch.
// Two barriers; safe to call back to back (second barrier protects the scratch).
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *ws, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan(v);
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    uint32_t wv = (lane < kTokThreads / 32) ? ws[lane] : 0u;
    uint32_t winc = warp_incl_scan(wv);
    uint32_t wbase = __shfl_sync(0xffffffffu, winc - wv, warp);
    total = __shfl_sync(0xffffffffu, winc, kTokThreads / 32 - 1);
    __syncthreads();
    return wbase + inc - v;
}

__device__ __forceinline__ void block_minmax(uint32_t &lo, uint32_t &hi, uint32_t *ws) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    lo = warp_min(lo);
    hi = warp_max(hi);
    if (lane == 0) { ws[warp] = lo; ws[8 + warp] = hi; }
    __syncthreads();
    uint32_t l = (lane < kTokThreads / 32) ? ws[lane] : 0xffffffffu;
    uint32_t h = (lane < kTokThreads / 32) ? ws[8 + lane] : 0u;
    lo = warp_min(l);
    hi = warp_max(h);
    __syncthreads();
}

// x << n with PTX semantics: shift amounts >= 32 give 0 (the C operator leaves them undefined)
__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, uint32_t n) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n));
    return r;
}

// order-preserving 32-bit key of a value (dense rank compares keys)
__device__ __forceinline__ uint32_t key_of(uint16_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t key_of(float x) {
    x = x + 0.0f;  // -0.0 -> +0.0 (equal under IEEE compare, must be one distinct value)
    uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_of_key(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint32_t raw_bits(uint16_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t raw_bits(float v) { return __float_as_uint(v); }

// 8 consecutive elements starting at row r0 (r0 % 8 == 0, base 16-byte aligned); rows >= R read as 0.
template <typename T>
__device__ __forceinline__ void load8(const T *__restrict__ p, uint32_t r0, uint32_t R, T (&out)[8]) {
    if (r0 + 8u <= R) {
        if constexpr (sizeof(T) == 2) {
            uint4 q = __ldg(reinterpret_cast<const uint4 *>(p + r0));
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint16_t a = (uint16_t)(w[i] & 0xffffu), b = (uint16_t)(w[i] >> 16);
                out[2 * i] = *reinterpret_cast<T *>(&a);
                out[2 * i + 1] = *reinterpret_cast<T *>(&b);
            }
        } else {
            uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(p + r0));
            uint4 q1 = __ldg(reinterpret_cast<const uint4 *>(p + r0) + 1);
            uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) out[i] = *reinterpret_cast<T *>(&w[i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            T z{};
            out[i] = (r0 + i < R) ? __ldg(p + r0 + i) : z;
        }
    }
}

// numpy `.astype(int)` on x86-64: out-of-range / NaN -> INT64_MIN (cvttss2si "integer indefinite")
__device__ __forceinline__ long long f32_to_i64_x86(float q) {
    if (!(q >= -9223372036854775808.0f && q < 9223372036854775808.0f)) return (long long)0x8000000000000000ull;
    return (long long)q;
}
// ScGPTTokenizer._expression_to_bin_vectorized on a float32 (tokenizers.py:521-543)
__device__ __forceinline__ long long expr_bin_token(float x, float bin_size, int n_bins, long long vocab) {
    float q = __fdiv_rn(x, bin_size);
    long long b = f32_to_i64_x86(q);
    b = b < 0 ? 0 : b;
    b = b > (long long)(n_bins - 1) ? (long long)(n_bins - 1) : b;
    return x > 0.0f ? vocab + b : 0ll;
}

// float32 log1p used by the binned window on float32 storage: polars keeps Float32 and calls Rust's f32::ln_1p, i.e. the
// platform's libm log1pf.  Two libms exist in practice: glibc up to 2.39 ships the fdlibm float algorithm (within 1 ulp, not
// correctly rounded: it differs from the correctly rounded value on 9,149,209 of the 2,139,095,040 non-negative floats), glibc
// from 2.40 the correctly rounded CORE-MATH one.  The device has both and uses the one the host's libm matches (probed at
// context creation, slaf_b200.cu: c_log1pf_fdlibm), so float32 bins equal what polars computes on the same machine bit for bit
// on every input.  log1pf_fdlibm restates the published fdlibm algorithm (s_log1pf.c, Sun Microsystems 1993; float
// version by Ian Lance Taylor, Cygnus) operation by operation with round-to-nearest intrinsics (no FMA contraction);
// scripts/micro/log1pf_check.c compares the same restatement with the host's log1pf on ALL non-negative floats (0 differences
// on glibc 2.39) and checks that the host's function is monotone there (the per-cell bin threshold relies on it: 0 violations).
__constant__ int c_log1pf_fdlibm;
// (round-to-nearest float operations that the compiler may not contract: intrinsics on the device, plain operators on the host,
// where x86-64 code is generated without FMA instructions)
#ifdef __CUDA_ARCH__
#define SLAFB_FADD(a, b) __fadd_rn((a), (b))
#define SLAFB_FSUB(a, b) __fsub_rn((a), (b))
#define SLAFB_FMUL(a, b) __fmul_rn((a), (b))
#define SLAFB_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define SLAFB_FADD(a, b) ((float)((a) + (b)))
#define SLAFB_FSUB(a, b) ((float)((a) - (b)))
#define SLAFB_FMUL(a, b) ((float)((a) * (b)))
#define SLAFB_FDIV(a, b) ((float)((a) / (b)))
#endif
__host__ __device__ inline uint32_t slafb_f2u(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
__host__ __device__ inline float slafb_u2f(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
__host__ __device__ inline float log1pf_fdlibm(float x) {
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f, two25 = 3.355443200e+07f;
    const float Lp1 = 6.6666668653e-01f, Lp2 = 4.0000000596e-01f, Lp3 = 2.8571429849e-01f, Lp4 = 2.2222198546e-01f,
                Lp5 = 1.8183572590e-01f, Lp6 = 1.5313838422e-01f, Lp7 = 1.4798198640e-01f;
    float f = 0.0f, c = 0.0f, u;
    int32_t hx = (int32_t)slafb_f2u(x), hu = 0, k = 1;
    const int32_t ax = hx & 0x7fffffff;
    if (hx < 0x3ed413d7) {                                                  // x < 0.41422
        if (ax >= 0x3f800000) {                                             // x <= -1.0
            if (x == -1.0f) return SLAFB_FDIV(-two25, 0.0f);
            return SLAFB_FDIV(SLAFB_FSUB(x, x), SLAFB_FSUB(x, x));
        }
        if (ax < 0x31000000) {                                              // |x| < 2**-29
            if (ax < 0x24800000) return x;                                  // |x| < 2**-54
            return SLAFB_FSUB(x, SLAFB_FMUL(SLAFB_FMUL(x, x), 0.5f));
        }
        if (hx > 0 || hx <= (int32_t)0xbe95f61f) { k = 0; f = x; hu = 1; }  // -0.2929 < x < 0.41422
    }
    if (hx >= 0x7f800000) return SLAFB_FADD(x, x);
    if (k != 0) {
        if (hx < 0x5a000000) {
            u = SLAFB_FADD(1.0f, x);
            hu = (int32_t)slafb_f2u(u);
            k = (hu >> 23) - 127;
            c = (k > 0) ? SLAFB_FSUB(1.0f, SLAFB_FSUB(u, x)) : SLAFB_FSUB(x, SLAFB_FSUB(u, 1.0f));      // correction term
            c = SLAFB_FDIV(c, u);
        } else {
            u = x;
            hu = (int32_t)slafb_f2u(u);
            k = (hu >> 23) - 127;
            c = 0.0f;
        }
        hu &= 0x007fffff;
        if (hu < 0x3504f7) {
            u = slafb_u2f((uint32_t)(hu | 0x3f800000));               // normalize u
        } else {
            k += 1;
            u = slafb_u2f((uint32_t)(hu | 0x3f000000));               // normalize u/2
            hu = (0x00800000 - hu) >> 2;
        }
        f = SLAFB_FSUB(u, 1.0f);
    }
    const float kf = (float)k;
    const float hfsq = SLAFB_FMUL(SLAFB_FMUL(0.5f, f), f);
    if (hu == 0) {                                                          // |f| < 2**-20
        if (f == 0.0f) {
            if (k == 0) return 0.0f;
            c = SLAFB_FADD(c, SLAFB_FMUL(kf, ln2_lo));
            return SLAFB_FADD(SLAFB_FMUL(kf, ln2_hi), c);
        }
        const float R = SLAFB_FMUL(hfsq, SLAFB_FSUB(1.0f, SLAFB_FMUL(0.66666666666666666f, f)));
        if (k == 0) return SLAFB_FSUB(f, R);
        return SLAFB_FSUB(SLAFB_FMUL(kf, ln2_hi), SLAFB_FSUB(SLAFB_FSUB(R, SLAFB_FADD(SLAFB_FMUL(kf, ln2_lo), c)), f));
    }
    const float s = SLAFB_FDIV(f, SLAFB_FADD(2.0f, f));
    const float z = SLAFB_FMUL(s, s);
    float R = SLAFB_FADD(Lp6, SLAFB_FMUL(z, Lp7));
    R = SLAFB_FADD(Lp5, SLAFB_FMUL(z, R));
    R = SLAFB_FADD(Lp4, SLAFB_FMUL(z, R));
    R = SLAFB_FADD(Lp3, SLAFB_FMUL(z, R));
    R = SLAFB_FADD(Lp2, SLAFB_FMUL(z, R));
    R = SLAFB_FADD(Lp1, SLAFB_FMUL(z, R));
    R = SLAFB_FMUL(z, R);
    if (k == 0) return SLAFB_FSUB(f, SLAFB_FSUB(hfsq, SLAFB_FMUL(s, SLAFB_FADD(hfsq, R))));
    return SLAFB_FSUB(SLAFB_FMUL(kf, ln2_hi),
                     SLAFB_FSUB(SLAFB_FSUB(hfsq, SLAFB_FADD(SLAFB_FMUL(s, SLAFB_FADD(hfsq, R)), SLAFB_FADD(SLAFB_FMUL(kf, ln2_lo), c))), f));
}
// (out of line on purpose: it is called from up to five places of a kernel, a few times per cell, and both evaluations are ~300
// instructions -- inlined they doubled the float32 kernels' code and cost 15 % through the instruction cache)
__device__ __noinline__ float log1pf_window(float x) { return c_log1pf_fdlibm ? log1pf_fdlibm(x) : (float)log1p((double)x); }

// ------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy primitives (PTX)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared (1-D, 16-byte aligned, size a multiple of 16), completion on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait_parity(uint64_t *bar, uint32_t parity) { mbar_wait(bar, parity); }

// ------------------------------------------------------------------------------------------
// K1: segmented CSR build from the COO cell column
// ------------------------------------------------------------------------------------------
// Each CTA owns one contiguous span of rows.  A producer warp streams the span through a ring of
// shared-memory stages with TMA bulk copies (kCsrStages x 8 KiB in flight per CTA, ~19 MB over the
// grid: the loaded DRAM latency is ~2.5 us, registers could not hold that much), eight consumer
// warps test 8 rows per thread for a change of cell id (2 x LDS.128 + xor/or) and append the rare
// segment heads to a per-CTA list.  Heads are row numbers, so sorting the (tiny) list restores row
// order; a flat gather of the preceding CTAs' counts gives the global offset.
struct CsrArgs {
    const uint32_t *cell;   // uint32 or int32 bits
    uint32_t n_rows;
    uint32_t max_cells;
    uint32_t rows_per_cta;  // multiple of kCsrStageRows
    uint32_t *seg_start;    // [max_cells + 1]
    uint32_t *seg_cell;     // [max_cells]
    unsigned long long *tile_state;  // [grid] : epoch<<34 | status<<32 | value
    uint32_t *ticket;       // dynamic span counter (never reset; host passes the base)
    uint32_t ticket_base;
    uint32_t epoch;         // 30-bit launch id, so tile_state needs no memset
    uint32_t *n_cells_out;  // [0] = number of segments, [1] = error flags
    uint32_t *host_counts;  // the same two words in mapped pinned host memory (no D2H copy in the stream)
    uint32_t *work_counters;  // [0] general-path count, [1] general-path head, [2] K2 row ticket (reset here)
};

constexpr unsigned long long kStAgg = 1ull;
__device__ __forceinline__ unsigned long long st_pack(uint32_t epoch, unsigned long long status, uint32_t v) {
    return ((unsigned long long)epoch << 34) | (status << 32) | (unsigned long long)v;
}

constexpr int kCsrConsumers = 256;                 // 8 consumer warps
constexpr int kCsrBlock = kCsrConsumers + 32;      // + 1 producer warp
constexpr int kCsrStageRows = kCsrConsumers * 8;   // 2048 rows = 8 KiB per stage
constexpr int kCsrStages = 4;

// The part of K1 after the scan: order the span's heads, publish the count, gather the counts of all preceding spans,
// write the span's rows of the segment table.  Shared by the TMA-fed and the direct-load variant of the kernel.
template <int BLOCK>
__device__ __forceinline__ void csr_finish(const CsrArgs &a, uint32_t span, uint32_t r_beg, uint32_t r_end, uint32_t *s_list,
                                           uint32_t *s_sorted, uint32_t &s_count, uint32_t &s_base, uint32_t *s_ws, uint32_t *s_part) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t R = a.n_rows;
    constexpr int kCsrBlock = BLOCK;          // (shadows the namespace constant inside this function)
    constexpr int kCsrConsumers = 256;
    // ---------------- order the heads of this span
    const uint32_t count = s_count;
    const bool listed = count <= (uint32_t)kCsrListCap;
    if (listed) {
        for (uint32_t i = tid; i < count; i += kCsrBlock) {
            const uint32_t v = s_list[i];
            uint32_t rank = 0;
            for (uint32_t k = 0; k < count; k++) rank += (s_list[k] < v) ? 1u : 0u;
            s_sorted[rank] = v;
        }
    }

    // ---------------- publish the count, gather the counts of all preceding CTAs
    // (a few hundred CTAs that finish at about the same time: a flat gather is one or two L2 round
    // trips, a chained look-back would serialise them.  Spinning is safe: spans are handed out by
    // ticket, so every predecessor has already started.)
    {
        volatile unsigned long long *stt = a.tile_state;
        if (tid == 0) stt[span] = st_pack(a.epoch, kStAgg, count);
        uint32_t part = 0;
        for (uint32_t i = tid; i < span; i += kCsrBlock) {
            unsigned long long sv;
            do { sv = stt[i]; } while ((uint32_t)(sv >> 34) != a.epoch || ((sv >> 32) & 3ull) == 0ull);
            part += (uint32_t)sv;
        }
        part = warp_sum(part);
        if (lane == 0) s_part[warp] = part;
        __syncthreads();
        if (tid == 0) {
            uint32_t excl = 0;
#pragma unroll
            for (int w = 0; w < kCsrBlock / 32; w++) excl += s_part[w];
            s_base = excl;
            if (span == gridDim.x - 1) {
                const uint32_t total = excl + count;
                a.n_cells_out[0] = total;
                a.n_cells_out[1] = (total > a.max_cells) ? 1u : 0u;
                a.host_counts[0] = total;
                a.host_counts[1] = (total > a.max_cells) ? 1u : 0u;
                if (total <= a.max_cells) a.seg_start[total] = R;
                a.work_counters[0] = 0u;
                a.work_counters[1] = 0u;
                a.work_counters[2] = 0u;   // K2 row ticket
            }
        }
        __syncthreads();
    }
    const uint32_t base = s_base;

    // ---------------- write the segment table rows of this span
    if (listed) {
        for (uint32_t i = tid; i < count; i += kCsrBlock) {
            const uint32_t o = base + i;
            if (o < a.max_cells) {
                const uint32_t r = s_sorted[i];
                SLAFB_CHECK(r < R, 101, r);
                a.seg_start[o] = r;
                a.seg_cell[o] = __ldg(a.cell + r);
            }
        }
    } else {
        // more heads than the list holds (very short cells): replay the span in row order with a
        // block-wide scan per 2048-row tile.  Correct for any input, only slower.
        uint32_t run = base;
        for (uint32_t t0 = r_beg; t0 < r_end; t0 += kCsrStageRows) {
            uint32_t flags = 0, c[8];
            const uint32_t r0 = t0 + (uint32_t)tid * 8u;
            if (tid < kCsrConsumers) {
                uint32_t p = (r0 > 0u && r0 < r_end) ? __ldg(a.cell + r0 - 1) : 0u;
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const uint32_t r = r0 + e;
                    c[e] = (r < r_end) ? __ldg(a.cell + r) : 0u;
                    if (r < r_end && (r == 0u || c[e] != p)) flags |= 1u << e;
                    p = c[e];
                }
            }
            uint32_t tot;
            uint32_t off = run + block_excl_scan(__popc(flags), s_ws, tot);
#pragma unroll
            for (int e = 0; e < 8; e++) {
                if (flags & (1u << e)) {
                    if (off < a.max_cells) { a.seg_start[off] = r0 + e; a.seg_cell[off] = c[e]; }
                    off++;
                }
            }
            run += tot;
        }
    }
}

__global__ void __launch_bounds__(kCsrBlock) csr_build_kernel(const CsrArgs a) {
    __shared__ __align__(128) uint32_t s_stage[kCsrStages][kCsrStageRows + 4];   // [0..3] = the 4 rows before the tile
    __shared__ __align__(8) uint64_t s_full[kCsrStages], s_empty[kCsrStages];
    __shared__ uint32_t s_list[kCsrListCap], s_sorted[kCsrListCap];
    __shared__ uint32_t s_count, s_span, s_base, s_ws[16], s_part[kCsrBlock / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_span = atomicAdd(a.ticket, 1u) - a.ticket_base;
        s_count = 0u;
        for (int i = 0; i < kCsrStages; i++) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], kCsrConsumers / 32); }
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t span = s_span, R = a.n_rows;
    const unsigned long long sb = (unsigned long long)span * a.rows_per_cta;
    const uint32_t r_beg = sb < R ? (uint32_t)sb : R;
    const uint32_t r_end = (sb + a.rows_per_cta < R) ? (uint32_t)(sb + a.rows_per_cta) : R;
    const uint32_t n_tiles = (r_end - r_beg + kCsrStageRows - 1) / kCsrStageRows;

    if (warp == kCsrConsumers / 32) {
        // ---------------- producer warp: keep kCsrStages bulk copies in flight
        if (lane == 0) {
            for (uint32_t t = 0; t < n_tiles; t++) {
                const uint32_t st = t % kCsrStages;
                if (t >= (uint32_t)kCsrStages) mbar_wait(&s_empty[st], ((t / kCsrStages) - 1u) & 1u);
                const uint32_t t0 = r_beg + t * kCsrStageRows;
                const uint32_t rows = min((uint32_t)kCsrStageRows, r_end - t0) & ~3u;   // 16-byte granules; a ragged tail is read directly
                const uint32_t lead = t0 >= 4u ? 4u : 0u;                               // the row before the tile, for the first compare
                const uint32_t bytes = (rows + lead) * 4u;
                if (bytes) {
                    mbar_arrive_expect_tx(&s_full[st], bytes);
                    bulk_g2s(&s_stage[st][4u - lead], a.cell + t0 - lead, bytes, &s_full[st]);
                } else {
                    mbar_arrive(&s_full[st]);
                }
            }
        }
    } else {
        // ---------------- consumer warps
        for (uint32_t t = 0; t < n_tiles; t++) {
            const uint32_t st = t % kCsrStages;
            mbar_wait(&s_full[st], (t / kCsrStages) & 1u);
            const uint32_t t0 = r_beg + t * kCsrStageRows;
            const uint32_t rows = min((uint32_t)kCsrStageRows, r_end - t0);
            const uint32_t rows4 = rows & ~3u;
            const uint32_t i0 = (uint32_t)tid * 8u;          // first of my 8 rows inside the tile
            if (i0 < rows) {
                uint32_t c[8], prev;
                const uint32_t *sp = &s_stage[st][4u + i0];
                if (i0 + 8u <= rows4) {
                    const uint4 q0 = *reinterpret_cast<const uint4 *>(sp), q1 = *reinterpret_cast<const uint4 *>(sp + 4);
                    c[0] = q0.x; c[1] = q0.y; c[2] = q0.z; c[3] = q0.w; c[4] = q1.x; c[5] = q1.y; c[6] = q1.z; c[7] = q1.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const uint32_t i = i0 + e;
                        c[e] = (i < rows4) ? sp[e] : ((i < rows) ? __ldg(a.cell + t0 + i) : 0u);
                    }
                }
                const uint32_t r0 = t0 + i0;
                if (r0 == 0u) prev = ~c[0];                                  // row 0 always opens a segment
                else prev = sp[-1];                                          // previous row (lead rows for i0 == 0)
                uint32_t diff = c[0] ^ prev;
#pragma unroll
                for (int e = 1; e < 8; e++) diff |= c[e] ^ c[e - 1];
                if (diff) {                                                   // rare: ~1 segment head per 2000 rows
                    uint32_t p = prev;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        if (i0 + e < rows && c[e] != p) {
                            const uint32_t k = atomicAdd(&s_count, 1u);
                            if (k < (uint32_t)kCsrListCap) s_list[k] = r0 + e;
                        }
                        p = c[e];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[st]);
        }
    }
    __syncthreads();

    csr_finish<kCsrBlock>(a, span, r_beg, r_end, s_list, s_sorted, s_count, s_base, s_ws, s_part);
}

// K1, direct-load variant: the same spans, heads and ordering, but the rows come through LDG.128 instead of the bulk-copy
// engine (cp.async.bulk moves 16 B per clock per SM = 4.65 TB/s over the chip; plain coalesced loads reach ~5.1 TB/s for
// this span pattern, profiles/r1i_read_probe.txt).  256 threads, a tile of 2048 rows per iteration: warp w, lane l holds rows
// [w*256 + l*4, +4) and [w*256 + 128 + l*4, +4) -- every load instruction of a warp covers 512 contiguous bytes -- and the
// next tile is already in flight while the current one is tested.
constexpr int kCsrLdgBlock = 256;

__global__ void __launch_bounds__(kCsrLdgBlock, 8) csr_build_ldg_kernel(const CsrArgs a) {
    __shared__ uint32_t s_list[kCsrListCap], s_sorted[kCsrListCap];
    __shared__ uint32_t s_count, s_span, s_base, s_ws[16], s_part[kCsrLdgBlock / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_span = atomicAdd(a.ticket, 1u) - a.ticket_base;
        s_count = 0u;
    }
    __syncthreads();
    const uint32_t span = s_span, R = a.n_rows;
    const unsigned long long sb = (unsigned long long)span * a.rows_per_cta;
    const uint32_t r_beg = sb < R ? (uint32_t)sb : R;
    const uint32_t r_end = (sb + a.rows_per_cta < R) ? (uint32_t)(sb + a.rows_per_cta) : R;
    const uint32_t n_tiles = (r_end - r_beg + kCsrStageRows - 1) / kCsrStageRows;
    const uint32_t in_tile = (uint32_t)warp * 256u + (uint32_t)lane * 4u;          // first row of my first group inside a tile

    auto load_tile = [&](uint32_t t, uint4 &q0, uint4 &q1, uint32_t &before) {
        const uint32_t t0 = r_beg + t * kCsrStageRows, g0 = t0 + in_tile, g1 = g0 + 128u;
        if (g1 + 4u <= r_end) {                                       // both groups inside the span: two aligned 16-byte loads
            q0 = __ldg(reinterpret_cast<const uint4 *>(a.cell + g0));
            q1 = __ldg(reinterpret_cast<const uint4 *>(a.cell + g1));
        } else {                                                      // ragged end of the table: element-wise, zero beyond
            uint32_t c[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const uint32_t r = (e < 4 ? g0 : g1) + (uint32_t)(e & 3);
                c[e] = r < r_end ? __ldg(a.cell + r) : 0u;
            }
            q0 = make_uint4(c[0], c[1], c[2], c[3]);
            q1 = make_uint4(c[4], c[5], c[6], c[7]);
        }
        before = (lane == 0 && g0 > 0u && g0 < r_end) ? __ldg(a.cell + g0 - 1) : 0u;   // the row before the warp's 256 rows
    };
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0, n0 = q0, n1 = q0;
    uint32_t before = 0, nbefore = 0;
    if (n_tiles) load_tile(0, q0, q1, before);
    for (uint32_t t = 0; t < n_tiles; t++) {
        if (t + 1 < n_tiles) load_tile(t + 1, n0, n1, nbefore);
        const uint32_t t0 = r_beg + t * kCsrStageRows, g0 = t0 + in_tile, g1 = g0 + 128u;
        // predecessor of each group's first row: the previous lane's last element (lane 0: the row before / lane 31's group 0)
        uint32_t p0 = __shfl_up_sync(0xffffffffu, q0.w, 1), p1 = __shfl_up_sync(0xffffffffu, q1.w, 1);
        const uint32_t last0 = __shfl_sync(0xffffffffu, q0.w, 31);
        if (lane == 0) { p0 = (g0 == 0u) ? ~q0.x : before; p1 = last0; }
        const uint32_t d0 = (q0.x ^ p0) | (q0.y ^ q0.x) | (q0.z ^ q0.y) | (q0.w ^ q0.z);
        const uint32_t d1 = (q1.x ^ p1) | (q1.y ^ q1.x) | (q1.z ^ q1.y) | (q1.w ^ q1.z);
        if ((d0 && g0 < r_end) || (d1 && g1 < r_end)) {                // rare: ~1 segment head per 2000 rows
            const uint32_t c[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const uint32_t r = (e < 4 ? g0 : g1) + (uint32_t)(e & 3);
                const uint32_t prev = (e == 0) ? p0 : (e == 4 ? p1 : c[e - 1]);
                if (r < r_end && c[e] != prev) {
                    const uint32_t k = atomicAdd(&s_count, 1u);
                    if (k < (uint32_t)kCsrListCap) s_list[k] = r;
                }
            }
        }
        q0 = n0; q1 = n1; before = nbefore;
    }
    __syncthreads();
    csr_finish<kCsrLdgBlock>(a, span, r_beg, r_end, s_list, s_sorted, s_count, s_base, s_ws, s_part);
}

// ------------------------------------------------------------------------------------------
// K2: tokenization
// ------------------------------------------------------------------------------------------
struct TokArgs {
    const void *gene;
    const void *value;
    uint32_t n_rows;
    const uint32_t *seg_start;
    const uint32_t *seg_cell;
    const int32_t *perm;          // nullable: output row -> segment
    uint32_t n_cells;
    int cell_signed;              // cell ids are int32 (sign-extend into int64 cell_ids)
    long long *ids;
    uint8_t *mask;
    long long *vals;
    long long *cell_ids;
    uint32_t max_items;
    uint32_t max_genes;
    uint32_t L;
    int n_bins;                   // tokenizer bins (ScGPTTokenizer.n_expression_bins)
    int w_bins;                   // window bins (ScGPTWindow n_expression_bins); normally == n_bins
    long long vocab;
    double min_pct;
    float bin_size;               // float32(1.0 / n_bins)
    const double *log1p_lut;      // [65536] glibc log1p(double(v)), uploaded once
    uint32_t *work_list;          // output rows deferred to the general kernel
    uint32_t *work_counters;      // [0] count, [1] head
    uint32_t *host_slow;          // mapped pinned host word: number of cells the general kernel handled
    uint32_t *sort_scratch;       // [n_rows] keys, general path with > kSortSmemKeys float rows
    int all_cells;                // general kernel: iterate every output row (percentile mode)
};

struct TokSmem {
    uint32_t stage_gene[kChunkRows];
    uint32_t stage_val[kChunkRows];
    uint32_t ws[32];
    uint32_t bcast[4];
};

// zero-fill int64 elements [A, B) of a flat array (8-byte aligned base), 16-byte stores inside
__device__ __forceinline__ void fill_zero_i64(long long *base, size_t A, size_t B) {
    if (A >= B) return;
    size_t a2 = (A + 1) & ~(size_t)1, b2 = B & ~(size_t)1;
    if (threadIdx.x == 0 && (A & 1)) base[A] = 0;
    if (threadIdx.x == 1 && (B & 1) && B - 1 >= A) base[B - 1] = 0;
    if (a2 < b2) {
        ulonglong2 *v = reinterpret_cast<ulonglong2 *>(base + a2);
        const size_t nv = (b2 - a2) >> 1;
        for (size_t i = threadIdx.x; i < nv; i += blockDim.x) v[i] = make_ulonglong2(0ull, 0ull);
    }
}

// attention mask bytes for flat range [A, A+L): 1 for the first t_len, then 0; 16-byte stores inside.  The work is spread over
// `nth` threads of which the caller is number `me` (a whole CTA, or one warp).
__device__ __forceinline__ void fill_mask_by(uint8_t *base, size_t A, uint32_t L, uint32_t t_len, uint32_t me, uint32_t nth) {
    const uint32_t lead = (uint32_t)((16u - (uint32_t)(A & 15)) & 15u);     // bytes before the first aligned chunk
    // ragged head and tail: byte stores
    const uint32_t head = lead < L ? lead : L;
    const uint32_t body = (L - head) & ~15u;
    const uint32_t tail0 = head + body;
    for (uint32_t i = me; i < head + (L - tail0); i += nth) {
        const uint32_t t = i < head ? i : tail0 + (i - head);
        base[A + t] = (t < t_len) ? 1 : 0;
    }
    uint4 *chunks = reinterpret_cast<uint4 *>(base + A + head);
    const uint32_t nchunks = body >> 4;
    for (uint32_t c = me; c < nchunks; c += nth) {
        const uint32_t t = head + (c << 4);                                   // first token of the chunk
        uint4 v;
        if (t + 16u <= t_len) v = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
        else if (t >= t_len) v = make_uint4(0u, 0u, 0u, 0u);
        else {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t x = 0;
#pragma unroll
                for (int b2 = 0; b2 < 4; b2++) x |= (t + q * 4 + b2 < t_len ? 1u : 0u) << (8 * b2);
                w[q] = x;
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        chunks[c] = v;
    }
}
__device__ __forceinline__ void fill_mask(uint8_t *base, size_t A, uint32_t L, uint32_t t_len) {
    fill_mask_by(base, A, L, t_len, threadIdx.x, blockDim.x);
}

// value stream token for one kept row (scGPT modes)
template <typename ValT, int MODE>
__device__ __forceinline__ long long value_token(uint32_t raw, const TokArgs &a, double m_f64, float m_f32) {
    if constexpr (MODE == MODE_SCGPT_RAW) {
        if constexpr (sizeof(ValT) == 2) {
            return (long long)raw + a.vocab;  // isinstance(int) branch, tokenizers.py:441-444
        } else {
            return expr_bin_token(__uint_as_float(raw), a.bin_size, a.n_bins, a.vocab);
        }
    } else {
        float bin;
        if constexpr (sizeof(ValT) == 2) {
            const double lv = __ldg(a.log1p_lut + raw);
            double b = 0.0;
            if (lv > 0.0) {
                b = floor(__ddiv_rn(__dmul_rn(lv, (double)a.w_bins), m_f64));
                b = fmin(fmax(b, 0.0), (double)(a.w_bins - 1));
            }
            bin = (float)b;
        } else {
            const float lv = log1pf_window(__uint_as_float(raw));
            float b = 0.0f;
            if (lv > 0.0f) {
                b = floorf(__fdiv_rn(__fmul_rn(lv, (float)a.w_bins), m_f32));
                b = fminf(fmaxf(b, 0.0f), (float)(a.w_bins - 1));
            }
            bin = b;
        }
        return expr_bin_token(bin, a.bin_size, a.n_bins, a.vocab);  // float branch applied to the bin
    }
}

// expr_bin of the binned window as double (facade output)
template <typename ValT>
__device__ __forceinline__ double window_bin(uint32_t raw, const TokArgs &a, double m_f64, float m_f32) {
    if constexpr (sizeof(ValT) == 2) {
        const double lv = __ldg(a.log1p_lut + raw);
        double b = 0.0;
        if (lv > 0.0) {
            b = floor(__ddiv_rn(__dmul_rn(lv, (double)a.w_bins), m_f64));
            b = fmin(fmax(b, 0.0), (double)(a.w_bins - 1));
        }
        return b;
    } else {
        const float lv = log1pf_window(__uint_as_float(raw));
        float b = 0.0f;
        if (lv > 0.0f) {
            b = floorf(__fdiv_rn(__fmul_rn(lv, (float)a.w_bins), m_f32));
            b = fminf(fmaxf(b, 0.0f), (float)(a.w_bins - 1));
        }
        return (double)b;
    }
}

// ---- exact threshold machinery (general path) ------------------------------------------------
// Target ascending dense rank a* in [1, D]: keep rows whose ascending rank >= a*.
//   plain:       a* = D - min(k, D) + 1
//   percentile:  additionally ascending rank >= ceil(p * D / 100)   (f64 compare, aggregators.py:184-188)
__device__ __forceinline__ uint32_t target_asc_rank(uint32_t D, uint32_t k, bool pct, double p) {
    const uint32_t q = k < D ? k : D;
    uint32_t astar = D - q + 1;
    if (pct) {
        const double t = __ddiv_rn(__dmul_rn(p, (double)D), 100.0);
        double ct = ceil(t);
        uint32_t ap = ct < 1.0 ? 1u : (ct > 4294967295.0 ? 0xffffffffu : (uint32_t)ct);
        if (ap > astar) astar = ap;
    }
    return astar;
}

// uint16 values: presence bitmap over [lo, hi]; returns the key of the a*-th distinct value (ascending)
template <typename ValT>
__device__ uint32_t threshold_bitmap(const ValT *__restrict__ value, uint32_t s, uint32_t e, uint32_t R,
                                     uint32_t lo, uint32_t hi, uint32_t k, bool pct, double p,
                                     uint32_t *bitmap, uint32_t *ws, uint32_t *bcast) {
    const int tid = threadIdx.x;
    const uint32_t w0 = lo >> 5, w1 = hi >> 5;
    for (uint32_t w = w0 + tid; w <= w1; w += kTokThreads) bitmap[w] = 0u;
    __syncthreads();
    for (uint32_t g0 = s & ~7u; g0 < e; g0 += kChunkRows) {
        const uint32_t r0 = g0 + tid * kGroup;
        if (r0 < e) {
            ValT v[8];
            load8(value, r0, R, v);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t r = r0 + i;
                if (r >= s && r < e) {
                    const uint32_t key = key_of(v[i]);
                    const uint32_t bit = 1u << (key & 31);
                    // counts repeat: after the first rows nearly every bit is already set, so look before the atomic
                    if (!(reinterpret_cast<volatile uint32_t *>(bitmap)[key >> 5] & bit)) atomicOr(&bitmap[key >> 5], bit);
                }
            }
        }
    }
    __syncthreads();
    // distinct count
    uint32_t cnt = 0;
    for (uint32_t w = w0 + tid; w <= w1; w += kTokThreads) cnt += __popc(bitmap[w]);
    uint32_t D;
    block_excl_scan(cnt, ws, D);
    const uint32_t astar = target_asc_rank(D, k, pct, p);
    // locate the a*-th set bit from the bottom
    uint32_t run = 0;
    for (uint32_t wb = w0; wb <= w1; wb += kTokThreads) {
        const uint32_t w = wb + tid;
        const uint32_t word = (w <= w1) ? bitmap[w] : 0u;
        const uint32_t pc = __popc(word);
        uint32_t tot;
        const uint32_t ex = run + block_excl_scan(pc, ws, tot);
        if (astar > ex && astar <= ex + pc) {
            uint32_t x = word;
            for (uint32_t i = ex + 1; i < astar; i++) x &= x - 1;  // drop lower set bits
            bcast[0] = (w << 5) + (__ffs(x) - 1);
        }
        run += tot;
        if (run >= astar) break;
    }
    __syncthreads();
    const uint32_t thr = bcast[0];
    __syncthreads();
    return thr;
}

// float32 values: bitonic sort of the keys (shared memory, or global scratch when the cell is
// larger than kSortSmemKeys), then the a*-th distinct key.  All compare-exchanges are ascending
// (mirrored first step), so the virtual +inf padding above n never moves.
template <typename KeyT>
__device__ void bitonic_sort_asc_t(KeyT *keys, uint32_t n) {
    const int tid = threadIdx.x;
    uint32_t P = 1;
    while (P < n) P <<= 1;
    const uint32_t half = P >> 1;
    // all strides are powers of two: index arithmetic with shifts and masks only
    for (uint32_t k = 2, lgk = 1; k <= P; k <<= 1, lgk++) {
        const uint32_t hk = k >> 1, lgh = lgk - 1;
        for (uint32_t t = tid; t < half; t += kTokThreads) {
            const uint32_t i = t & (hk - 1u);
            const uint32_t base = (t >> lgh) << lgk;
            const uint32_t lo = base | i, hi = base | ((k - 1u) ^ i);          // mirrored first step: i <-> k-1-i
            if (hi < n) {
                const KeyT x = keys[lo], y = keys[hi];
                if (x > y) { keys[lo] = y; keys[hi] = x; }
            }
        }
        __syncthreads();
        for (uint32_t j = hk >> 1, lgj = lgh - 1u; j >= 1; j >>= 1, lgj--) {
            for (uint32_t t = tid; t < half; t += kTokThreads) {
                const uint32_t lo = ((t >> lgj) << (lgj + 1u)) | (t & (j - 1u)), hi = lo | j;
                if (hi < n) {
                    const KeyT x = keys[lo], y = keys[hi];
                    if (x > y) { keys[lo] = y; keys[hi] = x; }
                }
            }
            __syncthreads();
        }
    }
}
__device__ void bitonic_sort_asc(uint32_t *keys, uint32_t n) { bitonic_sort_asc_t<uint32_t>(keys, n); }

template <typename ValT>
__device__ uint32_t threshold_sort(const ValT *__restrict__ value, uint32_t s, uint32_t e, uint32_t R,
                                   uint32_t k, bool pct, double p, uint32_t *keys, uint32_t *ws,
                                   uint32_t *bcast) {
    const int tid = threadIdx.x;
    const uint32_t n = e - s;
    for (uint32_t i = tid; i < n; i += kTokThreads) keys[i] = key_of(__ldg(value + s + i));
    __syncthreads();
    bitonic_sort_asc(keys, n);
    // distinct count
    uint32_t cnt = 0;
    for (uint32_t i = tid; i < n; i += kTokThreads) cnt += (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
    uint32_t D;
    block_excl_scan(cnt, ws, D);
    const uint32_t astar = target_asc_rank(D, k, pct, p);
    uint32_t run = 0;
    for (uint32_t ib = 0; ib < n; ib += kTokThreads) {
        const uint32_t i = ib + tid;
        const uint32_t head = (i < n && (i == 0 || keys[i] != keys[i - 1])) ? 1u : 0u;
        uint32_t tot;
        const uint32_t ex = run + block_excl_scan(head, ws, tot);
        if (head && ex + 1 == astar) bcast[0] = keys[i];
        run += tot;
        if (run >= astar) break;
    }
    __syncthreads();
    const uint32_t thr = bcast[0];
    __syncthreads();
    return thr;
}

// ---- one cell ------------------------------------------------------------------------------
// SINK: 0 = token tensors (K2/K2g), 1 = count only (window pass A), 2 = flat lists (window pass B)
struct WindowSink {
    uint32_t *kept_count;      // [C]   pass A
    const long long *offsets;  // [C+1] pass B (exclusive scan of kept_count)
    long long *gene_list;
    double *expr_list;
};

template <typename GeneT, typename ValT, int MODE, bool GENERAL, int SINK>
__device__ bool process_cell(const TokArgs &a, const WindowSink &wsnk, uint32_t out_row, TokSmem &sm,
                             uint32_t *bitmap, uint32_t *sort_smem) {
    constexpr bool SCGPT = (MODE == MODE_SCGPT_RAW || MODE == MODE_SCGPT_BINNED);
    constexpr bool BINNED = (MODE == MODE_SCGPT_BINNED);
    constexpr bool PCT = (MODE == MODE_GF_PCT);
    constexpr bool U16 = (sizeof(ValT) == 2);
    const int tid = threadIdx.x;
    const GeneT *__restrict__ gene = static_cast<const GeneT *>(a.gene);
    const ValT *__restrict__ value = static_cast<const ValT *>(a.value);
    const uint32_t c = a.perm ? (uint32_t)a.perm[out_row] : out_row;
    const uint32_t s = __ldg(a.seg_start + c), e = __ldg(a.seg_start + c + 1);
    const uint32_t n = e - s, R = a.n_rows, k = a.max_items;

    // ---- pass 1: statistics the filter / binning needs ------------------------------------
    bool keep_all = (n <= k) && !PCT;
    uint32_t thr = 0;
    uint32_t kmax = 0;
    float m_f32 = 0.0f;
    if (BINNED || !keep_all) {
        uint32_t lo = 0xffffffffu, hi = 0u;
        for (uint32_t g0 = s & ~7u; g0 < e; g0 += kChunkRows) {
            const uint32_t r0 = g0 + tid * kGroup;
            if (r0 < e) {
                ValT v[8];
                load8(value, r0, R, v);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint32_t r = r0 + i;
                    if (r >= s && r < e) {
                        const uint32_t key = key_of(v[i]);
                        lo = min(lo, key);
                        hi = max(hi, key);
                    }
                }
            }
        }
        block_minmax(lo, hi, sm.ws);
        kmax = hi;
        // log1p is monotone: max(log1pf(x)) over the cell = log1pf(max x)
        if constexpr (BINNED && !U16) m_f32 = log1pf_window(float_of_key(hi));
        if (!keep_all) {
            if (!PCT && U16 && (hi - lo + 1u <= k)) {
                keep_all = true;  // at most k distinct values can exist in [lo, hi]
            } else {
                if constexpr (!GENERAL) {
                    if (tid == 0) a.work_list[atomicAdd(a.work_counters, 1u)] = out_row;
                    return false;
                } else {
                    if constexpr (U16) {
                        thr = threshold_bitmap<ValT>(value, s, e, R, lo, hi, k, PCT, a.min_pct, bitmap, sm.ws, sm.bcast);
                    } else {
                        uint32_t *keys = (n <= (uint32_t)kSortSmemKeys) ? sort_smem : (a.sort_scratch + s);
                        thr = threshold_sort<ValT>(value, s, e, R, k, PCT, a.min_pct, keys, sm.ws, sm.bcast);
                    }
                }
            }
        }
    }
    double m_f64 = 0.0;
    if constexpr (BINNED && U16) m_f64 = __ldg(a.log1p_lut + kmax);
    // binned window: the max is taken over KEPT rows (aggregators.py:103); the cell maximum always
    // has dense rank 1 <= max_items, so it equals the max over all rows of the cell.

    // ---- pass 2: row-order compaction + emission ------------------------------------------
    const uint32_t L = a.L;
    const uint32_t limit = (SINK == 0) ? (SCGPT ? a.max_genes : (L - 1u)) : 0xffffffffu;
    const size_t row_base = (size_t)out_row * L;
    const long long list_base = (SINK == 2) ? wsnk.offsets[out_row] : 0;
    uint32_t base = 0;
    for (uint32_t g0 = s & ~7u; g0 < e && base < limit; g0 += kChunkRows) {
        const uint32_t r0 = g0 + tid * kGroup;
        GeneT g[8];
        ValT v[8];
        uint32_t flags = 0;
        if (r0 < e) {
            if constexpr (SINK != 1) load8(gene, r0, R, g);
            if (SCGPT || !keep_all) load8(value, r0, R, v);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t r = r0 + i;
                bool kp = (r >= s && r < e);
                if (!keep_all) kp = kp && (key_of(v[i]) >= thr);
                flags |= (kp ? 1u : 0u) << i;
            }
        }
        uint32_t chunk_total;
        uint32_t off = block_excl_scan(__popc(flags), sm.ws, chunk_total);
        if constexpr (SINK != 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (flags & (1u << i)) {
                    SLAFB_CHECK(off < (uint32_t)kChunkRows, 301, off);
                    sm.stage_gene[off] = (uint32_t)g[i];
                    if constexpr (SCGPT) sm.stage_val[off] = raw_bits(v[i]);
                    off++;
                }
            }
            __syncthreads();
            const uint32_t m = min(chunk_total, limit - base);
            if constexpr (SINK == 0) {
                long long *ids = a.ids + row_base + 1 + base;
                SLAFB_CHECK(out_row < a.n_cells && 1u + base + m <= L, 302, base + m);
                for (uint32_t i = tid; i < m; i += kTokThreads) {
                    long long gid;
                    if constexpr (sizeof(GeneT) == 2) gid = (long long)sm.stage_gene[i];
                    else gid = (long long)(int32_t)sm.stage_gene[i];
                    ids[i] = gid + 4;                               // tokenizers.py:439,686
                    if constexpr (SCGPT) a.vals[row_base + 1 + base + i] = value_token<ValT, MODE>(sm.stage_val[i], a, m_f64, m_f32);
                }
            } else {
                for (uint32_t i = tid; i < m; i += kTokThreads) {
                    long long gid;
                    if constexpr (sizeof(GeneT) == 2) gid = (long long)sm.stage_gene[i];
                    else gid = (long long)(int32_t)sm.stage_gene[i];
                    wsnk.gene_list[list_base + base + i] = gid;
                    if constexpr (SCGPT) {
                        double x;
                        if constexpr (BINNED) x = window_bin<ValT>(sm.stage_val[i], a, m_f64, m_f32);
                        else if constexpr (U16) x = (double)sm.stage_val[i];
                        else x = (double)__uint_as_float(sm.stage_val[i]);
                        wsnk.expr_list[list_base + base + i] = x;
                    }
                }
            }
            __syncthreads();
        }
        base += chunk_total;
    }

    if constexpr (SINK == 1) {
        if (tid == 0) wsnk.kept_count[out_row] = base;
        return true;
    } else if constexpr (SINK == 2) {
        return true;
    } else {
        // ---- frame: CLS / SEP / PAD, value-stream PADs, attention mask, cell id -------------
        const uint32_t n_emit = min(base, limit);
        const uint32_t sep_pos = 1u + n_emit;
        const uint32_t t_len = min(n_emit + 2u, L);
        if (tid == 0) {
            a.ids[row_base] = 1;                                   // CLS
            if (sep_pos < L) a.ids[row_base + sep_pos] = 2;        // SEP (dropped when truncated, tokenizers.py:706)
            if constexpr (SCGPT) a.vals[row_base] = 0;             // PAD under CLS
            const uint32_t cid = __ldg(a.seg_cell + c);
            a.cell_ids[out_row] = a.cell_signed ? (long long)(int32_t)cid : (long long)cid;
        }
        fill_zero_i64(a.ids, row_base + t_len, row_base + L);
        if constexpr (SCGPT) fill_zero_i64(a.vals, row_base + sep_pos, row_base + L);
        fill_mask(a.mask, row_base, L, t_len);
        return true;
    }
}

// ---- K2 fast path: persistent CTAs, TMA-staged cell segments ------------------------------------
// min / max of the uint16 values of rows [t0, e) of the column (the rows of a long cell beyond its value stage), read by the whole
// CTA from global memory: 128-bit loads over the aligned middle, single rows at the ragged ends
__device__ __forceinline__ void minmax_rows_u16(const uint16_t *__restrict__ v, uint32_t t0, uint32_t e, uint32_t &lo, uint32_t &hi) {
    if (t0 >= e) return;
    const uint32_t m0 = min((t0 + 7u) & ~7u, e), m1 = max(m0, e & ~7u);
    uint32_t mn2 = 0xffffffffu, mx2 = 0u;
    for (uint32_t g = (m0 >> 3) + threadIdx.x; g < (m1 >> 3); g += kFastThreads) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(v) + g);
        mn2 = __vminu2(__vminu2(mn2, __vminu2(q.x, q.y)), __vminu2(q.z, q.w));
        mx2 = __vmaxu2(__vmaxu2(mx2, __vmaxu2(q.x, q.y)), __vmaxu2(q.z, q.w));
    }
    const uint32_t nh = m0 - t0, nt = e - m1;
    if (threadIdx.x < nh + nt) {
        const uint32_t r = threadIdx.x < nh ? t0 + threadIdx.x : m1 + (threadIdx.x - nh);
        const uint32_t x = (uint32_t)__ldg(v + r);
        mn2 = __vminu2(mn2, x | (x << 16));
        mx2 = __vmaxu2(mx2, x | (x << 16));
    }
    lo = min(lo, min(mn2 & 0xffffu, mn2 >> 16));
    hi = max(hi, max(mx2 & 0xffffu, mx2 >> 16));
}
// rows of the cell [s, e) whose uint16 value exceeds T: staged rows from shared memory, the rest
// (cells larger than the value stage) from global memory.  Kept out of line: it runs for ~2% of
// the cells and must not cost the common path registers.
__device__ __noinline__ uint32_t count_above_u16(const uint16_t *sv, const uint16_t *__restrict__ value_g, uint32_t s,
                                                 uint32_t e, uint32_t a0, uint32_t n_v, uint32_t T) {
    uint32_t above = 0;
    const uint32_t sh = s - a0, ngroups = (sh + n_v + 7u) >> 3;
    for (uint32_t gi = threadIdx.x; gi < ngroups; gi += kFastThreads) {
        const uint32_t r0 = a0 + gi * 8u;
        const uint4 q = *reinterpret_cast<const uint4 *>(sv + gi * 8u);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t r = r0 + i;
            const uint32_t key = (w[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
            above += (r >= s && r < s + n_v && key > T) ? 1u : 0u;
        }
    }
    // rows beyond the stage: 128-bit loads over the aligned middle, single rows at the ragged ends
    const uint32_t t0 = s + n_v;
    if (t0 < e) {
        const uint32_t m0 = min((t0 + 7u) & ~7u, e), m1 = max(m0, e & ~7u);
        const uint32_t Tc = min(T, 0xffffu), T2 = Tc | (Tc << 16);
        uint32_t bits = 0;                                                  // 16 set bits per row above T
        for (uint32_t g = (m0 >> 3) + threadIdx.x; g < (m1 >> 3); g += kFastThreads) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(value_g) + g);
            bits += __popc(__vcmpgtu2(q.x, T2)) + __popc(__vcmpgtu2(q.y, T2)) + __popc(__vcmpgtu2(q.z, T2)) + __popc(__vcmpgtu2(q.w, T2));
        }
        above += bits >> 4;
        const uint32_t nh = m0 - t0, nt = e - m1;
        if (threadIdx.x < nh + nt) {
            const uint32_t r = threadIdx.x < nh ? t0 + threadIdx.x : m1 + (threadIdx.x - nh);
            above += ((uint32_t)__ldg(value_g + r) > T) ? 1u : 0u;
        }
    }
    return above;
}

// smallest key c in [klo, khi] with pred(c) true, for a monotone pred with pred(khi) true; one full warp, 32-ary search
// (each round every lane tests one candidate: seven rounds cover 32 bits)
template <typename Pred>
__device__ __forceinline__ uint32_t warp_lower_bound(uint32_t klo, uint32_t khi, Pred pred) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t L = klo, H = khi;
    while (L < H) {
        const uint32_t w = H - L;
        const uint32_t c = L + (uint32_t)(((unsigned long long)w * (lane + 1u)) / 33ull);      // in [L, H)
        const uint32_t m = __ballot_sync(0xffffffffu, pred(c));
        if (m) {
            const int f = __ffs(m) - 1;
            H = __shfl_sync(0xffffffffu, c, f);
            const uint32_t below = __shfl_sync(0xffffffffu, c, f > 0 ? f - 1 : 0);
            if (f > 0) L = below + 1u;
        } else {
            L = __shfl_sync(0xffffffffu, c, 31) + 1u;
        }
    }
    return H;
}

struct FastArgs {
    TokArgs t;
    uint32_t gcap;                // gene stage capacity (rows, multiple of 8)
    uint32_t vcap;                // value stage capacity (rows, multiple of 8)
    const uint32_t *vstar_table;  // [65536] smallest uint16 value whose window bin is >= 1, per cell maximum
    uint32_t *ticket;             // dynamic row counter (zeroed by K1): rows are handed out in order
    // packed wire format (pack_rows.h): one 16-byte aligned block per cell; blk_off[c] in 16-byte units
    const uint8_t *packed;
    const uint32_t *blk_off;
    uint32_t rawcap;              // bytes of one raw-block stage (a cell whose block is larger goes to the general kernel)
};

struct CellDesc {
    uint32_t j;                   // output row (0xffffffff: no more work)
    uint32_t s, e, c, a0;
    uint32_t g_bulk, v_bulk;      // rows staged by the bulk copies, counted from a0 (multiples of 8); packed: v_bulk = 0 marks an oversize block
    uint32_t n_g, n_v;            // rows of the cell the stages must hold (from s)
};

// packed route: the cell's block goes into the raw stage with ONE bulk copy (or not at all when it does not fit)
__device__ __forceinline__ void packed_issue(const FastArgs &fa, uint32_t j, uint32_t c, uint32_t s, uint32_t e, uint32_t boff, uint32_t bend,
                                             uint32_t n_g_want, CellDesc *d, uint8_t *raw, uint64_t *bar) {
    const TokArgs &a = fa.t;
    d->j = j;
    if (j >= a.n_cells) { mbar_arrive(bar); return; }
    const uint32_t n = e - s, bytes = (bend - boff) << 4;
    const bool fits = bytes <= fa.rawcap && n + 8u <= fa.vcap && n_g_want + 8u <= fa.gcap + 8u;
    d->s = s; d->e = e; d->c = c; d->a0 = s;
    d->g_bulk = 0u; d->v_bulk = fits ? 1u : 0u; d->n_g = n_g_want; d->n_v = n;
    SLAFB_CHECK(e <= a.n_rows && s <= e && bend >= boff, 221, bytes);
    if (!fits || bytes == 0u) { mbar_arrive(bar); return; }
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s(raw, fa.packed + ((size_t)boff << 4), bytes, bar);
}

// value of row `row` in an exception list {row, value} x n (ascending rows): the lists are short, a linear scan is enough
__device__ __forceinline__ uint32_t packed_exception(const uint32_t *ex, uint32_t n, uint32_t row) {
    for (uint32_t i = 0; i < n; i++)
        if (ex[2u * i] == row) return ex[2u * i + 1u];
    return 255u;      // (malformed block: the byte stands for itself)
}

template <typename GeneT, typename ValT, int MODE>
__device__ __forceinline__ void fast_issue(const FastArgs &fa, uint32_t j, uint32_t c, uint32_t s, uint32_t e, CellDesc *d,
                                           uint8_t *gstage, uint8_t *vstage, uint64_t *bar) {
    constexpr bool SCGPT = (MODE == MODE_SCGPT_RAW || MODE == MODE_SCGPT_BINNED);
    const TokArgs &a = fa.t;
    d->j = j;
    if (j >= a.n_cells) { mbar_arrive(bar); return; }     // end marker for the consumers
    const uint32_t n = e - s, a0 = s & ~7u, R8 = a.n_rows & ~7u;
    const uint32_t limit = SCGPT ? a.max_genes : (a.L - 1u);
    constexpr bool PCT = (MODE == MODE_GF_PCT);
    const bool need_val = SCGPT || PCT || (n > a.max_items);
    // percentile mode: any row may survive the filter, so the genes of all rows are staged (the gene stage is as long as the value stage)
    const uint32_t n_g = PCT ? min(n, fa.gcap - 8u) : min(n, limit);
    const uint32_t n_v = need_val ? min(n, fa.vcap - 8u) : 0u;
    const uint32_t g_end = min((s + n_g + 7u) & ~7u, R8), v_end = min((s + n_v + 7u) & ~7u, R8);
    const uint32_t g_bulk = (n_g && g_end > a0) ? g_end - a0 : 0u;
    const uint32_t v_bulk = (n_v && v_end > a0) ? v_end - a0 : 0u;
    d->s = s; d->e = e; d->c = c; d->a0 = a0;
    d->g_bulk = g_bulk; d->v_bulk = v_bulk; d->n_g = n_g; d->n_v = n_v;
    SLAFB_CHECK(g_bulk <= fa.gcap && v_bulk <= fa.vcap, 201, g_bulk > fa.gcap ? g_bulk : v_bulk);
    SLAFB_CHECK(a0 + max(g_bulk, v_bulk) <= a.n_rows && e <= a.n_rows && s <= e, 202, a0 + max(g_bulk, v_bulk));
    SLAFB_CHECK(c < 0x7fffffffu && j < a.n_cells, 203, j);
    const uint32_t bytes = g_bulk * (uint32_t)sizeof(GeneT) + v_bulk * (uint32_t)sizeof(ValT);
    if (bytes == 0) {
        mbar_arrive(bar);
    } else {
        mbar_arrive_expect_tx(bar, bytes);
        if (g_bulk) bulk_g2s(gstage, static_cast<const GeneT *>(a.gene) + a0, g_bulk * (uint32_t)sizeof(GeneT), bar);
        if (v_bulk) bulk_g2s(vstage, static_cast<const ValT *>(a.value) + a0, v_bulk * (uint32_t)sizeof(ValT), bar);
    }
}

static_assert(kHashMaxDistinct + kFastThreads < kHashSlots, "the distinct-value set must never fill up");

// CTAs per SM the register allocation is sized for: the uint16 count modes stage ~21-25 KB per CTA and run 8 per SM (32
// registers per thread); float32 values and the percentile mode stage more (6 per SM fit), which leaves them 40 registers.
template <typename ValT, int MODE>
constexpr int fast_ctas_per_sm() { return (sizeof(ValT) == 4 || MODE == MODE_GF_PCT) ? (1536 / kFastThreads) : (2048 / kFastThreads); }

template <typename GeneT, typename ValT, int MODE, bool PACKED = false>
__global__ void __launch_bounds__(kFastThreads, fast_ctas_per_sm<ValT, MODE>()) tokenize_cells_kernel(const FastArgs fa) {
    static_assert(!PACKED || sizeof(ValT) == 2, "the packed wire format carries uint16 counts");
    constexpr bool SCGPT = (MODE == MODE_SCGPT_RAW || MODE == MODE_SCGPT_BINNED);
    constexpr bool BINNED = (MODE == MODE_SCGPT_BINNED);
    constexpr bool PCT = (MODE == MODE_GF_PCT);
    constexpr bool U16 = (sizeof(ValT) == 2);
    static_assert(!PCT || U16, "percentile mode runs in this kernel for uint16 values only (float32: general kernel)");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ CellDesc s_desc[2];
    __shared__ uint32_t s_ws[3 * kFastWarps];
    __shared__ uint32_t s_xstar;      // float32 binned window: key of the smallest value whose bin is >= 1
    __shared__ uint32_t s_bitmap[PCT ? kPctWords : 1];   // percentile mode: which values of [32, 8192) occur in the cell (bit = value)
    __shared__ uint32_t s_thr;        // percentile mode: the smallest value that is kept
    __shared__ uint32_t s_dirty;      // percentile mode: the bitmap holds bits of the current cell (wiped when the cell is done)
    __shared__ uint32_t s_ap[PCT ? 64 : 1];   // percentile mode: ceil(p * D / 100) for D < 64 (f64, evaluated once per CTA instead of per cell and thread)
    const TokArgs &a = fa.t;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t gbytes = fa.gcap * (uint32_t)sizeof(GeneT), vbytes = fa.vcap * (uint32_t)sizeof(ValT);
    const uint32_t stage_bytes = gbytes + vbytes;   // both multiples of 16
    const GeneT *__restrict__ gene_g = static_cast<const GeneT *>(a.gene);
    const ValT *__restrict__ value_g = static_cast<const ValT *>(a.value);
    const uint32_t R = a.n_rows, R8 = R & ~7u, k = a.max_items, L = a.L;
    const uint32_t limit = SCGPT ? a.max_genes : (L - 1u);

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        s_dirty = 0u;
    }
    if constexpr (PCT) {
        for (uint32_t w = tid; w < (uint32_t)kPctWords; w += kFastThreads) s_bitmap[w] = 0u;
        if (tid < 64) s_ap[tid] = target_asc_rank((uint32_t)tid, 0xffffffffu, true, a.min_pct);    // k = "infinite": the percentile term alone (>= 1)
    }
    __syncthreads();
    // Rows are handed out IN ORDER by an atomic ticket (first row = blockIdx.x), so the rows being
    // written at any moment form one compact window of the output tensors -- measured 6.0 TB/s
    // for this store pattern against 4.3-5.2 TB/s with a fixed stride (scripts/micro/store_probe.cu).
    // Thread 0 runs three rows ahead: ticket -> permutation entry -> segment bounds -> bulk copy,
    // one step per iteration, so none of those latencies is ever waited on in line.
    const uint32_t G = gridDim.x, NONE = 0xffffffffu;
    uint32_t j1 = NONE, c1 = 0, s1 = 0, e1 = 0;   // next row: segment known, copy issued at the top of the loop
    uint32_t b1 = 0, be1 = 0;                     // (packed route: its block bounds)
    uint32_t j2 = NONE, c2 = 0;                   // row after: permutation entry loaded
    uint32_t j3 = NONE;                           // row after that: ticket drawn
    // packed route: [raw block stage 0][raw block stage 1][gene stage][value stage] -- the bulk copies double-buffer the raw blocks,
    // the decoded genes / values of the CURRENT cell live in one gene and one value stage
    auto want_genes = [&](uint32_t n) -> uint32_t { return PCT ? n : min(n, limit); };
    (void)b1; (void)be1;
    if (tid == 0) {
        const uint32_t j0 = blockIdx.x;
        if (j0 < a.n_cells) {
            const uint32_t c0 = a.perm ? (uint32_t)__ldg(a.perm + j0) : j0;
            const uint32_t s0 = __ldg(a.seg_start + c0), e0 = __ldg(a.seg_start + c0 + 1);
            if constexpr (PACKED) packed_issue(fa, j0, c0, s0, e0, __ldg(fa.blk_off + c0), __ldg(fa.blk_off + c0 + 1), want_genes(e0 - s0), &s_desc[0], smem_raw, &s_bar[0]);
            else fast_issue<GeneT, ValT, MODE>(fa, j0, c0, s0, e0, &s_desc[0], smem_raw, smem_raw + gbytes, &s_bar[0]);
        } else {
            if constexpr (PACKED) packed_issue(fa, NONE, 0, 0, 0, 0, 0, 0, &s_desc[0], smem_raw, &s_bar[0]);
            else fast_issue<GeneT, ValT, MODE>(fa, NONE, 0, 0, 0, &s_desc[0], smem_raw, smem_raw + gbytes, &s_bar[0]);
        }
        j1 = G + atomicAdd(fa.ticket, 1u);
        j2 = G + atomicAdd(fa.ticket, 1u);
        j3 = G + atomicAdd(fa.ticket, 1u);
        if (j1 < a.n_cells) {
            c1 = a.perm ? (uint32_t)__ldg(a.perm + j1) : j1;
            s1 = __ldg(a.seg_start + c1);
            e1 = __ldg(a.seg_start + c1 + 1);
            if constexpr (PACKED) { b1 = __ldg(fa.blk_off + c1); be1 = __ldg(fa.blk_off + c1 + 1); }
        }
        if (j2 < a.n_cells) c2 = a.perm ? (uint32_t)__ldg(a.perm + j2) : j2;
    }
    uint32_t phase_bits = 0u;   // bit st = parity the next wait on stage st must see

    for (uint32_t it = 0;; it++) {
        const uint32_t st = it & 1u;
        // stage st^1 was released by the barrier that ended the previous iteration
        if (tid == 0) {
            if constexpr (PACKED)
                packed_issue(fa, j1 < a.n_cells ? j1 : NONE, c1, s1, e1, b1, be1, want_genes(e1 - s1), &s_desc[st ^ 1u],
                             smem_raw + (st ^ 1u) * fa.rawcap, &s_bar[st ^ 1u]);
            else
                fast_issue<GeneT, ValT, MODE>(fa, j1 < a.n_cells ? j1 : NONE, c1, s1, e1, &s_desc[st ^ 1u],
                                              smem_raw + (st ^ 1u) * stage_bytes, smem_raw + (st ^ 1u) * stage_bytes + gbytes,
                                              &s_bar[st ^ 1u]);
            j1 = j2; c1 = c2;
            if (j1 < a.n_cells) {
                s1 = __ldg(a.seg_start + c1); e1 = __ldg(a.seg_start + c1 + 1);
                if constexpr (PACKED) { b1 = __ldg(fa.blk_off + c1); be1 = __ldg(fa.blk_off + c1 + 1); }
            }
            j2 = j3;
            if (j2 < a.n_cells) c2 = a.perm ? (uint32_t)__ldg(a.perm + j2) : j2;
            j3 = (j3 < a.n_cells) ? G + atomicAdd(fa.ticket, 1u) : NONE;
        }
        mbar_wait(&s_bar[st], (phase_bits >> st) & 1u);
        phase_bits ^= 1u << st;
        const CellDesc d = s_desc[st];
        const uint32_t j = d.j;
        if (j == NONE) break;                       // uniform: every thread reads the same descriptor
        GeneT *sg = reinterpret_cast<GeneT *>(smem_raw + (PACKED ? 2u * fa.rawcap : st * stage_bytes));
        ValT *sv = reinterpret_cast<ValT *>(smem_raw + (PACKED ? 2u * fa.rawcap : st * stage_bytes) + gbytes);
        const uint32_t s = d.s, e = d.e, n = e - s, a0 = d.a0, sh = s - a0;

        bool oversize = false;
        if constexpr (PACKED) {
            // ---- decode the cell's block (pack_rows.h) from the raw stage into the gene / value stages: 8 rows per thread,
            // values widened to uint16 (byte 255 = look the row up in the exception list), gene ids = base + running sum of
            // (delta + 1) -- a local prefix per thread and one block scan per 2,048 rows.
            oversize = (d.v_bulk == 0u) && n > 0u;
            if (!oversize && n > 0u) {
                const uint8_t *rb = smem_raw + st * fa.rawcap;
                const uint4 hdr = *reinterpret_cast<const uint4 *>(rb);
                const uint32_t np = (n + 15u) & ~15u, base_gene = hdr.y, n_eg = hdr.z, n_ev = hdr.w;
                const uint8_t *dl = rb + 16u, *vl = dl + np;
                const uint32_t *exg = reinterpret_cast<const uint32_t *>(vl + np), *exv = exg + 2u * n_eg;
                SLAFB_CHECK(hdr.x == n && 16u + 2u * np + 8u * (n_eg + n_ev) <= fa.rawcap, 222, hdr.x);
                uint32_t carry = 0u;
                for (uint32_t base = 0u; base < n; base += (uint32_t)kFastThreads * 8u) {
                    const uint32_t i0 = base + (uint32_t)tid * 8u;
                    uint32_t dd[8], vv[8], run = 0u;
                    if (i0 < n) {
                        const uint2 dq = *reinterpret_cast<const uint2 *>(dl + i0), vq = *reinterpret_cast<const uint2 *>(vl + i0);
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            dd[i] = ((i < 4 ? dq.x : dq.y) >> ((i & 3) * 8)) & 0xffu;
                            vv[i] = ((i < 4 ? vq.x : vq.y) >> ((i & 3) * 8)) & 0xffu;
                        }
                        if (n_ev) {
#pragma unroll
                            for (int i = 0; i < 8; i++) if (vv[i] == 255u && i0 + i < n) vv[i] = packed_exception(exv, n_ev, i0 + i);
                        }
                        if (n_eg) {
#pragma unroll
                            for (int i = 0; i < 8; i++) if (dd[i] == 255u && i0 + i < n) dd[i] = packed_exception(exg, n_eg, i0 + i);
                        }
#pragma unroll
                        for (int i = 0; i < 8; i++) { run += (i0 + i < n) ? dd[i] + 1u : 0u; dd[i] = run; }    // inclusive prefix inside the thread
                        SLAFB_CHECK(i0 + 8u <= fa.vcap, 223, i0);
                        *reinterpret_cast<uint4 *>(sv + i0) = make_uint4(vv[0] | (vv[1] << 16), vv[2] | (vv[3] << 16), vv[4] | (vv[5] << 16), vv[6] | (vv[7] << 16));
                    }
                    const uint32_t inc = warp_incl_scan(run);
                    if (lane == 31) s_ws[warp] = inc;
                    __syncthreads();
                    uint32_t wbase = 0u, tot = 0u;
#pragma unroll
                    for (int w2 = 0; w2 < kFastWarps; w2++) {
                        const uint32_t t = s_ws[w2];
                        if (w2 < warp) wbase += t;
                        tot += t;
                    }
                    if (i0 < d.n_g) {
                        const uint32_t g0 = base_gene + carry + wbase + inc - run - 1u;     // gene id of row i0 is g0 + dd[0]
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            if (i0 + i < d.n_g) {
                                SLAFB_CHECK(i0 + i < fa.gcap, 224, i0 + i);
                                sg[i0 + i] = (GeneT)(g0 + dd[i]);
                            }
                        }
                    }
                    carry += tot;
                    __syncthreads();        // the scratch is free again; after the last chunk: both stages are complete
                }
            }
        }

        if constexpr (PACKED) {
            if (oversize) {                         // a block that does not fit the raw stage: unpacked for, and tokenized by, the general kernel
                if (tid == 0) {
                    const uint32_t wi = atomicAdd(a.work_counters, 1u);
                    SLAFB_CHECK(wi < a.n_cells, 208, wi);
                    a.work_list[wi] = j;
                }
                __syncthreads();
                continue;
            }
        }

        // rows in the last partial 8-group of the column arrays cannot be bulk-copied (16-byte granules)
        const bool patch = !PACKED && (s + max(d.n_g, d.n_v) > R8);
        if (patch) {
            if (tid < 8) {
                const uint32_t r = R8 + tid;
                if (r < R && r >= a0) {
                    if (r < s + d.n_g) sg[r - a0] = gene_g[r];
                    if (r < s + d.n_v) sv[r - a0] = value_g[r];
                }
            }
            __syncthreads();
        }

        // ---- statistics of the cell's values (range test for the dense-rank filter, max for the bins)
        bool keep_all = PCT ? false : (n <= k);    // the percentile rule also filters cells with few rows
        bool defer = false;
        uint32_t kmax = 0, thr = 0, dirty = 0;
        float m_f32 = 0.0f;
        (void)thr; (void)dirty;
        if constexpr (!U16) {
            // float32 values: ONE pass over the cell's values gives the maximum (for the bins) and, when the cell has more
            // rows than max_items, the exact number of distinct values through an open-addressing set in shared memory.
            // Normalised expression values take as many distinct values per cell as the counts they came from (tens), so
            // "at most k distinct" holds almost always and the cell stays on this path.
            if (BINNED || !keep_all) {
                const bool do_hash = !keep_all;
                uint32_t *tab = reinterpret_cast<uint32_t *>(smem_raw + 2u * stage_bytes);
                const uint32_t cap = min(k, (uint32_t)kHashMaxDistinct);
                if (do_hash) {
                    for (uint32_t i = tid; i < (uint32_t)kHashSlots; i += kFastThreads) tab[i] = 0xffffffffu;
                    if (tid == 0) s_ws[0] = 0u;
                    __syncthreads();
                }
                volatile uint32_t *distinct = s_ws;
                volatile uint32_t *vt = tab;
                float xmax = -INFINITY;
                uint32_t last = 0xffffffffu;                        // the key this thread probed last (the table's "empty" pattern, never a key)
                bool hashing = do_hash;
                // set membership is about equality only: the raw bits of x + 0.0f (which folds -0.0 into +0.0) serve as the key
                auto probe = [&](float x) {
                    const uint32_t key = __float_as_uint(x + 0.0f);
                    if (key == last) return;                        // counts repeat: the same value again costs nothing
                    last = key;
                    uint32_t h = (key * 2654435761u) >> (32 - kHashBits);
                    for (;;) {
                        SLAFB_CHECK(h < (uint32_t)kHashSlots, 204, h);
                        uint32_t cur = vt[h];                       // plain read first: the common case is "already present"
                        if (cur == 0xffffffffu) {
                            cur = atomicCAS(tab + h, 0xffffffffu, key);
                            if (cur == 0xffffffffu) { atomicAdd(s_ws, 1u); break; }
                        }
                        if (cur == key) break;
                        h = (h + 1u) & (uint32_t)(kHashSlots - 1);
                        if (*distinct > cap) break;                 // the set will not decide any more: stop probing
                    }
                };
                // staged rows: aligned groups of four (one 128-bit shared-memory load), ragged ends element by element
                const uint32_t nst = d.n_v, nq = (sh + nst + 3u) >> 2;
                const float4 *sv4 = reinterpret_cast<const float4 *>(sv);
                for (uint32_t c = tid; c < nq; c += kFastThreads) {
                    const float4 q = sv4[c];
                    const uint32_t p0 = c << 2;
                    if (hashing && *distinct > cap) hashing = false;        // decided (deferred): only the maximum is still needed
                    if (p0 >= sh && p0 + 4u <= sh + nst) {
                        xmax = fmaxf(fmaxf(xmax, fmaxf(q.x, q.y)), fmaxf(q.z, q.w));
                        if (hashing) { probe(q.x); probe(q.y); probe(q.z); probe(q.w); }
                    } else {
                        const float x4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint32_t pp = p0 + (uint32_t)i;
                            if (pp >= sh && pp < sh + nst) {
                                xmax = fmaxf(xmax, x4[i]);
                                if (hashing) probe(x4[i]);
                            }
                        }
                    }
                }
                for (uint32_t r = s + nst + tid; r < e; r += kFastThreads) {    // rows beyond the stage: from global memory
                    const float x = __ldg(value_g + r);
                    xmax = fmaxf(xmax, x);
                    if (hashing && *distinct <= cap) probe(x);
                }
                uint32_t hi = __reduce_max_sync(0xffffffffu, key_of(xmax));
                if (lane == 0) s_ws[kFastWarps + warp] = hi;
                __syncthreads();
#pragma unroll
                for (int w = 0; w < kFastWarps; w++) hi = max(hi, s_ws[kFastWarps + w]);
                kmax = hi;
                // log1p is monotone: max(log1pf(x)) = log1pf(max x).  Only warp 0 (the threshold search below) needs it.
                if constexpr (BINNED) { if (warp == (int)(it % (uint32_t)kFastWarps)) m_f32 = log1pf_window(float_of_key(hi)); }
                if (do_hash) {
                    const uint32_t D = s_ws[0];
                    if (D <= cap) keep_all = true; else defer = true;            // D <= cap <= k distinct values: every row has rank <= k
                }
                __syncthreads();                                                  // s_ws is reused by the next cell
            }
        }
        if (U16 && !PCT && (BINNED || !keep_all)) {
            uint32_t lo = 0xffffffffu, hi = 0u;
            const uint32_t ngroups = (sh + d.n_v + 7u) >> 3;
            for (uint32_t gi = tid; gi < ngroups; gi += kFastThreads) {
                const uint32_t r0 = a0 + gi * 8u;
                {
                    const uint4 q = *reinterpret_cast<const uint4 *>(sv + gi * 8u);
                    if (r0 >= s && r0 + 8u <= s + d.n_v) {
                        uint32_t mn = __vminu2(__vminu2(q.x, q.y), __vminu2(q.z, q.w));
                        uint32_t mx = __vmaxu2(__vmaxu2(q.x, q.y), __vmaxu2(q.z, q.w));
                        lo = min(lo, min(mn & 0xffffu, mn >> 16));
                        hi = max(hi, max(mx & 0xffffu, mx >> 16));
                    } else {
                        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const uint32_t r = r0 + i;
                            if (r >= s && r < s + d.n_v) {
                                const uint32_t key = (w[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                                lo = min(lo, key);
                                hi = max(hi, key);
                            }
                        }
                    }
                }
            }
            // rows beyond the value stage (cells larger than vcap): straight from global memory
            if constexpr (U16) minmax_rows_u16(reinterpret_cast<const uint16_t *>(value_g), s + d.n_v, e, lo, hi);
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if (lane == 0) { s_ws[warp] = lo; s_ws[kFastWarps + warp] = hi; }
            __syncthreads();
#pragma unroll
            for (int w = 0; w < kFastWarps; w++) {
                lo = min(lo, s_ws[w]);
                hi = max(hi, s_ws[kFastWarps + w]);
            }
            kmax = hi;
            if (!keep_all) {
                if (U16 && (hi - lo + 1u <= k)) {
                    keep_all = true;                                // at most k distinct values fit in [lo, hi]
                } else if (U16 && k >= 8u) {
                    // Second sufficient test (count data has a handful of very large values): with
                    // T = lo + k - 1 - M, the values <= T take at most k - M distinct values and the
                    // rows above T at most one each, so  #rows(value > T) <= M  implies  #distinct <= k.
                    const uint32_t M = min(64u, k >> 1);
                    const uint32_t T = lo + (k - 1u - M);
                    uint32_t above = count_above_u16(reinterpret_cast<const uint16_t *>(sv), reinterpret_cast<const uint16_t *>(value_g),
                                                     s, e, a0, d.n_v, T);
                    above = warp_sum(above);
                    __syncthreads();                                // scratch rows are free again
                    if (lane == 0) s_ws[warp] = above;
                    __syncthreads();
                    above = 0;
#pragma unroll
                    for (int w2 = 0; w2 < kFastWarps; w2++) above += s_ws[w2];
                    if (above <= M) keep_all = true; else defer = true;
                } else {
                    defer = true;                                   // exact k-th distinct value: general kernel
                }
            }
        }

        uint32_t n_keep = n;
        bool compacted = false;          // percentile mode: the kept genes were gathered into the value stage's memory
        if constexpr (PCT) {
            // Percentile branch (aggregators.py:167-196): the filter needs the exact number of distinct values D and the a*-th
            // smallest of them.  ONE pass over the staged values gives the range and a presence mask: the values below 32 --
            // nearly every row of count data -- are collected in a register and reduced over the warps, the rare larger ones
            // set their bit in a shared-memory bitmap indexed by the value itself (a handful of atomics per cell); a value
            // beyond the bitmap sends the cell to the general kernel.  Thread t owns the contiguous 4-row chunks
            // [t * Q, (t + 1) * Q), so the compaction below needs a single block scan per cell.
            const uint32_t nst = d.n_v;                                 // staged rows (== n unless the cell is longer than the stage)
            const uint32_t nch = (sh + nst + 3u) >> 2;
            const uint32_t Q = (nch + (uint32_t)kFastThreads - 1u) / (uint32_t)kFastThreads;
            const uint2 *sv2 = reinterpret_cast<const uint2 *>(sv);
            volatile uint32_t *vb = s_bitmap;
            uint32_t mn2 = 0xffffffffu, mx2 = 0u, m0 = 0u, big = 0u;
            auto one = [&](uint32_t v) {                                // a valid value
                m0 |= shl_clamp(1u, v);
                if (v >= 32u) {
                    big = 1u;
                    if (v < (uint32_t)kPctBits) {
                        const uint32_t bit = 1u << (v & 31u);
                        SLAFB_CHECK((v >> 5) < (uint32_t)kPctWords, 205, v);
                        if (!(vb[v >> 5] & bit)) atomicOr(&s_bitmap[v >> 5], bit);      // look before the atomic
                    }
                }
            };
            for (uint32_t q = 0; q < Q; q++) {
                const uint32_t c = (uint32_t)tid * Q + q;
                if (c >= nch) break;
                const uint2 w = sv2[c];
                const uint32_t p0 = c << 2;
                if (p0 >= sh && p0 + 4u <= sh + nst) {                  // whole chunk inside the cell
                    mn2 = __vminu2(__vminu2(mn2, w.x), w.y);
                    mx2 = __vmaxu2(__vmaxu2(mx2, w.x), w.y);
                    m0 |= shl_clamp(1u, w.x & 0xffffu) | shl_clamp(1u, w.x >> 16) | shl_clamp(1u, w.y & 0xffffu) | shl_clamp(1u, w.y >> 16);
                    if ((w.x | w.y) & 0xffe0ffe0u) {                    // rare: a value >= 32
                        const uint32_t v4[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
#pragma unroll
                        for (int i = 0; i < 4; i++) if (v4[i] >= 32u) one(v4[i]);
                    }
                } else {
                    const uint32_t v4[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t pp = p0 + (uint32_t)i;
                        if (pp >= sh && pp < sh + nst) {
                            mn2 = __vminu2(mn2, v4[i] | (v4[i] << 16));
                            mx2 = __vmaxu2(mx2, v4[i] | (v4[i] << 16));
                            one(v4[i]);
                        }
                    }
                }
            }
            for (uint32_t r = s + nst + tid; r < e; r += kFastThreads) {        // rows beyond the stage: from global memory
                const uint32_t v = (uint32_t)__ldg(value_g + r);
                mn2 = __vminu2(mn2, v | (v << 16));
                mx2 = __vmaxu2(mx2, v | (v << 16));
                one(v);
            }
            uint32_t lo = __reduce_min_sync(0xffffffffu, min(mn2 & 0xffffu, mn2 >> 16));
            uint32_t hi = __reduce_max_sync(0xffffffffu, max(mx2 & 0xffffu, mx2 >> 16));
            m0 = __reduce_or_sync(0xffffffffu, m0);
            big = __reduce_or_sync(0xffffffffu, big);
            if (lane == 0) { s_ws[warp] = lo; s_ws[kFastWarps + warp] = hi; s_ws[2 * kFastWarps + warp] = m0; if (big) s_dirty = 1u; }
            __syncthreads();
            m0 = 0u;
#pragma unroll
            for (int w = 0; w < kFastWarps; w++) {
                lo = min(lo, s_ws[w]);
                hi = max(hi, s_ws[kFastWarps + w]);
                m0 |= s_ws[2 * kFastWarps + w];
            }
            kmax = hi;
            dirty = s_dirty;                                            // uniform; the bitmap is wiped at the end of the iteration
            (void)lo;
            if (hi >= (uint32_t)kPctBits) {
                defer = true;
            } else {
                const uint32_t d0 = __popc(m0);                         // distinct values below 32
                constexpr int WPT = (kPctWords + kFastThreads - 1) / kFastThreads;     // bitmap words per thread (contiguous)
                uint32_t D = d0, pc = 0, wv[WPT], inc = 0, wbase = 0;
#pragma unroll
                for (int q2 = 0; q2 < WPT; q2++) wv[q2] = 0u;
                if (dirty) {                                            // some value >= 32: count the bitmap's words 1 .. hi / 32
#pragma unroll
                    for (int q2 = 0; q2 < WPT; q2++) {
                        const uint32_t wi = (uint32_t)tid * WPT + q2;
                        wv[q2] = (wi >= 1u && wi <= (hi >> 5) && wi < (uint32_t)kPctWords) ? s_bitmap[wi] : 0u;
                        pc += __popc(wv[q2]);
                    }
                    inc = warp_incl_scan(pc);
                    __syncthreads();                                    // (everyone has read lo / hi / m0 from the scratch)
                    if (lane == 31) s_ws[warp] = inc;
                    __syncthreads();
#pragma unroll
                    for (int w2 = 0; w2 < kFastWarps; w2++) {
                        const uint32_t t = s_ws[w2];
                        if (w2 < warp) wbase += t;
                        D += t;
                    }
                }
                // a* = max(D - min(k, D) + 1, ceil(p * D / 100)); the second term from the per-CTA table when D is small
                const uint32_t astar = (D < 64u) ? max(D - min(k, D) + 1u, s_ap[D]) : target_asc_rank(D, k, true, a.min_pct);
                if (astar <= d0) {                                      // the usual case: the threshold is one of the small values
                    uint32_t x = m0;
                    for (uint32_t i = 1u; i < astar; i++) x &= x - 1u;  // drop the lower set bits (astar is small: a percentile of D)
                    thr = (uint32_t)(__ffs(x) - 1);
                } else {                                                // it lies in the bitmap: the (astar - d0)-th set bit of words >= 1
                    const uint32_t want = astar - d0;
                    uint32_t ex = wbase + inc - pc;
                    if (want > ex && want <= ex + pc) {
#pragma unroll
                        for (int q2 = 0; q2 < WPT; q2++) {
                            const uint32_t c2 = __popc(wv[q2]);
                            if (want > ex && want <= ex + c2) {
                                uint32_t x = wv[q2];
                                for (uint32_t i = ex + 1u; i < want; i++) x &= x - 1u;
                                s_thr = (((uint32_t)tid * WPT + q2) << 5) + (uint32_t)(__ffs(x) - 1);
                            }
                            ex += c2;
                        }
                    }
                    __syncthreads();
                    thr = s_thr;
                }
                if (astar == 1u) keep_all = true;                       // every distinct value passes both rules
            }
            if (!defer && !keep_all && nst == n && Q <= 8u && limit * (uint32_t)sizeof(GeneT) <= vbytes) {
                // Row-order compaction, whole cell staged: flags from the values (phase A), ONE block scan, then every thread gathers
                // the kept genes of its own rows into the value stage's memory (phase B; the values are not needed any more and
                // everybody finished reading them before the scan's barrier).  The emission then reads a dense gene list from there.
                const uint32_t thr2 = thr | (thr << 16);
                uint32_t fl = 0u;                                       // 4 bits per chunk, Q <= 8 chunks
                for (uint32_t q = 0; q < Q; q++) {
                    const uint32_t c = (uint32_t)tid * Q + q;
                    if (c >= nch) break;
                    const uint2 w = sv2[c];
                    const uint32_t p0 = c << 2;
                    const uint32_t cx = __vcmpgeu2(w.x, thr2), cy = __vcmpgeu2(w.y, thr2);     // 0xffff per half that passes
                    uint32_t f4 = (cx & 1u) | ((cx >> 15) & 2u) | ((cy & 1u) << 2) | ((cy >> 13) & 8u);
                    if (!(p0 >= sh && p0 + 4u <= sh + n)) {             // ragged ends: rows outside the cell do not count
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const uint32_t pp = p0 + (uint32_t)i;
                            if (!(pp >= sh && pp < sh + n)) f4 &= ~(1u << i);
                        }
                    }
                    fl |= f4 << (4u * q);
                }
                const uint32_t pc = __popc(fl), inc = warp_incl_scan(pc);
                __syncthreads();                                        // the scratch is free, and every thread has read its values
                if (lane == 31) s_ws[warp] = inc;
                __syncthreads();
                uint32_t wbase = 0, tot = 0;
#pragma unroll
                for (int w2 = 0; w2 < kFastWarps; w2++) {
                    const uint32_t t = s_ws[w2];
                    if (w2 < warp) wbase += t;
                    tot += t;
                }
                uint32_t off = wbase + inc - pc;
                const bool room = off + pc <= limit;                    // all of this thread's kept genes are emitted: no per-row test
                GeneT *dst = reinterpret_cast<GeneT *>(sv);
                for (uint32_t q = 0; q < Q && fl; q++) {
                    const uint32_t c = (uint32_t)tid * Q + q, f4 = (fl >> (4u * q)) & 15u;
                    if (!f4) continue;
                    GeneT g4[4];
                    if constexpr (sizeof(GeneT) == 2) {
                        const uint2 gq = reinterpret_cast<const uint2 *>(sg)[c];
                        g4[0] = (GeneT)(gq.x & 0xffffu); g4[1] = (GeneT)(gq.x >> 16); g4[2] = (GeneT)(gq.y & 0xffffu); g4[3] = (GeneT)(gq.y >> 16);
                    } else {
                        const uint4 gq = reinterpret_cast<const uint4 *>(sg)[c];
                        g4[0] = (GeneT)gq.x; g4[1] = (GeneT)gq.y; g4[2] = (GeneT)gq.z; g4[3] = (GeneT)gq.w;
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (f4 & (1u << i)) {
                            SLAFB_CHECK(off >= limit || (off + 1u) * sizeof(GeneT) <= vbytes, 207, off);
                            if (room || off < limit) dst[off] = g4[i];
                            off++;
                        }
                    }
                }
                n_keep = tot;
                compacted = true;
                __syncthreads();                                        // the gathered list is complete
            } else if (!defer && !keep_all) {
                // a cell longer than the stages: chunks of 8 rows per thread, rows beyond the stages from global memory, kept
                // genes written back IN PLACE in the gene stage at or before the positions they were read from
                const uint32_t staged = min(d.n_g, d.n_v);            // rows [s, s + staged) lie in both stages
                const uint32_t ng_all = (sh + n + 7u) >> 3;
                uint32_t base = 0;
                for (uint32_t g0 = 0; g0 < ng_all && base < limit; g0 += kFastThreads) {
                    const uint32_t gi = g0 + tid, r0 = a0 + gi * 8u;
                    uint32_t flags = 0;
                    GeneT g[8];
                    if (gi < ng_all) {
                        if (r0 >= s && r0 + 8u <= s + staged) {       // whole group staged: vector loads
                            const uint4 qv = *reinterpret_cast<const uint4 *>(sv + gi * 8u);
                            const uint32_t w4[4] = {qv.x, qv.y, qv.z, qv.w};
                            if constexpr (sizeof(GeneT) == 2) {
                                const uint4 qg = *reinterpret_cast<const uint4 *>(sg + gi * 8u);
                                const uint32_t g4[4] = {qg.x, qg.y, qg.z, qg.w};
#pragma unroll
                                for (int i = 0; i < 8; i++) g[i] = (GeneT)((g4[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
                            } else {
                                const uint4 qa = *reinterpret_cast<const uint4 *>(sg + gi * 8u), qb = *(reinterpret_cast<const uint4 *>(sg + gi * 8u) + 1);
                                g[0] = (GeneT)qa.x; g[1] = (GeneT)qa.y; g[2] = (GeneT)qa.z; g[3] = (GeneT)qa.w;
                                g[4] = (GeneT)qb.x; g[5] = (GeneT)qb.y; g[6] = (GeneT)qb.z; g[7] = (GeneT)qb.w;
                            }
#pragma unroll
                            for (int i = 0; i < 8; i++)
                                flags |= ((((w4[i >> 1] >> ((i & 1) * 16)) & 0xffffu) >= thr) ? 1u : 0u) << i;
                        } else {                                       // ragged ends, rows beyond the stages
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const uint32_t r = r0 + i;
                                g[i] = GeneT{};
                                if (r >= s && r < e) {
                                    const bool in = r < s + staged;
                                    const uint32_t v = in ? (uint32_t)sv[r - a0] : (uint32_t)__ldg(value_g + r);
                                    g[i] = in ? sg[r - a0] : __ldg(gene_g + r);
                                    flags |= (v >= thr ? 1u : 0u) << i;
                                }
                            }
                        }
                    }
                    const uint32_t pc = __popc(flags), inc = warp_incl_scan(pc);
                    if (lane == 31) s_ws[warp] = inc;
                    __syncthreads();                                   // all reads of this chunk are done as well
                    uint32_t wbase = 0, tot = 0;
#pragma unroll
                    for (int w2 = 0; w2 < kFastWarps; w2++) {
                        const uint32_t t = s_ws[w2];
                        if (w2 < warp) wbase += t;
                        tot += t;
                    }
                    uint32_t off = base + wbase + inc - pc;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        if (flags & (1u << i)) {
                            SLAFB_CHECK(off >= limit || sh + off < fa.gcap, 207, sh + off);
                            if (off < limit) sg[sh + off] = g[i];
                            off++;
                        }
                    }
                    base += tot;
                    __syncthreads();                                   // scratch and stage are consistent for the next chunk / the emission
                }
                n_keep = base;
            }
        }

        if (defer) {
            if (tid == 0) {                                                   // exact k-th distinct value: general kernel
                const uint32_t wi = atomicAdd(a.work_counters, 1u);
                SLAFB_CHECK(wi < a.n_cells, 208, wi);
                a.work_list[wi] = j;
            }
        } else {
            // ---- emission: token pairs (128-bit stores), value stream, PAD fill, mask, cell id
            const uint32_t n_emit = min(n_keep, limit);
            const GeneT *sge = compacted ? reinterpret_cast<const GeneT *>(sv) : sg;      // where the genes to emit start
            const uint32_t she = compacted ? 0u : sh;
            const size_t A = (size_t)j * L;
            double m_f64 = 0.0;
            uint32_t vstar = 0xffffffffu;
            const long long top = a.vocab + (long long)(a.n_bins - 1);
            if constexpr (BINNED && U16) vstar = __ldg(fa.vstar_table + kmax);
            if constexpr (BINNED && !U16) {
                // The binned window followed by the tokenizer's float branch gives  token = (bin >= 1) ? top : PAD, and
                // bin(x) = floor(log1pf(x) * n / m) is monotone in x: instead of one log1p per row, ONE warp finds the
                // smallest float whose bin is >= 1 (one round of 32 candidates around a closed-form guess, a 32-ary search
                // only if that misses) while the other warps already emit the gene stream; rows then need one compare each.
                // (the searching warp rotates with the cell: the double-precision log1p evaluations of the CTAs sharing an SM then
                // spread over its four sub-partitions instead of all landing on warp 0's)
                if (warp == (int)(it % (uint32_t)kFastWarps)) {
                    const float wb = (float)a.w_bins;
                    auto pred = [&](uint32_t c) -> bool {
                        const float lv = log1pf_window(float_of_key(c));
                        return lv > 0.0f && __fdiv_rn(__fmul_rn(lv, wb), m_f32) >= 1.0f;
                    };
                    const uint32_t kzero = key_of(0.0f);
                    uint32_t xs = 0xffffffffu;                       // no value reaches bin 1
                    if (a.w_bins >= 2 && kmax > kzero && pred(kmax)) {
                        // bin(x) >= 1  <=>  log1pf(x) * n / m >= 1, so the threshold sits within a few ulps of expm1(m / n):
                        // the 32 lanes test the 32 floats around that guess in ONE round; the exact predicate decides, the
                        // full 32-ary search only runs when the boundary is not inside the window.
                        const float x0 = (float)expm1((double)m_f32 / (double)wb);
                        uint32_t k0 = key_of(x0);
                        k0 = min(max(k0, kzero + 1u), kmax);
                        const uint32_t wlo = (k0 - kzero > 16u) ? k0 - 16u : kzero + 1u;
                        const uint32_t c = min(wlo + (uint32_t)lane, kmax);
                        const uint32_t hit = __ballot_sync(0xffffffffu, pred(c));
                        if (hit == 0u) xs = warp_lower_bound(wlo + 32u, kmax, pred);             // boundary above the window
                        else if ((hit & 1u) && wlo > kzero + 1u) xs = warp_lower_bound(kzero + 1u, wlo, pred);   // at or below its first float
                        else xs = min(wlo + (uint32_t)(__ffs(hit) - 1), kmax);
                    }
                    if (lane == 0) s_xstar = xs;
                }
            }
            float xstar_f = 0.0f;
            (void)m_f64; (void)top; (void)vstar; (void)xstar_f;
            auto id_tok = [&](long long t) -> long long {
                if (t == 0) return 1ll;                                        // CLS
                if (t <= (long long)n_emit) {
                    if constexpr (sizeof(GeneT) == 2) return (long long)sge[she + (uint32_t)t - 1u] + 4ll;
                    else return (long long)(int32_t)sge[she + (uint32_t)t - 1u] + 4ll;
                }
                return (t == (long long)n_emit + 1) ? 2ll : 0ll;               // SEP, then PAD
            };
            auto val_tok = [&](long long t) -> long long {
                if (t < 1 || t > (long long)n_emit) return 0ll;
                const ValT x = sv[sh + (uint32_t)t - 1u];
                if constexpr (BINNED && U16) return ((uint32_t)x >= vstar) ? top : 0ll;
                else if constexpr (BINNED) return ((float)x >= xstar_f) ? top : 0ll;     // xstar_f: set after the barrier below
                else return value_token<ValT, MODE>(raw_bits(x), a, m_f64, m_f32);
            };
            // Token t of the row lives at flat index A + t.  Pairs are 16-byte aligned in the flat
            // array: pair p holds tokens t0 = 2p - off and t0 + 1, off = A & 1 (odd L only).
            const uint32_t off = (uint32_t)(A & 1);
            long long *ids_al = a.ids + (A - off);           // 16-byte aligned
            long long *vals_al = SCGPT ? a.vals + (A - off) : nullptr;
            const uint32_t n_pairs = (L + off + 1u) >> 1;
            // interior pairs: both tokens are gene tokens (1 <= t0, t0 + 1 <= n_emit): no branches
            const uint32_t p_lo = (off + 2u) >> 1;                       // smallest p with t0 >= 1
            const uint32_t p_hi = n_emit >= 1u ? (n_emit - 1u + off) / 2u + 1u : 0u;   // one past the largest p with t0 + 1 <= n_emit
#pragma unroll 2
            SLAFB_CHECK(compacted || (sh + n_emit <= fa.gcap && (!SCGPT || sh + n_emit <= fa.vcap)), 209, sh + n_emit);
            SLAFB_CHECK(j < a.n_cells && 2u * n_pairs <= L + 2u, 210, n_pairs);
            for (uint32_t p = p_lo + tid; p < p_hi; p += kFastThreads) {
                const uint32_t gi = she + 2u * p - off - 1u;             // stage index of token t0
                SLAFB_CHECK((compacted || gi + 1u < fa.gcap) && 2u * p + 1u - off < L, 211, gi);
                long long x0, x1;
                if constexpr (sizeof(GeneT) == 2) { x0 = (long long)sge[gi] + 4ll; x1 = (long long)sge[gi + 1] + 4ll; }
                else { x0 = (long long)(int32_t)sge[gi] + 4ll; x1 = (long long)(int32_t)sge[gi + 1] + 4ll; }
                store_pair(ids_al + 2u * p, x0, x1);
            }
            if constexpr (BINNED && !U16) {
                __syncthreads();                                 // the searching warp has published the bin threshold
                vstar = s_xstar;
                // keys order like the floats they encode, so the rows compare as floats (one instruction instead of the key
                // transform); "no value reaches bin 1" (key 0xffffffff) decodes to a NaN, which compares false as it must
                xstar_f = float_of_key(vstar);
            }
            if constexpr (SCGPT) {
                // the value stream second: the bin-threshold load above has landed by now
#pragma unroll 2
                for (uint32_t p = p_lo + tid; p < p_hi; p += kFastThreads) {
                    const uint32_t gi = sh + 2u * p - off - 1u;
                    long long y0, y1;
                    if constexpr (BINNED && U16) {
                        y0 = ((uint32_t)sv[gi] >= vstar) ? top : 0ll;
                        y1 = ((uint32_t)sv[gi + 1] >= vstar) ? top : 0ll;
                    } else if constexpr (BINNED) {
                        y0 = ((float)sv[gi] >= xstar_f) ? top : 0ll;
                        y1 = ((float)sv[gi + 1] >= xstar_f) ? top : 0ll;
                    } else {
                        y0 = value_token<ValT, MODE>(raw_bits(sv[gi]), a, m_f64, m_f32);
                        y1 = value_token<ValT, MODE>(raw_bits(sv[gi + 1]), a, m_f64, m_f32);
                    }
                    store_pair(vals_al + 2u * p, y0, y1);
                }
            }
            // PAD pairs: both tokens beyond SEP
            const uint32_t p_pad = (n_emit + 2u + off + 1u) >> 1;         // smallest p with t0 >= n_emit + 2
            const uint32_t p_full_end = (L + off) >> 1;                   // pairs whose second token is < L
            for (uint32_t p = p_pad + tid; p < p_full_end; p += kFastThreads) {
                SLAFB_CHECK(2u * p + 1u - off < L && 2u * p >= off, 212, p);
                store_pair(ids_al + 2u * p, 0ll, 0ll);
                if constexpr (SCGPT) store_pair(vals_al + 2u * p, 0ll, 0ll);
            }
            // boundary pairs (CLS, the last gene / SEP / first PAD, row ends): generic, a handful per row
            {
                const uint32_t nb_head = min(p_lo, n_pairs);                          // pairs [0, p_lo)
                const uint32_t mid_lo = max(p_hi, p_lo), mid_hi = min(max(p_pad, mid_lo), n_pairs);   // pairs [p_hi, p_pad)
                const uint32_t tail_lo = max(p_full_end, mid_hi);                     // pairs [p_full_end, n_pairs)
                const uint32_t n_b = nb_head + (mid_hi - mid_lo) + (n_pairs - tail_lo);
                for (uint32_t i = tid; i < n_b; i += kFastThreads) {
                    uint32_t p;
                    if (i < nb_head) p = i;
                    else if (i < nb_head + (mid_hi - mid_lo)) p = mid_lo + (i - nb_head);
                    else p = tail_lo + (i - nb_head - (mid_hi - mid_lo));
                    const long long t0 = (long long)(2u * p) - (long long)off, t1 = t0 + 1;
                    const bool v0 = t0 >= 0, v1 = t1 < (long long)L;
                    const long long x0 = v0 ? id_tok(t0) : 0ll, x1 = v1 ? id_tok(t1) : 0ll;
                    if (v0 && v1) *reinterpret_cast<longlong2 *>(ids_al + 2u * p) = make_longlong2(x0, x1);
                    else if (v0) ids_al[2u * p] = x0;
                    else if (v1) ids_al[2u * p + 1] = x1;
                    if constexpr (SCGPT) {
                        const long long y0 = v0 ? val_tok(t0) : 0ll, y1 = v1 ? val_tok(t1) : 0ll;
                        if (v0 && v1) *reinterpret_cast<longlong2 *>(vals_al + 2u * p) = make_longlong2(y0, y1);
                        else if (v0) vals_al[2u * p] = y0;
                        else if (v1) vals_al[2u * p + 1] = y1;
                    }
                }
            }
            fill_mask(a.mask, A, L, min(n_emit + 2u, L));
            if (tid == 0) {
                const uint32_t cid = __ldg(a.seg_cell + d.c);
                a.cell_ids[j] = a.cell_signed ? (long long)(int32_t)cid : (long long)cid;
            }
        }
        if constexpr (PCT) {
            if (dirty) {                                // (uniform) the next cell starts from an empty bitmap again
                for (uint32_t w = tid; w < (uint32_t)kPctWords; w += kFastThreads) s_bitmap[w] = 0u;
                if (tid == 0) s_dirty = 0u;
            }
        }
        __syncthreads();   // every thread is done with stage `st` and the scratch before they are reused
    }
}

// ---- K2w: percentile mode, ONE WARP PER CELL -------------------------------------------------------------------------------
// The CTA-per-cell kernel above is issue-bound in percentile mode: 256 threads share ~2,100 rows, so the per-cell fixed work (block
// scans, cross-warp reductions through shared memory, barriers, ragged-chunk handling in every warp) is paid by eight warps for
// eight rows per thread -- 7,300 warp instructions per cell.  Here a warp owns a cell: every reduction is a `redux` / ballot, the
// row-order compaction is a warp prefix per 256 rows with predicated (branch-free) stores, and the emission is the same 128-bit
// store pattern with a stride of 32.  Warps take output rows in order from the same ticket, so the rows in flight still form one
// compact window of the output tensors.
//   The gene segment is bulk-copied into the warp's private stage (the emission reads it in place when nothing is filtered).  The
// values are either bulk-copied as well (VSTAGE) or read with 128-bit loads straight from global memory -- twice, the second time
// from L2 -- which halves the shared memory per warp and doubles the warps per SM.  Of a cell longer than the stage the head is
// staged and the rows beyond it are read from global memory (two loads per lane in flight); their kept genes land behind the
// head's survivors in the gene stage (at most L - 2 genes are ever emitted, and the stage holds that many).  Only a value beyond
// the bitmap sends a cell to the general kernel.
constexpr int kPctWarps = 4;                        // warps per CTA
template <typename GeneT, bool VSTAGE>
__host__ __device__ constexpr uint32_t pct_warp_bytes(uint32_t vcap) {
    return (VSTAGE ? ((vcap * 2u + 127u) & ~127u) : 0u) + ((vcap * (uint32_t)sizeof(GeneT) + 127u) & ~127u) + (uint32_t)kPctWords * 4u;
}

// predicated shared-memory store of one gene id + pointer bump (no branch: the compiler turns `if (bit) { *p++ = g; }` into one)
template <typename GeneT>
__device__ __forceinline__ void sts_if(uint32_t &addr, GeneT g, uint32_t bit) {
    if constexpr (sizeof(GeneT) == 2) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u16 [%0], %1;\n\t@p add.u32 %0, %0, 2;\n\t}"
                     : "+r"(addr) : "h"((uint16_t)g), "r"(bit) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t@p add.u32 %0, %0, 4;\n\t}"
                     : "+r"(addr) : "r"((uint32_t)g), "r"(bit) : "memory");
    }
}

template <typename GeneT, bool VSTAGE>
__global__ void __launch_bounds__(kPctWarps * 32) tokenize_pct_warp_kernel(const FastArgs fa) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[kPctWarps][2];
    const TokArgs &a = fa.t;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t vbytes = VSTAGE ? ((fa.vcap * 2u + 127u) & ~127u) : 0u, gbytes = (fa.vcap * (uint32_t)sizeof(GeneT) + 127u) & ~127u;
    uint8_t *mine = smem_raw + warp * pct_warp_bytes<GeneT, VSTAGE>(fa.vcap);
    uint16_t *sv = reinterpret_cast<uint16_t *>(mine);
    GeneT *sg = reinterpret_cast<GeneT *>(mine + vbytes);
    uint32_t *bitmap = reinterpret_cast<uint32_t *>(mine + vbytes + gbytes);
    uint64_t *bar = &s_bar[warp][0], *bar_v = &s_bar[warp][1];       // gene stage / value stage
    const GeneT *__restrict__ gene_g = static_cast<const GeneT *>(a.gene);
    const uint16_t *__restrict__ value_g = static_cast<const uint16_t *>(a.value);
    const uint32_t R = a.n_rows, k = a.max_items, L = a.L, limit = L - 1u;
    const uint32_t NC = a.n_cells;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t gcap = (fa.vcap >> 3) & ~31u;                            // 8-row groups the stage holds (a multiple of the warp width)
    if (lane == 0) { mbar_init(bar, 1); mbar_init(bar_v, 1); mbar_fence_init(); }
    for (uint32_t w = lane; w < (uint32_t)kPctWords; w += 32u) bitmap[w] = 0u;
    __syncwarp();
    const uint32_t G = gridDim.x * (uint32_t)kPctWarps;
    // the next cell's indices are fetched one cell ahead (ticket -> permutation entry -> segment bounds and cell id), so only the
    // bulk copy itself is waited for in line
    uint32_t jn = blockIdx.x * (uint32_t)kPctWarps + warp, cn = 0, sn = 0, en = 0, cidn = 0;
    if (jn < NC) {
        cn = a.perm ? (uint32_t)__ldg(a.perm + jn) : jn;
        sn = __ldg(a.seg_start + cn); en = __ldg(a.seg_start + cn + 1); cidn = __ldg(a.seg_cell + cn);
    }
    // rows the bulk copies stage for the cell [s_, e_): its whole 8-row groups up to the capacity of the stage
    auto n_staged = [&](uint32_t s_, uint32_t e_) -> uint32_t {
        const uint32_t n_ = e_ - s_, a0_ = s_ & ~7u, ng_ = ((s_ - a0_) + n_ + 7u) >> 3;
        const uint32_t ce_ = (n_ && a0_ + (ng_ << 3) > R) ? 1u : 0u;
        return n_ ? (min(ng_ - ce_, gcap) << 3) : 0u;
    };
    // The value stage is refilled for the NEXT cell as soon as the compaction has read it for the last time, so that copy flies
    // during the emission; the gene stage is refilled at the top of the iteration and its copy flies during pass 1.
    auto stage_values = [&](uint32_t s_, uint32_t e_) {
        const uint32_t b = n_staged(s_, e_);
        if (b) {
            mbar_arrive_expect_tx(bar_v, b * 2u);
            bulk_g2s(sv, value_g + (s_ & ~7u), b * 2u, bar_v);
        }
    };
    if constexpr (VSTAGE) { if (jn < NC && lane == 0) stage_values(sn, en); }
    uint32_t phase = 0u, phase_v = 0u;
    while (jn < NC) {
        const uint32_t j = jn, s = sn, e = en, cid = cidn, n = e - s, a0 = s & ~7u, sh = s - a0;
        // The cell's rows lie in `ngroups` aligned 8-row groups of the columns.  The hot loops below only ever touch groups that
        // can be read with one 128-bit load: all but the column's last partial group (`colend`; it can only be the last group of
        // the last cell), whose rows are visited one by one.  The first `ngs` groups are staged, the others read from global memory.
        const uint32_t ngroups = (sh + n + 7u) >> 3;
        const uint32_t colend = (n && a0 + (ngroups << 3) > R) ? 1u : 0u;
        const uint32_t ghot = ngroups - colend;
        const uint32_t ngs = min(ghot, gcap);
        const uint32_t bulk = n ? (ngs << 3) : 0u;                          // rows staged by the bulk copies
        if (lane == 0) {
            if (bulk) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the stage was written through the generic proxy (in-place compaction)
                mbar_arrive_expect_tx(bar, bulk * (uint32_t)sizeof(GeneT));
                bulk_g2s(sg, gene_g + a0, bulk * (uint32_t)sizeof(GeneT), bar);
            }
            jn = G + atomicAdd(fa.ticket, 1u);
        }
        const uint4 *sv4 = reinterpret_cast<const uint4 *>(sv);
        // values of rows a0 + 8 gi .. a0 + 8 gi + 7 (two per word); gi < ghot
        auto ldv8 = [&](uint32_t gi) -> uint4 {
            if (VSTAGE && gi < ngs) return sv4[gi];
            return __ldg(reinterpret_cast<const uint4 *>(value_g + a0 + (gi << 3)));
        };
        // gene ids of the same rows as raw words: two per word in .a (uint16), or one per word in .a and .b (int32); gi < ghot
        struct G8 { uint4 a, b; };
        auto ldg8 = [&](uint32_t gi) -> G8 {
            G8 r;
            r.b = make_uint4(0u, 0u, 0u, 0u);
            if (gi < ngs) {
                if constexpr (sizeof(GeneT) == 2) r.a = reinterpret_cast<const uint4 *>(sg)[gi];
                else { r.a = reinterpret_cast<const uint4 *>(sg)[2u * gi]; r.b = reinterpret_cast<const uint4 *>(sg)[2u * gi + 1u]; }
            } else {
                const uint4 *gp = reinterpret_cast<const uint4 *>(gene_g + a0 + (gi << 3));
                r.a = __ldg(gp);
                if constexpr (sizeof(GeneT) == 4) r.b = __ldg(gp + 1);
            }
            return r;
        };
        // which of the 8 rows of the group starting at row a0 + p0 belong to the cell (bit i = row a0 + p0 + i); p0 < sh + n
        auto group_mask = [&](uint32_t p0) -> uint32_t {
            const int lo_i = max((int)sh - (int)p0, 0);                     // 0..7: only group 0 starts before the cell
            const int hi_i = min((int)(sh + n) - (int)p0, 8);               // 1..8
            return (0xffu >> (8 - hi_i)) & (0xffu << lo_i) & 0xffu;
        };
        bool defer = false, keep_all = false;
        uint32_t n_keep = n, thr = 0u, dirty = 0u;
        uint32_t mx2 = 0u, m0 = 0u, rare = 0u;
        volatile uint32_t *vb = bitmap;
        // one row (the slow paths): maximum, presence bit -- in the register mask below 32, in the warp's bitmap above
        auto one = [&](uint32_t v) {
            mx2 = __vmaxu2(mx2, v | (v << 16));
            m0 |= shl_clamp(1u, v);
            if (v >= 32u && v < (uint32_t)kPctBits) {
                const uint32_t bit = 1u << (v & 31u);
                dirty = 1u;
                if (!(vb[v >> 5] & bit)) atomicOr(&bitmap[v >> 5], bit);
            }
        };
        auto eight = [&](const uint4 q, uint32_t valid) {
            uint64_t lo = (uint64_t)q.x | ((uint64_t)q.y << 32), hi = (uint64_t)q.z | ((uint64_t)q.w << 32);
#pragma unroll 1
            for (uint32_t i = 0; i < 8u; i++) {
                const uint32_t v = (uint32_t)lo & 0xffffu;
                lo = (lo >> 16) | (hi << 48);
                hi >>= 16;
                if ((valid >> i) & 1u) one(v);
            }
        };
        // ---- pass 1: maximum and presence of the values.  The sweeps are branch-free and cover the groups that lie wholly inside the
        // cell, [gw0, gw1): packed maximum, presence of the values below 32 in a register mask, and a flag if any value is larger.
        const uint32_t gw0 = sh ? 1u : 0u, gw1 = min(ghot, (sh + n) >> 3);
        auto sweep = [&](const uint4 q, uint32_t gi) {
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
            uint32_t bits = 0u;
#pragma unroll
            for (int i = 0; i < 4; i++) bits |= shl_clamp(1u, w4[i] & 0xffffu) | shl_clamp(1u, w4[i] >> 16);
            const uint32_t mx = __vmaxu2(__vmaxu2(q.x, q.y), __vmaxu2(q.z, q.w));
            const bool in = gi >= gw0 && gi < gw1;
            mx2 = __vmaxu2(mx2, in ? mx : 0u);
            m0 |= in ? bits : 0u;
            rare |= in ? ((q.x | q.y | q.z | q.w) & 0xffe0ffe0u) : 0u;
        };
        // the ragged first / last group of the cell goes row by row in lanes 0 / 1 (after the sweeps); without a value stage their
        // loads are issued here, ahead of the sweep
        const uint32_t rgi = lane ? ngroups - 1u : 0u;
        const bool rneed = n && lane < 2u && (lane ? (rgi >= gw1 && rgi >= gw0) : (sh != 0u));
        uint4 rq = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (!VSTAGE) { if (rneed && rgi < ghot) rq = ldv8(rgi); }
        const uint32_t gfirst = VSTAGE ? ngs : 0u;                          // groups [gfirst, ghot): values from global memory
        if (gfirst < ghot) {
            // two 128-bit loads per lane are outstanding while a group is processed (clamped indices: the warp stays converged);
            // this runs while the bulk copy is in flight
            const uint32_t glast = ghot - 1u;
            uint4 qn = ldv8(min(gfirst + lane, glast)), qn2 = ldv8(min(gfirst + lane + 32u, glast));
            for (uint32_t g0 = gfirst; g0 < ghot; g0 += 32u) {
                const uint4 q = qn;
                qn = qn2;
                qn2 = ldv8(min(g0 + lane + 64u, glast));
                sweep(q, g0 + lane);
            }
        }
        jn = __shfl_sync(0xffffffffu, jn, 0);
        if (jn < NC) cn = a.perm ? (uint32_t)__ldg(a.perm + jn) : jn;        // (its latency hides under this cell's work)
        if (VSTAGE && bulk) {
            mbar_wait(bar_v, phase_v);
            phase_v ^= 1u;
        }
        if (n) {
            if constexpr (VSTAGE) {
                for (uint32_t g0 = 0u; g0 < ngs; g0 += 32u) sweep(sv4[min(g0 + lane, ngs - 1u)], g0 + lane);
            }
            // the slow paths, row by row: the ragged first / last group (lanes 0 / 1), the column's last partial group, and -- for
            // cells that have one -- every group with a value of 32 or more
            if (rneed) {
                const uint32_t r0 = a0 + (rgi << 3), valid = group_mask(rgi << 3);
                if (rgi < ghot) {
                    eight(VSTAGE ? ldv8(rgi) : rq, valid);
                } else {
                    for (uint32_t i = 0; i < 8u; i++) if ((valid >> i) & 1u) one((uint32_t)__ldg(value_g + r0 + i));
                }
            }
            __syncwarp();
            if (__any_sync(0xffffffffu, rare != 0u)) {
                for (uint32_t g0 = gw0; g0 < gw1; g0 += 32u) {
                    const uint32_t gi = g0 + lane;
                    if (gi < gw1) {
                        const uint4 q = ldv8(gi);
                        if ((q.x | q.y | q.z | q.w) & 0xffe0ffe0u) eight(q, 0xffu);
                    }
                }
                __syncwarp();
            }
            const uint32_t hi = __reduce_max_sync(0xffffffffu, max(mx2 & 0xffffu, mx2 >> 16));
            m0 = __reduce_or_sync(0xffffffffu, m0);
            dirty = __reduce_or_sync(0xffffffffu, dirty);
            if (hi >= (uint32_t)kPctBits) {
                defer = true;
            } else {
                const uint32_t d0 = __popc(m0);
                constexpr int WPL = kPctWords / 32;                           // bitmap words per lane (contiguous)
                uint32_t D = d0, wv[WPL], pc = 0u;
#pragma unroll
                for (int q2 = 0; q2 < WPL; q2++) wv[q2] = 0u;
                if (dirty) {
                    __syncwarp();
#pragma unroll
                    for (int q2 = 0; q2 < WPL; q2++) {
                        const uint32_t wi = lane * WPL + q2;
                        wv[q2] = (wi >= 1u && wi <= (hi >> 5)) ? bitmap[wi] : 0u;
                        pc += __popc(wv[q2]);
                    }
                    D += __reduce_add_sync(0xffffffffu, pc);
                }
                const uint32_t astar = target_asc_rank(D, k, true, a.min_pct);
                if (astar <= d0) {
                    uint32_t x = m0;
                    for (uint32_t i = 1u; i < astar; i++) x &= x - 1u;
                    thr = (uint32_t)(__ffs(x) - 1);
                } else {                                                      // the (astar - d0)-th set bit of the bitmap's words >= 1
                    const uint32_t want = astar - d0, inc = warp_incl_scan(pc);
                    uint32_t ex = inc - pc, found = 0u;
#pragma unroll
                    for (int q2 = 0; q2 < WPL; q2++) {
                        const uint32_t c2 = __popc(wv[q2]);
                        if (want > ex && want <= ex + c2) {
                            uint32_t x = wv[q2];
                            for (uint32_t i = ex + 1u; i < want; i++) x &= x - 1u;
                            found = ((lane * WPL + q2) << 5) + (uint32_t)(__ffs(x) - 1);
                        }
                        ex += c2;
                    }
                    thr = __reduce_max_sync(0xffffffffu, found);
                }
                keep_all = (astar == 1u);
            }
        }
        if (jn < NC) { sn = __ldg(a.seg_start + cn); en = __ldg(a.seg_start + cn + 1); cidn = __ldg(a.seg_cell + cn); }   // (cn has arrived by now)
        if (bulk) {
            mbar_wait(bar, phase);                                          // the genes (their copy flew during pass 1)
            phase ^= 1u;
        }
        if (n && !defer) {
            {
                if (!keep_all || ngs < ngroups) {
                    // ---- row-order compaction: 256 rows per step (8 per lane), offsets from a warp prefix over the lanes' counts,
                    // IN PLACE in the gene stage (a kept gene never moves to a higher index; the genes of the rows beyond the stage
                    // come from global memory and land behind the staged ones' survivors).  The loads of the next step are issued
                    // before this step is processed: they read rows beyond everything this step writes.
                    const uint32_t thr2 = keep_all ? 0u : (thr | (thr << 16));
                    const uint32_t sg_addr = smem_u32(sg + sh);
                    uint32_t base = 0u;
                    if (ghot) {
                        const uint32_t glast = ghot - 1u;
                        uint4 qvn = ldv8(min(lane, glast));
                        G8 qgn = ldg8(min(lane, glast));
                        for (uint32_t g0 = 0u; g0 < ghot && base < limit; g0 += 32u) {
                            const uint32_t gi = g0 + lane, p0 = gi << 3;
                            const uint4 q = qvn;
                            const G8 gq = qgn;
                            qvn = ldv8(min(gi + 32u, glast));                    // (clamped, not predicated: the warp stays converged)
                            qgn = ldg8(min(gi + 32u, glast));
                            const uint32_t cx = __vcmpgeu2(q.x, thr2), cy = __vcmpgeu2(q.y, thr2), cz = __vcmpgeu2(q.z, thr2), cw = __vcmpgeu2(q.w, thr2);
                            uint32_t flags = (cx & 1u) | ((cx >> 15) & 2u) | ((cy & 1u) << 2) | ((cy >> 13) & 8u) | ((cz & 1u) << 4) | ((cz >> 11) & 32u) |
                                             ((cw & 1u) << 6) | ((cw >> 9) & 128u);
                            flags = gi < ghot ? (flags & group_mask(min(p0, sh + n - 1u))) : 0u;
                            GeneT g[8];
                            if constexpr (sizeof(GeneT) == 2) {
                                const uint32_t g4[4] = {gq.a.x, gq.a.y, gq.a.z, gq.a.w};
#pragma unroll
                                for (int i = 0; i < 8; i++) g[i] = (GeneT)((g4[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
                            } else {
                                g[0] = (GeneT)gq.a.x; g[1] = (GeneT)gq.a.y; g[2] = (GeneT)gq.a.z; g[3] = (GeneT)gq.a.w;
                                g[4] = (GeneT)gq.b.x; g[5] = (GeneT)gq.b.y; g[6] = (GeneT)gq.b.z; g[7] = (GeneT)gq.b.w;
                            }
                            // every lane holds its genes in registers and the warp is converged: the stage may be overwritten below
                            __syncwarp();
                            // exclusive prefix of the lanes' counts (0..8) from four ballots: no dependent shuffle chain
                            const uint32_t pcf = __popc(flags);
                            const uint32_t b0 = __ballot_sync(0xffffffffu, pcf & 1u), b1 = __ballot_sync(0xffffffffu, pcf & 2u),
                                           b2 = __ballot_sync(0xffffffffu, pcf & 4u), b3 = __ballot_sync(0xffffffffu, pcf & 8u);
                            const uint32_t off = base + __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt) + 8u * __popc(b3 & lt);
                            base += __reduce_add_sync(0xffffffffu, pcf);
                            SLAFB_CHECK(off + pcf > limit || sh + off + pcf <= fa.vcap, 231, sh + off);
                            if (off + pcf <= limit) {                             // (all lanes but the one that crosses the emission limit)
                                uint32_t ad = sg_addr + off * (uint32_t)sizeof(GeneT);
#pragma unroll
                                for (int i = 0; i < 8; i++) sts_if<GeneT>(ad, g[i], flags & (1u << i));
                            } else {
                                uint32_t o2 = off;
#pragma unroll
                                for (int i = 0; i < 8; i++) {
                                    if (flags & (1u << i)) {
                                        if (o2 < limit) sg[sh + o2] = g[i];
                                        o2++;
                                    }
                                }
                            }
                        }
                    }
                    if (colend && base < limit) {
                        // the rows of the column's last partial group, one by one (the last cell of a batch only)
                        if (lane == 0) {
                            for (uint32_t r = max(s, a0 + (ghot << 3)); r < e; r++) {
                                if ((uint32_t)__ldg(value_g + r) >= (keep_all ? 0u : thr)) {
                                    if (base < limit) sg[sh + base] = __ldg(gene_g + r);
                                    base++;
                                }
                            }
                        }
                        base = __shfl_sync(0xffffffffu, base, 0);
                    }
                    n_keep = base;
                }
            }
        }
        __syncwarp();                                                       // the value stage has been read for the last time
        if constexpr (VSTAGE) { if (jn < NC && lane == 0) stage_values(sn, en); }
        if (defer) {
            if (lane == 0) {
                const uint32_t wi = atomicAdd(a.work_counters, 1u);
                SLAFB_CHECK(wi < NC, 208, wi);
                a.work_list[wi] = j;
            }
        } else {
            // ---- emission: the CTA kernel's store pattern with a stride of 32 (token t of the row at flat index A + t)
            const uint32_t n_emit = min(n_keep, limit);
            const size_t A = (size_t)j * L;
            const uint32_t off = (uint32_t)(A & 1);
            long long *ids_al = a.ids + (A - off);
            const uint32_t n_pairs = (L + off + 1u) >> 1;
            auto id_tok = [&](long long t) -> long long {
                if (t == 0) return 1ll;
                if (t <= (long long)n_emit) {
                    if constexpr (sizeof(GeneT) == 2) return (long long)sg[sh + (uint32_t)t - 1u] + 4ll;
                    else return (long long)(int32_t)sg[sh + (uint32_t)t - 1u] + 4ll;
                }
                return (t == (long long)n_emit + 1) ? 2ll : 0ll;
            };
            const uint32_t p_lo = (off + 2u) >> 1;
            const uint32_t p_hi = n_emit >= 1u ? (n_emit - 1u + off) / 2u + 1u : 0u;
            SLAFB_CHECK(j < NC && sh + n_emit <= fa.vcap, 232, sh + n_emit);
#pragma unroll 2
            for (uint32_t p = p_lo + lane; p < p_hi; p += 32u) {
                const uint32_t gi = sh + 2u * p - off - 1u;
                long long x0, x1;
                if constexpr (sizeof(GeneT) == 2) { x0 = (long long)sg[gi] + 4ll; x1 = (long long)sg[gi + 1] + 4ll; }
                else { x0 = (long long)(int32_t)sg[gi] + 4ll; x1 = (long long)(int32_t)sg[gi + 1] + 4ll; }
                store_pair(ids_al + 2u * p, x0, x1);
            }
            const uint32_t p_pad = (n_emit + 2u + off + 1u) >> 1;
            const uint32_t p_full_end = (L + off) >> 1;
            for (uint32_t p = p_pad + lane; p < p_full_end; p += 32u) store_pair(ids_al + 2u * p, 0ll, 0ll);
            {
                const uint32_t nb_head = min(p_lo, n_pairs);
                const uint32_t mid_lo = max(p_hi, p_lo), mid_hi = min(max(p_pad, mid_lo), n_pairs);
                const uint32_t tail_lo = max(p_full_end, mid_hi);
                const uint32_t n_b = nb_head + (mid_hi - mid_lo) + (n_pairs - tail_lo);
                for (uint32_t i = lane; i < n_b; i += 32u) {
                    uint32_t p;
                    if (i < nb_head) p = i;
                    else if (i < nb_head + (mid_hi - mid_lo)) p = mid_lo + (i - nb_head);
                    else p = tail_lo + (i - nb_head - (mid_hi - mid_lo));
                    const long long t0 = (long long)(2u * p) - (long long)off, t1 = t0 + 1;
                    const bool v0 = t0 >= 0, v1 = t1 < (long long)L;
                    const long long x0 = v0 ? id_tok(t0) : 0ll, x1 = v1 ? id_tok(t1) : 0ll;
                    if (v0 && v1) *reinterpret_cast<longlong2 *>(ids_al + 2u * p) = make_longlong2(x0, x1);
                    else if (v0) ids_al[2u * p] = x0;
                    else if (v1) ids_al[2u * p + 1] = x1;
                }
            }
            fill_mask_by(a.mask, A, L, min(n_emit + 2u, L), lane, 32u);
            if (lane == 0) a.cell_ids[j] = a.cell_signed ? (long long)(int32_t)cid : (long long)cid;
        }
        if (dirty) {
            for (uint32_t w = lane; w < (uint32_t)kPctWords; w += 32u) bitmap[w] = 0u;
        }
        __syncwarp();          // the stage and the bitmap are free for the next cell
    }
}

// ---- K2f: float32 values, ONE WARP PER CELL (plain Geneformer and both scGPT modes) ---------------------------------------
// The CTA-per-cell kernel is issue-bound on float32 storage for the same reason as in percentile mode (~6,100 warp instructions per
// cell: eight warps pay the per-cell reductions, barriers and ragged-chunk handling for eight rows per thread).  Here a warp owns a
// cell.  Normalised expression values have no integer range to test, so a cell with more rows than max_items proves "at most
// max_items distinct values" (then the dense-rank filter keeps every row) with an open-addressing set of the value bits in the
// warp's shared memory: the sweep looks every value up with plain reads (eight independent loads per lane and step) and only the
// steps in which some lane met a value that is not in the set yet -- a few per cell -- take the inserting path.  A cell with more
// than kF32HashCap distinct values goes to the work list (general kernel); the library falls back to the CTA kernel for data where
// that happens often (slaf_b200.cu).  The cell maximum comes from the same sweep; the binned window + tokenizer collapse to one
// threshold per cell (see the CTA kernel), found by the warp with the same candidate round / 32-ary search.
//   Staging: the genes of the rows that can be emitted, the values of the first `vcap` rows (the emitted rows are among them); the
// values of longer cells' tails come from global memory with two loads per lane in flight.
constexpr int kF32HashBits = 9;
constexpr int kF32HashSlots = 1 << kF32HashBits;    // per warp (2 KiB)
constexpr int kF32HashCap = 255;                    // trusted entries: one inserting step adds at most 256 more, so a free slot always exists
static_assert(kF32HashCap + 256 < kF32HashSlots, "the distinct-value set must never fill up");
template <typename GeneT>
__host__ __device__ constexpr uint32_t f32_warp_bytes(uint32_t vcap, uint32_t gcap) {
    return ((vcap * 4u + 127u) & ~127u) + ((gcap * (uint32_t)sizeof(GeneT) + 127u) & ~127u) + (uint32_t)kF32HashSlots * 4u;
}

template <typename GeneT, int MODE>
__global__ void __launch_bounds__(kPctWarps * 32) tokenize_f32_warp_kernel(const FastArgs fa) {
    constexpr bool SCGPT = (MODE == MODE_SCGPT_RAW || MODE == MODE_SCGPT_BINNED);
    constexpr bool BINNED = (MODE == MODE_SCGPT_BINNED);
    constexpr bool PCT = (MODE == MODE_GF_PCT);
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar[kPctWarps][2];
    const TokArgs &a = fa.t;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t vbytes = (fa.vcap * 4u + 127u) & ~127u, gbytes = (fa.gcap * (uint32_t)sizeof(GeneT) + 127u) & ~127u;
    uint8_t *mine = smem_raw + warp * f32_warp_bytes<GeneT>(fa.vcap, fa.gcap);
    float *sv = reinterpret_cast<float *>(mine);
    GeneT *sg = reinterpret_cast<GeneT *>(mine + vbytes);
    uint32_t *tab = reinterpret_cast<uint32_t *>(mine + vbytes + gbytes);
    uint64_t *bar = &s_bar[warp][0], *bar_g = &s_bar[warp][1];       // value stage / gene stage
    const GeneT *__restrict__ gene_g = static_cast<const GeneT *>(a.gene);
    const float *__restrict__ value_g = static_cast<const float *>(a.value);
    const uint32_t R = a.n_rows, k = a.max_items, L = a.L;
    const uint32_t limit = SCGPT ? a.max_genes : (L - 1u);
    const uint32_t NC = a.n_cells;
    const uint32_t gcv = (fa.vcap >> 3) & ~31u;                             // 8-row groups the value stage holds (a multiple of the warp width)
    const uint32_t gcg = (fa.gcap >> 3) & ~31u;                             // percentile mode: ... and the gene stage (any row may survive the filter)
    const uint32_t lt = (1u << lane) - 1u;
    if (lane == 0) { mbar_init(bar, 1); mbar_init(bar_g, 1); mbar_fence_init(); }
    __syncwarp();
    const uint32_t G = gridDim.x * (uint32_t)kPctWarps;
    uint32_t jn = blockIdx.x * (uint32_t)kPctWarps + warp, cn = 0, sn = 0, en = 0, cidn = 0;
    if (jn < NC) {
        cn = a.perm ? (uint32_t)__ldg(a.perm + jn) : jn;
        sn = __ldg(a.seg_start + cn); en = __ldg(a.seg_start + cn + 1); cidn = __ldg(a.seg_cell + cn);
    }
    uint32_t phase = 0u, phase_g = 0u;
    while (jn < NC) {
        const uint32_t j = jn, s = sn, e = en, cid = cidn, n = e - s, a0 = s & ~7u, sh = s - a0;
        // (group geometry as in the percentile kernel: `colend` = the column's last partial 8-row group, visited row by row)
        const uint32_t ngroups = (sh + n + 7u) >> 3;
        const uint32_t colend = (n && a0 + (ngroups << 3) > R) ? 1u : 0u;
        const uint32_t ghot = ngroups - colend;
        uint32_t n_emit = min(n, limit);                                    // (outside the percentile mode this kernel only emits cells whose every row is kept)
        const uint32_t nge = (sh + n_emit + 7u) >> 3;                       // groups that hold the rows to emit
        bool keep_all = PCT ? false : (n <= k);                             // the percentile rule also filters cells with few rows
        const bool pass1 = n && (BINNED || !keep_all);                      // the values are swept (maximum, distinct count)
        const bool hashing = n && !keep_all;
        const uint32_t ngg = PCT ? min(ghot, gcg) : min(ghot, nge);         // gene groups staged
        const uint32_t ngv = pass1 ? min(ghot, gcv) : (SCGPT ? ngg : 0u);   // value groups staged
        if (lane == 0) {
            if (ngv | ngg) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // (the stages' colend rows are written through the generic proxy)
            if (ngv) { mbar_arrive_expect_tx(bar, ngv * 32u); bulk_g2s(sv, value_g + a0, ngv * 32u, bar); }
            if (ngg) { mbar_arrive_expect_tx(bar_g, ngg * 8u * (uint32_t)sizeof(GeneT)); bulk_g2s(sg, gene_g + a0, ngg * 8u * (uint32_t)sizeof(GeneT), bar_g); }
            jn = G + atomicAdd(fa.ticket, 1u);
        }
        if (hashing) {
#pragma unroll
            for (int i = 0; i < kF32HashSlots / 32; i++) tab[lane + 32 * i] = 0xffffffffu;
        }
        // rows of the column's last partial group that are emitted: into the stages one by one (the last cell of a batch only)
        if (!PCT && colend && ngroups - 1u < nge && lane < 8u) {
            const uint32_t r = a0 + ((ngroups - 1u) << 3) + lane;
            if (r >= s && r < s + n_emit) { sg[r - a0] = __ldg(gene_g + r); if constexpr (SCGPT) sv[r - a0] = __ldg(value_g + r); }
        }
        __syncwarp();
        const float4 *sv4 = reinterpret_cast<const float4 *>(sv);
        volatile uint32_t *vt = tab;
        float xmax = -INFINITY;
        uint32_t ins = 0u;                                                  // keys this lane added to the set
        bool defer = false, live = hashing;                                 // live: the set still decides (uniform)
        float thr_f = 0.0f;                                                 // percentile mode: the smallest value that is kept
        (void)thr_f; (void)lt; (void)gcg;
        auto group_mask = [&](uint32_t p0) -> uint32_t {
            const int lo_i = max((int)sh - (int)p0, 0);
            const int hi_i = min((int)(sh + n) - (int)p0, 8);
            return (0xffu >> (8 - hi_i)) & (0xffu << lo_i) & 0xffu;
        };
        // the inserting probe (slow paths only); set membership is about equality, so the raw bits of x + 0.0f (which folds -0.0
        // into +0.0) are the key
        auto insert = [&](float x) {
            const uint32_t key = __float_as_uint(x + 0.0f);
            uint32_t h = (key * 2654435761u) >> (32 - kF32HashBits);
            for (;;) {
                SLAFB_CHECK(h < (uint32_t)kF32HashSlots, 240, h);
                uint32_t cur = vt[h];
                if (cur == 0xffffffffu) {
                    cur = atomicCAS(tab + h, 0xffffffffu, key);
                    if (cur == 0xffffffffu) { ins++; break; }
                }
                if (cur == key) break;
                h = (h + 1u) & (uint32_t)(kF32HashSlots - 1);
            }
        };
        auto eight = [&](const float4 qa, const float4 qb, uint32_t valid, bool ins_too) {
            const float x8[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if ((valid >> i) & 1u) {
                    xmax = fmaxf(xmax, x8[i]);
                    if (ins_too) insert(x8[i]);
                }
            }
        };
        // one step of the sweep: 8 values per lane; branch-free look-ups, then -- only if some lane missed -- the inserting path
        const uint32_t gw0 = sh ? 1u : 0u, gw1 = min(ghot, (sh + n) >> 3);  // the groups that lie wholly inside the cell
        auto sweep = [&](const float4 qa, const float4 qb, uint32_t gi) {
            const bool in = gi >= gw0 && gi < gw1;
            const float mx = fmaxf(fmaxf(fmaxf(qa.x, qa.y), fmaxf(qa.z, qa.w)), fmaxf(fmaxf(qb.x, qb.y), fmaxf(qb.z, qb.w)));
            xmax = in ? fmaxf(xmax, mx) : xmax;
            if (live) {
                // a value counts as seen when it sits in its HOME slot of the set.  Keys that linear probing displaced miss every time,
                // but the frequent values of a cell arrive first, when the set is all but empty, and are hardly ever displaced.
                const float x8[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                uint32_t miss = 0u;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint32_t key = __float_as_uint(x8[i] + 0.0f);
                    miss |= (vt[(key * 2654435761u) >> (32 - kF32HashBits)] != key) ? (1u << i) : 0u;
                }
                miss = in ? miss : 0u;
                if (__any_sync(0xffffffffu, miss != 0u)) {
#pragma unroll 1
                    while (miss) {                                          // (one or two values in a few lanes, typically)
                        const uint32_t i = (uint32_t)__ffs(miss) - 1u;
                        miss &= miss - 1u;
                        insert(i < 4u ? (i < 2u ? (i ? qa.y : qa.x) : (i == 2u ? qa.z : qa.w)) : (i < 6u ? (i == 4u ? qb.x : qb.y) : (i == 6u ? qb.z : qb.w)));
                    }
                    if (__reduce_add_sync(0xffffffffu, ins) > (uint32_t)kF32HashCap) live = false;   // too many distinct values for this set: deferred
                }
            }
        };
        const uint32_t rgi = lane ? ngroups - 1u : 0u;
        const bool rneed = pass1 && lane < 2u && (lane ? (rgi >= gw1 && rgi >= gw0) : (sh != 0u));
        if (pass1 && ngv < ghot) {
            // values beyond the stage, from global memory while the bulk copy is in flight (clamped indices, uniform trip count)
            const uint32_t glast = ghot - 1u;
            auto ldv = [&](uint32_t gi, float4 &qa, float4 &qb) {
                const float4 *pp = reinterpret_cast<const float4 *>(value_g + a0 + (min(gi, glast) << 3));
                qa = __ldg(pp); qb = __ldg(pp + 1);
            };
            float4 na, nb, ma, mb;
            ldv(ngv + lane, na, nb);
            ldv(ngv + lane + 32u, ma, mb);
            for (uint32_t g0 = ngv; g0 < ghot; g0 += 32u) {
                const float4 qa = na, qb = nb;
                na = ma; nb = mb;
                ldv(g0 + lane + 64u, ma, mb);
                sweep(qa, qb, g0 + lane);
            }
        }
        jn = __shfl_sync(0xffffffffu, jn, 0);
        if (jn < NC) cn = a.perm ? (uint32_t)__ldg(a.perm + jn) : jn;        // (its latency hides under this cell's work)
        if (ngv) {
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        uint32_t kmax = 0u;
        if (pass1) {
            for (uint32_t g0 = 0u; g0 < ngv; g0 += 32u) {
                const uint32_t gi = min(g0 + lane, ngv - 1u);
                sweep(sv4[2u * gi], sv4[2u * gi + 1u], g0 + lane);
            }
            // the ragged first / last group (lanes 0 / 1) and the column's last partial group, row by row
            if (rneed) {
                const uint32_t r0 = a0 + (rgi << 3), valid = group_mask(rgi << 3);
                if (rgi < ngv) {
                    eight(sv4[2u * rgi], sv4[2u * rgi + 1u], valid, live);
                } else if (rgi < ghot) {
                    const float4 *pp = reinterpret_cast<const float4 *>(value_g + r0);
                    eight(__ldg(pp), __ldg(pp + 1), valid, live);
                } else {
                    for (uint32_t i = 0; i < 8u; i++) {
                        if ((valid >> i) & 1u) { const float x = __ldg(value_g + r0 + i); xmax = fmaxf(xmax, x); if (live) insert(x); }
                    }
                }
            }
            __syncwarp();
            kmax = __reduce_max_sync(0xffffffffu, key_of(xmax));
            if (hashing) {
                const uint32_t D = __reduce_add_sync(0xffffffffu, ins);
                if constexpr (PCT) {
                    // Percentile branch (aggregators.py:167-196): the set holds every distinct value of the cell; keep the rows whose
                    // value is at least the a*-th smallest of them.  That value is found by a* rounds of "smallest key not below
                    // the bound" over the set (16 slots per lane, one warp minimum per round; a* is a percentile of D: small).
                    if (!live || D > (uint32_t)kF32HashCap) {
                        defer = true;
                    } else {
                        const uint32_t astar = target_asc_rank(D, k, true, a.min_pct);
                        keep_all = (astar == 1u);
                        if (!keep_all) {
                            uint32_t ok[kF32HashSlots / 32];
#pragma unroll
                            for (int i = 0; i < kF32HashSlots / 32; i++) {
                                const uint32_t bits = vt[lane + 32 * i];
                                ok[i] = bits == 0xffffffffu ? 0xffffffffu : key_of(__uint_as_float(bits));
                            }
                            uint32_t lowb = 0u, sel = 0u;
                            for (uint32_t t = 0; t < astar; t++) {
                                uint32_t m = 0xffffffffu;
#pragma unroll
                                for (int i = 0; i < kF32HashSlots / 32; i++) m = ok[i] >= lowb ? min(m, ok[i]) : m;
                                sel = __reduce_min_sync(0xffffffffu, m);
                                lowb = sel + 1u;
                            }
                            thr_f = float_of_key(sel);
                        }
                    }
                } else {
                    if (live && D <= min(k, (uint32_t)kF32HashCap)) keep_all = true; else defer = true;   // D <= k distinct values: every row has rank <= k
                }
            }
        }
        if (jn < NC) { sn = __ldg(a.seg_start + cn); en = __ldg(a.seg_start + cn + 1); cidn = __ldg(a.seg_cell + cn); }
        if (ngg) {
            mbar_wait(bar_g, phase_g);                                      // the genes (their copy flew during the sweep)
            phase_g ^= 1u;
        }
        if constexpr (PCT) {
            if (n && !defer && (!keep_all || ngg < ngroups)) {
                // ---- row-order compaction, IN PLACE in the gene stage (the percentile kernel's: 256 rows per step, offsets from a
                // ballot prefix, predicated stores; rows beyond the stages come from global memory, next step's loads in flight)
                struct V8 { float4 a, b; };
                struct G8 { uint4 a, b; };
                auto ldv8 = [&](uint32_t gi) -> V8 {
                    V8 r;
                    if (gi < ngv) { r.a = sv4[2u * gi]; r.b = sv4[2u * gi + 1u]; }
                    else { const float4 *pp = reinterpret_cast<const float4 *>(value_g + a0 + (gi << 3)); r.a = __ldg(pp); r.b = __ldg(pp + 1); }
                    return r;
                };
                auto ldg8 = [&](uint32_t gi) -> G8 {
                    G8 r;
                    r.b = make_uint4(0u, 0u, 0u, 0u);
                    if (gi < ngg) {
                        if constexpr (sizeof(GeneT) == 2) r.a = reinterpret_cast<const uint4 *>(sg)[gi];
                        else { r.a = reinterpret_cast<const uint4 *>(sg)[2u * gi]; r.b = reinterpret_cast<const uint4 *>(sg)[2u * gi + 1u]; }
                    } else {
                        const uint4 *gp = reinterpret_cast<const uint4 *>(gene_g + a0 + (gi << 3));
                        r.a = __ldg(gp);
                        if constexpr (sizeof(GeneT) == 4) r.b = __ldg(gp + 1);
                    }
                    return r;
                };
                const float thr = keep_all ? -INFINITY : thr_f;
                const uint32_t sg_addr = smem_u32(sg + sh);
                uint32_t base = 0u;
                if (ghot) {
                    const uint32_t glast = ghot - 1u;
                    V8 qvn = ldv8(min(lane, glast));
                    G8 qgn = ldg8(min(lane, glast));
                    for (uint32_t g0 = 0u; g0 < ghot && base < limit; g0 += 32u) {
                        const uint32_t gi = g0 + lane, p0 = gi << 3;
                        const V8 q = qvn;
                        const G8 gq = qgn;
                        qvn = ldv8(min(gi + 32u, glast));
                        qgn = ldg8(min(gi + 32u, glast));
                        uint32_t flags = (q.a.x >= thr ? 1u : 0u) | (q.a.y >= thr ? 2u : 0u) | (q.a.z >= thr ? 4u : 0u) | (q.a.w >= thr ? 8u : 0u) |
                                         (q.b.x >= thr ? 16u : 0u) | (q.b.y >= thr ? 32u : 0u) | (q.b.z >= thr ? 64u : 0u) | (q.b.w >= thr ? 128u : 0u);
                        flags = gi < ghot ? (flags & group_mask(min(p0, sh + n - 1u))) : 0u;
                        GeneT g[8];
                        if constexpr (sizeof(GeneT) == 2) {
                            const uint32_t g4[4] = {gq.a.x, gq.a.y, gq.a.z, gq.a.w};
#pragma unroll
                            for (int i = 0; i < 8; i++) g[i] = (GeneT)((g4[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
                        } else {
                            g[0] = (GeneT)gq.a.x; g[1] = (GeneT)gq.a.y; g[2] = (GeneT)gq.a.z; g[3] = (GeneT)gq.a.w;
                            g[4] = (GeneT)gq.b.x; g[5] = (GeneT)gq.b.y; g[6] = (GeneT)gq.b.z; g[7] = (GeneT)gq.b.w;
                        }
                        __syncwarp();                                         // every lane holds its genes in registers: the stage may be overwritten
                        const uint32_t pcf = __popc(flags);
                        const uint32_t b0 = __ballot_sync(0xffffffffu, pcf & 1u), b1 = __ballot_sync(0xffffffffu, pcf & 2u),
                                       b2 = __ballot_sync(0xffffffffu, pcf & 4u), b3 = __ballot_sync(0xffffffffu, pcf & 8u);
                        const uint32_t off = base + __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt) + 8u * __popc(b3 & lt);
                        base += __reduce_add_sync(0xffffffffu, pcf);
                        SLAFB_CHECK(off + pcf > limit || sh + off + pcf <= fa.gcap, 242, sh + off);
                        if (off + pcf <= limit) {
                            uint32_t ad = sg_addr + off * (uint32_t)sizeof(GeneT);
#pragma unroll
                            for (int i = 0; i < 8; i++) sts_if<GeneT>(ad, g[i], flags & (1u << i));
                        } else {
                            uint32_t o2 = off;
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                if (flags & (1u << i)) {
                                    if (o2 < limit) sg[sh + o2] = g[i];
                                    o2++;
                                }
                            }
                        }
                    }
                }
                if (colend && base < limit) {
                    if (lane == 0) {
                        for (uint32_t r = max(s, a0 + (ghot << 3)); r < e; r++) {
                            if (__ldg(value_g + r) >= thr) {
                                if (base < limit) sg[sh + base] = __ldg(gene_g + r);
                                base++;
                            }
                        }
                    }
                    base = __shfl_sync(0xffffffffu, base, 0);
                }
                n_emit = min(base, limit);
                __syncwarp();
            }
        }
        if (defer) {
            if (lane == 0) {
                const uint32_t wi = atomicAdd(a.work_counters, 1u);
                SLAFB_CHECK(wi < NC, 208, wi);
                a.work_list[wi] = j;
            }
        } else {
            // ---- emission: token pairs (128-bit stores), value stream, PAD fill, mask, cell id
            const size_t A = (size_t)j * L;
            const long long top = a.vocab + (long long)(a.n_bins - 1);
            float xstar_f = 0.0f;
            (void)top; (void)xstar_f; (void)kmax;
            if constexpr (BINNED) {
                // token = (bin >= 1) ? top : PAD with bin(x) = floor(log1pf(x) * n / m) monotone in x: the warp finds the smallest float
                // whose bin is >= 1 (one round of 32 candidates around a closed-form guess, a 32-ary search only if that misses)
                const float m_f32 = log1pf_window(float_of_key(kmax));
                const float wb = (float)a.w_bins;
                auto pred = [&](uint32_t c) -> bool {
                    const float lv = log1pf_window(float_of_key(c));
                    return lv > 0.0f && __fdiv_rn(__fmul_rn(lv, wb), m_f32) >= 1.0f;
                };
                const uint32_t kzero = key_of(0.0f);
                uint32_t xs = 0xffffffffu;                                  // no value reaches bin 1
                if (a.w_bins >= 2 && kmax > kzero && pred(kmax)) {
                    const float x0 = (float)expm1((double)m_f32 / (double)wb);
                    uint32_t k0 = key_of(x0);
                    k0 = min(max(k0, kzero + 1u), kmax);
                    const uint32_t wlo = (k0 - kzero > 16u) ? k0 - 16u : kzero + 1u;
                    const uint32_t c = min(wlo + lane, kmax);
                    const uint32_t hit = __ballot_sync(0xffffffffu, pred(c));
                    if (hit == 0u) xs = warp_lower_bound(wlo + 32u, kmax, pred);
                    else if ((hit & 1u) && wlo > kzero + 1u) xs = warp_lower_bound(kzero + 1u, wlo, pred);
                    else xs = min(wlo + (uint32_t)(__ffs(hit) - 1), kmax);
                }
                xstar_f = float_of_key(xs);                                 // ("no value": a NaN, which compares false as it must)
            }
            auto id_tok = [&](long long t) -> long long {
                if (t == 0) return 1ll;
                if (t <= (long long)n_emit) {
                    if constexpr (sizeof(GeneT) == 2) return (long long)sg[sh + (uint32_t)t - 1u] + 4ll;
                    else return (long long)(int32_t)sg[sh + (uint32_t)t - 1u] + 4ll;
                }
                return (t == (long long)n_emit + 1) ? 2ll : 0ll;
            };
            auto val_of = [&](float x) -> long long {
                if constexpr (BINNED) return (x >= xstar_f) ? top : 0ll;
                else return expr_bin_token(x, a.bin_size, a.n_bins, a.vocab);
            };
            auto val_tok = [&](long long t) -> long long {
                if (t < 1 || t > (long long)n_emit) return 0ll;
                return val_of(sv[sh + (uint32_t)t - 1u]);
            };
            const uint32_t off = (uint32_t)(A & 1);
            long long *ids_al = a.ids + (A - off);
            long long *vals_al = SCGPT ? a.vals + (A - off) : nullptr;
            const uint32_t n_pairs = (L + off + 1u) >> 1;
            const uint32_t p_lo = (off + 2u) >> 1;
            const uint32_t p_hi = n_emit >= 1u ? (n_emit - 1u + off) / 2u + 1u : 0u;
            SLAFB_CHECK(j < NC && sh + n_emit <= fa.gcap && (!SCGPT || sh + n_emit <= fa.vcap), 241, sh + n_emit);
#pragma unroll 2
            for (uint32_t p = p_lo + lane; p < p_hi; p += 32u) {
                const uint32_t gi = sh + 2u * p - off - 1u;
                long long x0, x1;
                if constexpr (sizeof(GeneT) == 2) { x0 = (long long)sg[gi] + 4ll; x1 = (long long)sg[gi + 1] + 4ll; }
                else { x0 = (long long)(int32_t)sg[gi] + 4ll; x1 = (long long)(int32_t)sg[gi + 1] + 4ll; }
                store_pair(ids_al + 2u * p, x0, x1);
            }
            if constexpr (SCGPT) {
#pragma unroll 2
                for (uint32_t p = p_lo + lane; p < p_hi; p += 32u) {
                    const uint32_t gi = sh + 2u * p - off - 1u;
                    store_pair(vals_al + 2u * p, val_of(sv[gi]), val_of(sv[gi + 1]));
                }
            }
            const uint32_t p_pad = (n_emit + 2u + off + 1u) >> 1;
            const uint32_t p_full_end = (L + off) >> 1;
            for (uint32_t p = p_pad + lane; p < p_full_end; p += 32u) {
                store_pair(ids_al + 2u * p, 0ll, 0ll);
                if constexpr (SCGPT) store_pair(vals_al + 2u * p, 0ll, 0ll);
            }
            {
                const uint32_t nb_head = min(p_lo, n_pairs);
                const uint32_t mid_lo = max(p_hi, p_lo), mid_hi = min(max(p_pad, mid_lo), n_pairs);
                const uint32_t tail_lo = max(p_full_end, mid_hi);
                const uint32_t n_b = nb_head + (mid_hi - mid_lo) + (n_pairs - tail_lo);
                for (uint32_t i = lane; i < n_b; i += 32u) {
                    uint32_t p;
                    if (i < nb_head) p = i;
                    else if (i < nb_head + (mid_hi - mid_lo)) p = mid_lo + (i - nb_head);
                    else p = tail_lo + (i - nb_head - (mid_hi - mid_lo));
                    const long long t0 = (long long)(2u * p) - (long long)off, t1 = t0 + 1;
                    const bool v0 = t0 >= 0, v1 = t1 < (long long)L;
                    const long long x0 = v0 ? id_tok(t0) : 0ll, x1 = v1 ? id_tok(t1) : 0ll;
                    if (v0 && v1) *reinterpret_cast<longlong2 *>(ids_al + 2u * p) = make_longlong2(x0, x1);
                    else if (v0) ids_al[2u * p] = x0;
                    else if (v1) ids_al[2u * p + 1] = x1;
                    if constexpr (SCGPT) {
                        const long long y0 = v0 ? val_tok(t0) : 0ll, y1 = v1 ? val_tok(t1) : 0ll;
                        if (v0 && v1) *reinterpret_cast<longlong2 *>(vals_al + 2u * p) = make_longlong2(y0, y1);
                        else if (v0) vals_al[2u * p] = y0;
                        else if (v1) vals_al[2u * p + 1] = y1;
                    }
                }
            }
            fill_mask_by(a.mask, A, L, min(n_emit + 2u, L), lane, 32u);
            if (lane == 0) a.cell_ids[j] = a.cell_signed ? (long long)(int32_t)cid : (long long)cid;
        }
        __syncwarp();          // the stages and the set are free for the next cell
    }
}

// table for the binned window on uint16 storage: for a cell whose maximum value is vmax, the
// smallest value v with  floor(log1p(v) * n_bins / log1p(vmax)) >= 1  (f64, evaluation order of
// aggregators.py:100-105).  Every such row gets the token vocab + n_bins - 1, every other row PAD
// (the float branch of ScGPTTokenizer applied to an integral bin >= 1 always clips to the top bin).
__global__ void vstar_table_kernel(const double *__restrict__ lut, int n_bins, uint32_t *__restrict__ table) {
    const uint32_t vmax = blockIdx.x * blockDim.x + threadIdx.x;
    if (vmax > 65535u) return;
    if (n_bins < 2 || vmax == 0u) { table[vmax] = 0xffffffffu; return; }
    const double m = lut[vmax], nb = (double)n_bins;
    uint32_t lo = 1u, hi = vmax;                 // the predicate holds at vmax and is monotone in v
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const bool ok = __ddiv_rn(__dmul_rn(lut[mid], nb), m) >= 1.0;
        if (ok) hi = mid; else lo = mid + 1u;
    }
    table[vmax] = lo;
}

// general path: persistent CTAs over the work list (or over all cells)
template <typename GeneT, typename ValT, int MODE, int SINK>
__global__ void __launch_bounds__(kTokThreads) tokenize_general_kernel(const TokArgs a, const WindowSink wsnk) {
    __shared__ TokSmem sm;
    __shared__ uint32_t s_item;
    extern __shared__ __align__(16) uint32_t dyn[];   // bitmap (u16) or sort keys (f32)
    const uint32_t count = a.all_cells ? a.n_cells : a.work_counters[0];
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.host_slow) *a.host_slow = count;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(a.work_counters + 1, 1u);
        __syncthreads();
        const uint32_t it = s_item;
        if (it >= count) break;
        const uint32_t row = a.all_cells ? it : a.work_list[it];
        process_cell<GeneT, ValT, MODE, true, SINK>(a, wsnk, row, sm, dyn, dyn);
    }
}

// ------------------------------------------------------------------------------------------
// K6 (extension, BASELINE.json config 4): Geneformer rank-value encoding with per-gene normalisation
// ------------------------------------------------------------------------------------------
// Not in the reference (it has no normalisation and keeps row order).  Per cell:
//   x_i   = float32(value_i) / norm[gene_i]          (norm == nullptr: x_i = value_i)
//   keep  = rows whose dense rank of x (descending) is <= max_items        -- the reference's filter
//   order = rank_order ? descending x, ties in row order (stable)  :  row order
// then the Geneformer framing (CLS, gene + 4, SEP, truncate to max_genes, PAD, mask).
// One CTA per cell, persistent over a ticket; the cell's rows are sorted as 64-bit composites
// (~key << 32 | row) with a bitonic network in shared memory (global scratch beyond 8192 rows).
struct RankArgs {
    TokArgs t;
    const float *norm;               // [n_norm] or nullptr
    uint32_t n_norm;
    int rank_order;
    unsigned long long *scratch64;   // [n_rows] composites of cells larger than kRankSmemKeys
    int from_list;                   // tokenize_rank_kernel: take the output rows from the work list (what the bucket kernel left over)
};
constexpr int kRankSmemKeys = 4096;       // composites sorted in registers + shared memory up to here (32 KiB per CTA: six CTAs per SM)

__device__ void bitonic_sort_asc_u64(unsigned long long *keys, uint32_t n) { bitonic_sort_asc_t<unsigned long long>(keys, n); }

// Bitonic sort of E * kTokThreads 64-bit keys in shared memory (callers pad with ~0), ascending, with the network tiled over the
// memory hierarchy: thread t owns the E contiguous elements [t * E, t * E + E) in REGISTERS.  Strides below E are compare-
// exchanges between a thread's own registers, strides below 32 * E are exchanged with another lane by warp shuffles, and only
// the strides that cross warps go through shared memory (two barriers each).  For 4,096 keys that is 6 shared-memory steps
// instead of 78 barrier-separated passes over shared memory.
// Physical position of logical element i: one slot of padding after every 16 elements, so that a warp's accesses to
// "element r of every thread" (stride E elements = up to 128 bytes) spread over all banks instead of hitting one.
__device__ __forceinline__ uint32_t rk_phys(uint32_t i) { return i + (i >> 4); }
constexpr int kRankSmemSlots = kRankSmemKeys + kRankSmemKeys / 16;

template <int E>
__device__ __forceinline__ void bitonic_sort_tiled(unsigned long long *sm) {
    constexpr uint32_t N = (uint32_t)E * kTokThreads;
    const uint32_t base = threadIdx.x * (uint32_t)E;
    unsigned long long v[E];
#pragma unroll
    for (int r = 0; r < E; r++) v[r] = sm[rk_phys(base + r)];
    for (uint32_t k = 2; k <= N; k <<= 1) {
        uint32_t j = k >> 1;
        for (; j >= (uint32_t)E * 32u; j >>= 1) {                       // partner in another warp
            __syncthreads();                                            // everyone has read what the previous step left in sm
#pragma unroll
            for (int r = 0; r < E; r++) sm[rk_phys(base + r)] = v[r];
            __syncthreads();
#pragma unroll
            for (int r = 0; r < E; r++) {
                const uint32_t i = base + r;
                const unsigned long long pv = sm[rk_phys(i ^ j)];
                const bool keep_min = (((i & j) == 0u) == ((i & k) == 0u));
                v[r] = keep_min ? (pv < v[r] ? pv : v[r]) : (pv > v[r] ? pv : v[r]);
            }
        }
        for (; j >= (uint32_t)E; j >>= 1) {                             // partner in another lane of this warp
            const uint32_t lm = j / (uint32_t)E;
            const bool keep_min = (((base & j) == 0u) == ((base & k) == 0u));   // (k > j >= E: neither bit depends on r)
#pragma unroll
            for (int r = 0; r < E; r++) {
                const unsigned long long pv = __shfl_xor_sync(0xffffffffu, v[r], lm);
                v[r] = keep_min ? (pv < v[r] ? pv : v[r]) : (pv > v[r] ? pv : v[r]);
            }
        }
#pragma unroll
        for (int jj = E / 2; jj >= 1; jj >>= 1) {                       // partner in my own registers
            if ((uint32_t)jj <= (k >> 1)) {
#pragma unroll
                for (int r = 0; r < E; r++) {
                    if ((r & jj) == 0) {
                        const bool asc = (((base + (uint32_t)r) & k) == 0u);
                        const unsigned long long x = v[r], y = v[r | jj];
                        const bool sw = (x > y) == asc;
                        v[r] = sw ? y : x;
                        v[r | jj] = sw ? x : y;
                    }
                }
            }
        }
    }
    // back to the plain (unpadded) order the callers index: through the padded layout once more, then every thread moves the
    // elements tid, tid + 256, ... (consecutive addresses on both sides)
    __syncthreads();
#pragma unroll
    for (int r = 0; r < E; r++) sm[rk_phys(base + r)] = v[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < E; r++) v[r] = sm[rk_phys(threadIdx.x + (uint32_t)r * kTokThreads)];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < E; r++) sm[threadIdx.x + (uint32_t)r * kTokThreads] = v[r];
    __syncthreads();
}

template <typename GeneT, typename ValT>
__device__ __forceinline__ uint32_t rank_key(const RankArgs &ra, GeneT g, ValT v) {
    if (ra.norm == nullptr) return key_of(v);
    const uint32_t gi = (uint32_t)g;
    const float d = gi < ra.n_norm ? __ldg(ra.norm + gi) : 1.0f;
    return key_of(__fdiv_rn((float)v, d));
}

// after the sort: `keys` holds the cell's composites (~key << 32 | row) in ascending order.  Dense-rank cut, emission in rank or
// row order, CLS / SEP / PAD / mask / cell id.  Whole CTA; s_ws = 32 words, s_bcast = 4 words of shared scratch.
template <typename GeneT, typename ValT>
__device__ __forceinline__ void rank_emit(const RankArgs &ra, const unsigned long long *keys, uint32_t j, uint32_t c, uint32_t s, uint32_t n,
                                          uint32_t *s_ws, uint32_t *s_bcast) {
    const TokArgs &a = ra.t;
    const int tid = threadIdx.x;
    const GeneT *__restrict__ gene = static_cast<const GeneT *>(a.gene);
    const ValT *__restrict__ value = static_cast<const ValT *>(a.value);
    const uint32_t k = a.max_items, L = a.L, limit = L - 1u;
    // m = rows with dense rank <= k: the prefix before the (k+1)-th distinct key.  Thread t counts the distinct-key heads of its
    // contiguous slice of the sorted array, ONE block scan places the slices, and the thread whose slice holds head number k
    // (0-based) walks it once more to find the position.
    if (tid == 0) s_bcast[0] = n;
    const uint32_t per = (n + kTokThreads - 1u) / kTokThreads;
    const uint32_t i0 = min((uint32_t)tid * per, n), i1 = min(i0 + per, n);
    uint32_t heads = 0u;
    {
        uint32_t prev = i0 ? (uint32_t)(keys[i0 - 1u] >> 32) : 0u;
        for (uint32_t i = i0; i < i1; i++) {
            const uint32_t cur = (uint32_t)(keys[i] >> 32);
            heads += (i == 0u || cur != prev) ? 1u : 0u;
            prev = cur;
        }
    }
    uint32_t tot_heads;
    const uint32_t ex = block_excl_scan(heads, s_ws, tot_heads);           // (its barriers also publish s_bcast[0] = n)
    if (ex <= k && k < ex + heads) {
        uint32_t run = ex, prev = i0 ? (uint32_t)(keys[i0 - 1u] >> 32) : 0u;
        for (uint32_t i = i0; i < i1; i++) {
            const uint32_t cur = (uint32_t)(keys[i] >> 32);
            if (i == 0u || cur != prev) {
                if (run == k) { s_bcast[0] = i; break; }
                run++;
            }
            prev = cur;
        }
    }
    __syncthreads();
    const uint32_t m = s_bcast[0];
    const size_t row_base = (size_t)j * L;
    uint32_t n_emit;
    if (ra.rank_order) {
        n_emit = min(m, limit);
        for (uint32_t r = tid; r < n_emit; r += kTokThreads) {
            const GeneT g = __ldg(gene + s + (uint32_t)(keys[r] & 0xffffffffull));
            long long gid;
            if constexpr (sizeof(GeneT) == 2) gid = (long long)g; else gid = (long long)(int32_t)g;
            a.ids[row_base + 1 + r] = gid + 4;
        }
    } else {
        // row order: keep rows whose key is >= the smallest kept key
        const uint32_t thr_inv = m ? (uint32_t)(keys[m - 1] >> 32) : 0u;      // largest ~key kept
        __syncthreads();
        uint32_t base = 0;
        for (uint32_t ib = 0; ib < n && base < limit; ib += kTokThreads) {
            const uint32_t i = ib + tid;
            GeneT g{};
            uint32_t kp = 0;
            if (i < n) {
                g = __ldg(gene + s + i);
                kp = (m && (~rank_key<GeneT, ValT>(ra, g, __ldg(value + s + i)) <= thr_inv)) ? 1u : 0u;
            }
            uint32_t tot;
            const uint32_t off = base + block_excl_scan(kp, s_ws, tot);
            if (kp && off < limit) {
                long long gid;
                if constexpr (sizeof(GeneT) == 2) gid = (long long)g; else gid = (long long)(int32_t)g;
                a.ids[row_base + 1 + off] = gid + 4;
            }
            base += tot;
        }
        n_emit = min(base, limit);
    }
    const uint32_t sep_pos = 1u + n_emit, t_len = min(n_emit + 2u, L);
    if (tid == 0) {
        a.ids[row_base] = 1;
        if (sep_pos < L) a.ids[row_base + sep_pos] = 2;
        const uint32_t cid = __ldg(a.seg_cell + c);
        a.cell_ids[j] = a.cell_signed ? (long long)(int32_t)cid : (long long)cid;
    }
    fill_zero_i64(a.ids, row_base + t_len, row_base + L);
    fill_mask(a.mask, row_base, L, t_len);
}

template <typename GeneT, typename ValT>
__global__ void __launch_bounds__(kTokThreads) tokenize_rank_kernel(const RankArgs ra) {
    extern __shared__ __align__(16) unsigned long long rk_keys[];
    __shared__ uint32_t s_ws[32], s_bcast[4], s_item;
    const TokArgs &a = ra.t;
    const int tid = threadIdx.x;
    const GeneT *__restrict__ gene = static_cast<const GeneT *>(a.gene);
    const ValT *__restrict__ value = static_cast<const ValT *>(a.value);
    const uint32_t count = ra.from_list ? a.work_counters[0] : a.n_cells;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(a.work_counters + 1, 1u);
        __syncthreads();
        if (s_item >= count) break;
        const uint32_t j = ra.from_list ? a.work_list[s_item] : s_item;
        if (j == 0xffffffffu) continue;                                     // (tokenized by the second bucket launch)
        const uint32_t c = a.perm ? (uint32_t)a.perm[j] : j;
        const uint32_t s = __ldg(a.seg_start + c), e = __ldg(a.seg_start + c + 1), n = e - s;
        unsigned long long *keys = (n <= (uint32_t)kRankSmemKeys) ? rk_keys : (ra.scratch64 + s);
        if (n <= (uint32_t)kRankSmemKeys) {
            // composites into shared memory, padded with ~0 (sorts last) up to the next power of two >= 256 per the tiling
            uint32_t P = kTokThreads;
            while (P < n) P <<= 1;
            for (uint32_t i = tid; i < P; i += kTokThreads) {
                unsigned long long c = ~0ull;
                if (i < n) {
                    const uint32_t key = rank_key<GeneT, ValT>(ra, __ldg(gene + s + i), __ldg(value + s + i));
                    c = ((unsigned long long)(~key) << 32) | (unsigned long long)i;
                }
                keys[rk_phys(i)] = c;
            }
            __syncthreads();
            switch (P / kTokThreads) {
            case 1: bitonic_sort_tiled<1>(keys); break;
            case 2: bitonic_sort_tiled<2>(keys); break;
            case 4: bitonic_sort_tiled<4>(keys); break;
            case 8: bitonic_sort_tiled<8>(keys); break;
            default: bitonic_sort_tiled<16>(keys); break;
            }
        } else {
            for (uint32_t i = tid; i < n; i += kTokThreads) {
                const uint32_t key = rank_key<GeneT, ValT>(ra, __ldg(gene + s + i), __ldg(value + s + i));
                keys[i] = ((unsigned long long)(~key) << 32) | (unsigned long long)i;
            }
            __syncthreads();
            bitonic_sort_asc_u64(keys, n);
        }
        rank_emit<GeneT, ValT>(ra, keys, j, c, s, n, s_ws, s_bcast);
    }
}

// K6b (extension): the rank-value kernel without a sorting network.  The composites of a cell are DISTRIBUTED over buckets by a
// monotone function of the key, the bucket counts are prefix-summed, every composite goes to its bucket's range, and the order
// inside a bucket comes from every row counting the smaller composites of its own bucket (work spread by rows: a cluster of keys
// keeps all threads busy).  The bucket function has two levels so that it adapts to the cell: 256 coarse buckets cut the range
// [min, max] of the order-preserving key bits evenly; a coarse bucket that received c rows is cut evenly again into the power of
// two >= c / 2 fine buckets.  Keys spread like the bits of a log-normal and keys clustered around a few values (a normaliser that
// hardly varies between genes) both end up with one or two rows per fine bucket -- a few passes over ~2,000 rows instead of 78
// compare-exchange stages over 4,096 slots.  E = rows per thread held in registers: the first launch (E = 16) takes the cells of
// up to 4,096 rows, a second one (E = 32, from the work list) those of up to 8,192; what is left -- longer cells, or a fine bucket
// of more than kRbMaxBucket rows (heavy ties: e.g. no normaliser at all) -- is drained by tokenize_rank_kernel (the sorting
// network).  Exact: the bucket index is non-decreasing in the composite, and inside a bucket the full 64-bit composites
// (~key, row) are compared.
// a cell too long for the shared-memory paths: composites into the global scratch, generic bitonic sort by the whole CTA, emission
template <typename GeneT, typename ValT>
__device__ __noinline__ void rank_giant_cell(const RankArgs &ra, uint32_t j, uint32_t c, uint32_t s, uint32_t n, uint32_t *s_ws, uint32_t *s_bcast) {
    const GeneT *__restrict__ gene = static_cast<const GeneT *>(ra.t.gene);
    const ValT *__restrict__ value = static_cast<const ValT *>(ra.t.value);
    unsigned long long *keys = ra.scratch64 + s;
    for (uint32_t i = threadIdx.x; i < n; i += kTokThreads) {
        const uint32_t key = rank_key<GeneT, ValT>(ra, __ldg(gene + s + i), __ldg(value + s + i));
        keys[i] = ((unsigned long long)(~key) << 32) | (unsigned long long)i;
    }
    __syncthreads();
    bitonic_sort_asc_u64(keys, n);
    rank_emit<GeneT, ValT>(ra, keys, j, c, s, n, s_ws, s_bcast);
}

constexpr int kRbCoarse = 256;
constexpr int kRbMaxBucket = 64;
template <int E>
__host__ __device__ constexpr size_t rank_bucket_smem() {
    return (size_t)E * kTokThreads * sizeof(unsigned long long)             // composites
           + ((size_t)E * kTokThreads + 2 * kRbCoarse) * sizeof(uint32_t)   // fine counts / offsets (sum of the fine-bucket numbers <= n + 256)
           + 2 * (size_t)kRbCoarse * sizeof(uint32_t);                      // coarse counts -> first fine bucket, log2 of the fine-bucket number
}
template <typename GeneT, typename ValT, int E>
__global__ void __launch_bounds__(kTokThreads, E == 16 ? 4 : 2) tokenize_rank_bucket_kernel(const RankArgs ra) {
    constexpr uint32_t NK = (uint32_t)E * kTokThreads;                      // rows the shared-memory array holds
    constexpr uint32_t NF = NK + 2u * kRbCoarse;                            // fine counters
    constexpr int FPT = (int)(NF / kTokThreads);                            // fine counters per thread (contiguous), a multiple of 2
    extern __shared__ __align__(16) unsigned long long rb_keys[];
    uint32_t *cnt = reinterpret_cast<uint32_t *>(rb_keys + NK);            // [NF]
    uint32_t *cfirst = cnt + NF;                                            // [256] coarse count, then the first fine bucket of the coarse one
    uint32_t *clog = cfirst + kRbCoarse;                                    // [256] log2 of its number of fine buckets
    __shared__ uint32_t s_ws[32], s_bcast[4], s_item;
    const TokArgs &a = ra.t;
    const int tid = threadIdx.x;
    const GeneT *__restrict__ gene = static_cast<const GeneT *>(a.gene);
    const ValT *__restrict__ value = static_cast<const ValT *>(a.value);
    const uint32_t count = ra.from_list ? a.work_counters[0] : a.n_cells;
    uint32_t *ticket = ra.from_list ? a.work_counters + 3 : a.work_counters + 2;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= count) break;
        const uint32_t j = ra.from_list ? a.work_list[item] : item;
        if (j == 0xffffffffu) continue;
        const uint32_t c = a.perm ? (uint32_t)a.perm[j] : j;
        const uint32_t s = __ldg(a.seg_start + c), e = __ldg(a.seg_start + c + 1), n = e - s;
        // what this launch cannot take stays on / goes to the work list (from_list: the entry simply stays)
        auto leave = [&]() { if (!ra.from_list && tid == 0) a.work_list[atomicAdd(a.work_counters, 1u)] = j; };
        if (!ra.from_list && n > 2u * NK) {
            // a giant cell (beyond the second launch as well): sorted in global scratch by this CTA right here, while the other CTAs
            // go on -- left to the sorting kernel it would be that launch's only work item and the step's tail
            rank_giant_cell<GeneT, ValT>(ra, j, c, s, n, s_ws, s_bcast);
            continue;
        }
        if (n > NK) { leave(); continue; }                                  // (uniform)
        // 1. inverted keys of the rows tid, tid + 256, ... in registers (loads in batches of eight rows: gene and value first, then
        // the gathers of the normaliser, so that eight of each are in flight); range of the cell; counters cleared
        uint32_t ik[E], lo = 0xffffffffu, hi = 0u;
#pragma unroll
        for (int r0 = 0; r0 < E; r0 += 8) {
            GeneT g8[8];
            ValT v8[8];
            float d8[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t i = (uint32_t)tid + (uint32_t)(r0 + q) * kTokThreads;
                g8[q] = GeneT{}; v8[q] = ValT{};
                if (i < n) { g8[q] = __ldg(gene + s + i); v8[q] = __ldg(value + s + i); }
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t i = (uint32_t)tid + (uint32_t)(r0 + q) * kTokThreads;
                const uint32_t gi = (uint32_t)g8[q];
                d8[q] = (i < n && ra.norm != nullptr && gi < ra.n_norm) ? __ldg(ra.norm + gi) : 1.0f;
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t i = (uint32_t)tid + (uint32_t)(r0 + q) * kTokThreads;
                ik[r0 + q] = 0xffffffffu;
                if (i < n) {
                    ik[r0 + q] = ~(ra.norm == nullptr ? key_of(v8[q]) : key_of(__fdiv_rn((float)v8[q], d8[q])));
                    lo = min(lo, ik[r0 + q]);
                    hi = max(hi, ik[r0 + q]);
                }
            }
        }
        {
            uint2 *c2 = reinterpret_cast<uint2 *>(cnt) + tid * (FPT / 2);
#pragma unroll
            for (int q = 0; q < FPT / 2; q++) c2[q] = make_uint2(0u, 0u);
            cfirst[tid] = 0u;                                               // (kRbCoarse == kTokThreads)
        }
        block_minmax(lo, hi, s_ws);                                         // (two barriers: the counters are cleared for everybody after it)
        const uint32_t range = hi - lo;
        const uint32_t sh1 = range >= (uint32_t)kRbCoarse ? (32u - (uint32_t)__clz(range)) - 8u : 0u;      // (range >> sh1) < 256
        // 2. coarse histogram
#pragma unroll
        for (int r = 0; r < E; r++) {
            const uint32_t i = (uint32_t)tid + (uint32_t)r * kTokThreads;
            if (i < n) atomicAdd(&cfirst[(ik[r] - lo) >> sh1], 1u);
        }
        __syncthreads();
        // 3. fine buckets per coarse bucket: the power of two >= count / 2 (at most the 2^sh1 key values it spans), and their prefix
        {
            const uint32_t cc = cfirst[tid];
            uint32_t lg = cc > 2u ? 32u - (uint32_t)__clz((cc - 1u) >> 1) : 0u;      // ceil(log2(cc / 2)) for cc > 2
            lg = min(lg, sh1);
            uint32_t tot;
            const uint32_t fb = block_excl_scan(1u << lg, s_ws, tot);      // (barriers inside: every coarse count has been read)
            SLAFB_CHECK(fb + (1u << lg) <= NF, 253, fb);
            cfirst[tid] = fb;
            clog[tid] = lg;
        }
        __syncthreads();
        // fine bucket of a key: its coarse bucket's first fine bucket + the leading bits of its offset inside the coarse bucket
        auto fine = [&](uint32_t k32) -> uint32_t {
            const uint32_t d = k32 - lo, cb = d >> sh1;
            return cfirst[cb] + ((d & ((1u << sh1) - 1u)) >> (sh1 - clog[cb]));
        };
        // 4. fine histogram; the value the atomic returns is the row's slot inside its bucket (four 8-bit slots per register; the
        // bucket itself is recomputed where it is needed again: registers are what limits the CTAs per SM here)
        uint32_t slot[E / 4];
#pragma unroll
        for (int q = 0; q < E / 4; q++) slot[q] = 0u;
#pragma unroll
        for (int r = 0; r < E; r++) {
            const uint32_t i = (uint32_t)tid + (uint32_t)r * kTokThreads;
            if (i < n) {
                const uint32_t f = fine(ik[r]);
                SLAFB_CHECK(f < NF, 250, f);
                slot[r >> 2] |= min(atomicAdd(&cnt[f], 1u), 255u) << (8 * (r & 3));
            }
        }
        __syncthreads();
        // 5. exclusive prefix over the fine counts (thread t owns FPT contiguous ones); largest bucket
        uint32_t sum = 0u, mx = 0u;
        {
            const uint2 *c2 = reinterpret_cast<const uint2 *>(cnt) + tid * (FPT / 2);
#pragma unroll
            for (int q = 0; q < FPT / 2; q++) { const uint2 v = c2[q]; sum += v.x + v.y; mx = max(mx, max(v.x, v.y)); }
        }
        uint32_t tot;
        const uint32_t first = block_excl_scan(sum, s_ws, tot);
        if (__syncthreads_or(mx > (uint32_t)kRbMaxBucket)) { leave(); continue; }     // heavy ties: the sorting kernel takes the cell
        {
            uint32_t run = first;
            uint2 *c2 = reinterpret_cast<uint2 *>(cnt) + tid * (FPT / 2);
#pragma unroll
            for (int q = 0; q < FPT / 2; q++) {
                const uint2 v = c2[q];
                uint2 o;
                o.x = run; run += v.x;
                o.y = run; run += v.y;
                c2[q] = o;
            }
        }
        __syncthreads();
        // 6. every composite into its bucket's range (any order inside the bucket)
#pragma unroll
        for (int r = 0; r < E; r++) {
            const uint32_t i = (uint32_t)tid + (uint32_t)r * kTokThreads;
            if (i < n) {
                const uint32_t pos = cnt[fine(ik[r])] + ((slot[r >> 2] >> (8 * (r & 3))) & 0xffu);
                SLAFB_CHECK(pos < n, 251, pos);
                rb_keys[pos] = ((unsigned long long)ik[r] << 32) | (unsigned long long)i;
            }
        }
        __syncthreads();
        // 7. order inside the buckets: every row counts the composites of its bucket that are smaller than its own -- its final
        // place.  Places first, barrier, then the composites -- still in registers -- are written to their final positions.
        uint32_t place[E / 2];                                              // two 16-bit positions per register
#pragma unroll
        for (int q = 0; q < E / 2; q++) place[q] = 0u;
#pragma unroll
        for (int r = 0; r < E; r++) {
            const uint32_t i = (uint32_t)tid + (uint32_t)r * kTokThreads;
            if (i < n) {
                const uint32_t f = fine(ik[r]);
                const uint32_t o = cnt[f], oe = (f + 1u < NF) ? cnt[f + 1u] : n;
                const unsigned long long mine = ((unsigned long long)ik[r] << 32) | (unsigned long long)i;
                uint32_t rank = 0u;
                for (uint32_t x = o; x < oe; x++) rank += rb_keys[x] < mine ? 1u : 0u;
                place[r >> 1] |= (o + rank) << (16 * (r & 1));
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < E; r++) {
            const uint32_t i = (uint32_t)tid + (uint32_t)r * kTokThreads;
            if (i < n) {
                const uint32_t pos = (place[r >> 1] >> (16 * (r & 1))) & 0xffffu;
                SLAFB_CHECK(pos < n, 252, pos);
                rb_keys[pos] = ((unsigned long long)ik[r] << 32) | (unsigned long long)i;
            }
        }
        __syncthreads();
        rank_emit<GeneT, ValT>(ra, rb_keys, j, c, s, n, s_ws, s_bcast);
        if (ra.from_list && tid == 0) a.work_list[item] = 0xffffffffu;      // done: the sorting kernel skips the entry
    }
}

// K7 (extension): per-gene histogram of uint16 values (exact nonzero medians), [n_genes x n_bins] uint32;
// values >= n_bins - 1 land in the last bin.  Integer counts: any reduction order gives the same table.
template <typename GeneT>
__global__ void __launch_bounds__(256) gene_hist_kernel(const GeneT *__restrict__ gene, const uint16_t *__restrict__ value,
                                                        uint32_t n_rows, uint32_t n_genes, uint32_t n_bins, uint32_t *hist) {
    const uint32_t stride = gridDim.x * blockDim.x * kGroup;
    for (uint32_t r0 = (blockIdx.x * blockDim.x + threadIdx.x) * kGroup; r0 < n_rows; r0 += stride) {
        GeneT g[8];
        uint16_t v[8];
        load8(gene, r0, n_rows, g);
        load8(value, r0, n_rows, v);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t gi = (uint32_t)g[i];
            if (r0 + i < n_rows && gi < n_genes && v[i] != 0)
                atomicAdd(hist + (size_t)gi * n_bins + min((uint32_t)v[i], n_bins - 1u), 1u);
        }
    }
}

// median of the nonzero values of every gene from its histogram row (numpy.median: mean of the two middle
// values for an even count); one warp per gene.  exact[g] = 0 when a middle value fell into the overflow bin.
__global__ void __launch_bounds__(256) gene_median_kernel(const uint32_t *__restrict__ hist, uint32_t n_genes, uint32_t n_bins,
                                                          float *__restrict__ median, uint8_t *__restrict__ exact) {
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_genes) return;
    const uint32_t *row = hist + (size_t)g * n_bins;
    unsigned long long total = 0;
    for (uint32_t b = lane; b < n_bins; b += 32) total += row[b];
#pragma unroll
    for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
    if (total == 0) { if (lane == 0) { median[g] = 0.0f; exact[g] = 1; } return; }
    const unsigned long long r_lo = (total - 1) >> 1, r_hi = total >> 1;      // 0-based ranks of the middle values
    uint32_t v_lo = 0xffffffffu, v_hi = 0xffffffffu;
    unsigned long long run = 0;
    for (uint32_t b0 = 0; b0 < n_bins && v_hi == 0xffffffffu; b0 += 32) {
        const uint32_t b = b0 + lane;
        const unsigned long long cnt = b < n_bins ? row[b] : 0ull;
        unsigned long long inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
            if ((int)lane >= d) inc += t;
        }
        const unsigned long long ex = run + inc - cnt;
        const bool has_lo = cnt && r_lo >= ex && r_lo < ex + cnt, has_hi = cnt && r_hi >= ex && r_hi < ex + cnt;
        const uint32_t m_lo = __ballot_sync(0xffffffffu, has_lo), m_hi = __ballot_sync(0xffffffffu, has_hi);
        if (m_lo && v_lo == 0xffffffffu) v_lo = b0 + (__ffs(m_lo) - 1);
        if (m_hi) v_hi = b0 + (__ffs(m_hi) - 1);
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) {
        median[g] = 0.5f * ((float)v_lo + (float)v_hi);
        exact[g] = (v_hi < n_bins - 1u) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------
// facade helpers
// ------------------------------------------------------------------------------------------
// exclusive scan of kept counts (one CTA; C is at most a few hundred thousand)
__global__ void __launch_bounds__(1024) scan_counts_kernel(const uint32_t *cnt, long long *offsets, uint32_t n,
                                                           long long *total_out) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0ull;
    __syncthreads();
    for (uint32_t b = 0; b < n; b += 1024) {
        const uint32_t i = b + tid;
        unsigned long long v = (i < n) ? cnt[i] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = wsum[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += t;
            }
            wsum[lane] = winc - w;
        }
        __syncthreads();
        const unsigned long long ex = carry + wsum[warp] + inc - v;
        if (i < n) offsets[i] = (long long)ex;
        __syncthreads();
        if (tid == 1023) carry = ex + v;
        __syncthreads();
    }
    if (tid == 0) { offsets[n] = (long long)carry; *total_out = (long long)carry; }
}

__global__ void cell_ids_kernel(const uint32_t *seg_cell, const int32_t *perm, uint32_t n, int cell_signed,
                                long long *cell_ids) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t c = perm ? (uint32_t)perm[j] : j;
    const uint32_t cid = seg_cell[c];
    cell_ids[j] = cell_signed ? (long long)(int32_t)cid : (long long)cid;
}

// The permutation reaches the device through the SMs, not the copy engine: 4 bytes per cell read from mapped pinned host
// memory and written to HBM.  (A cudaMemcpyAsync would queue behind whole batches of column copies on the H2D engine.)
__global__ void perm_upload_kernel(const int32_t *__restrict__ host_perm, int32_t *__restrict__ perm, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = host_perm[i];
}

// duplicate detection over the segment table's cell ids (a cell split into two separated row runs
// would silently become two output rows): open-addressing hash set, 64-bit slots (1<<32 | id)
__global__ void dup_check_kernel(const uint32_t *seg_cell, uint32_t n, unsigned long long *table, uint32_t mask, uint32_t *flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = seg_cell[i];
    const unsigned long long key = (1ull << 32) | id;
    uint32_t h = (id * 2654435761u) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(table + h, 0ull, key);
        if (old == 0ull) return;
        if (old == key) { *flag = 1u; return; }
        h = (h + 1u) & mask;
    }
}

// K8: raw mode on the device (RawPrefetchBatch, reference slaf/ml/datasets.py:288-296,958-1008): the batch as a
// CSR matrix whose rows are the cells in shuffled order.  Lengths in output order -> scan -> segment copies.
__global__ void csr_lengths_kernel(const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ seg_cell,
                                   const int32_t *__restrict__ perm, uint32_t n, int cell_signed, uint32_t *__restrict__ len,
                                   long long *__restrict__ cell_ids) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t c = perm ? (uint32_t)perm[j] : j;
    len[j] = seg_start[c + 1] - seg_start[c];
    const uint32_t cid = seg_cell[c];
    cell_ids[j] = cell_signed ? (long long)(int32_t)cid : (long long)cid;
}

// one warp per output row copies its segment (genes widened to int32 column indices, values kept in their dtype)
template <typename GeneT, typename ValT>
__global__ void __launch_bounds__(256) csr_gather_kernel(const GeneT *__restrict__ gene, const ValT *__restrict__ value,
                                                         const uint32_t *__restrict__ seg_start, const int32_t *__restrict__ perm,
                                                         uint32_t n, const long long *__restrict__ crow, int32_t *__restrict__ col,
                                                         ValT *__restrict__ val) {
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
    for (uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += warps) {
        const uint32_t c = perm ? (uint32_t)perm[j] : j;
        const uint32_t s = seg_start[c], m = seg_start[c + 1] - s;
        const long long o = crow[j];
        for (uint32_t i = lane; i < m; i += 32) {
            col[o + i] = (int32_t)__ldg(gene + s + i);
            val[o + i] = __ldg(value + s + i);
        }
    }
}

// K4: SLAFTokenizer.tokenize on flattened pre-windowed lists: one CTA per list
struct ListArgs {
    const long long *offsets;
    const long long *genes;
    const double *exprs;
    int exprs_are_int;
    int scgpt;
    uint32_t n_lists;
    uint32_t max_genes;
    uint32_t L;
    int n_bins;
    long long vocab;
    float bin_size;
    long long *ids;
    uint8_t *mask;
    long long *vals;
};

__global__ void __launch_bounds__(kTokThreads) tokenize_lists_kernel(const ListArgs a) {
    const uint32_t row = blockIdx.x;
    const long long s = a.offsets[row], e = a.offsets[row + 1];
    const unsigned long long n = (unsigned long long)(e - s);
    const uint32_t L = a.L;
    const uint32_t limit = a.scgpt ? a.max_genes : (L - 1u);
    const uint32_t n_emit = (uint32_t)(n < (unsigned long long)limit ? n : (unsigned long long)limit);
    const size_t row_base = (size_t)row * L;
    for (uint32_t i = threadIdx.x; i < n_emit; i += kTokThreads) {
        a.ids[row_base + 1 + i] = a.genes[s + i] + 4;
        if (a.scgpt) {
            const double x = a.exprs[s + i];
            a.vals[row_base + 1 + i] = a.exprs_are_int ? (long long)x + a.vocab
                                                       : expr_bin_token((float)x, a.bin_size, a.n_bins, a.vocab);
        }
    }
    const uint32_t sep_pos = 1u + n_emit, t_len = min(n_emit + 2u, L);
    if (threadIdx.x == 0) {
        a.ids[row_base] = 1;
        if (sep_pos < L) a.ids[row_base + sep_pos] = 2;
        if (a.scgpt) a.vals[row_base] = 0;
    }
    fill_zero_i64(a.ids, row_base + t_len, row_base + L);
    if (a.scgpt) fill_zero_i64(a.vals, row_base + sep_pos, row_base + L);
    fill_mask(a.mask, row_base, L, t_len);
}

// K5: per-gene nonzero sum / count (integer sums: identical for any rank count / order)
template <typename GeneT, typename ValT>
__global__ void __launch_bounds__(256) gene_stats_kernel(const GeneT *__restrict__ gene, const ValT *__restrict__ value,
                                                         uint32_t n_rows, uint32_t n_genes,
                                                         unsigned long long *sum, unsigned long long *cnt) {
    const uint32_t stride = gridDim.x * blockDim.x * kGroup;
    for (uint32_t r0 = (blockIdx.x * blockDim.x + threadIdx.x) * kGroup; r0 < n_rows; r0 += stride) {
        GeneT g[8];
        ValT v[8];
        load8(gene, r0, n_rows, g);
        load8(value, r0, n_rows, v);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (r0 + i < n_rows) {
                const uint32_t gi = (uint32_t)g[i];
                if (gi < n_genes) {
                    if constexpr (sizeof(ValT) == 2) {
                        if (v[i] != 0) { atomicAdd(sum + gi, (unsigned long long)v[i]); atomicAdd(cnt + gi, 1ull); }
                    } else {
                        // float32 values: fixed-point 2^-20 accumulation keeps the reduction order-independent
                        if (v[i] != 0.0f) {
                            atomicAdd(sum + gi, (unsigned long long)llrintf(v[i] * 1048576.0f));
                            atomicAdd(cnt + gi, 1ull);
                        }
                    }
                }
            }
        }
    }
}

// K5, uint16 values: ONE 64-bit reduction per nonzero row instead of two -- the row's value in the upper 40 bits, a count of one in
// the lower 24 -- into a per-context scratch vector, unpacked into the caller's sum / count vectors by a second, tiny kernel.
// The host keeps a launch below 2^24 rows, so neither field can overflow (sum <= 65535 * (2^24 - 1) < 2^40).
constexpr uint32_t kStatsRowsPerLaunch = (1u << 24) - 8u;       // multiple of 8: chunks keep the 16-byte alignment of the columns
template <typename GeneT>
__global__ void __launch_bounds__(256) gene_stats_packed_kernel(const GeneT *__restrict__ gene, const uint16_t *__restrict__ value,
                                                                uint32_t n_rows, uint32_t n_genes, unsigned long long *packed) {
    const uint32_t stride = gridDim.x * blockDim.x * kGroup;
    for (uint32_t r0 = (blockIdx.x * blockDim.x + threadIdx.x) * kGroup; r0 < n_rows; r0 += stride) {
        GeneT g[8];
        uint16_t v[8];
        load8(gene, r0, n_rows, g);
        load8(value, r0, n_rows, v);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t gi = (uint32_t)g[i];
            if (r0 + i < n_rows && gi < n_genes && v[i] != 0) atomicAdd(packed + gi, ((unsigned long long)v[i] << 24) | 1ull);
        }
    }
}
__global__ void gene_stats_unpack_kernel(unsigned long long *packed, uint32_t n_genes, unsigned long long *sum, unsigned long long *cnt) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_genes) return;
    const unsigned long long p = packed[g];
    if (p) {
        sum[g] += p >> 24;
        cnt[g] += p & 0xffffffull;
        packed[g] = 0ull;
    }
}

// Packed route, cells the tokenizing kernel deferred (its work list: blocks larger than the raw stage, cells that need the exact
// k-th distinct value): their blocks are decoded into plain gene / value columns at the cells' row offsets, where the general
// kernel reads them.  One CTA per deferred cell, the same decode as in the tokenizing kernel but straight from global memory.
template <typename GeneT>
__global__ void __launch_bounds__(kTokThreads) unpack_deferred_kernel(const uint8_t *__restrict__ packed, const uint32_t *__restrict__ blk_off,
                                                                      const uint32_t *__restrict__ seg_start, const int32_t *__restrict__ perm,
                                                                      const uint32_t *__restrict__ work_list, const uint32_t *__restrict__ work_count,
                                                                      GeneT *__restrict__ gene, uint16_t *__restrict__ value) {
    __shared__ uint32_t s_ws[kTokThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t count = *work_count;
    for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
        const uint32_t row = work_list[it];
        const uint32_t c = perm ? (uint32_t)perm[row] : row;
        const uint32_t s = seg_start[c], n = seg_start[c + 1] - s;
        const uint8_t *rb = packed + ((size_t)blk_off[c] << 4);
        const uint4 hdr = __ldg(reinterpret_cast<const uint4 *>(rb));
        const uint32_t np = (n + 15u) & ~15u, base_gene = hdr.y, n_eg = hdr.z, n_ev = hdr.w;
        const uint8_t *dl = rb + 16u, *vl = dl + np;
        const uint32_t *exg = reinterpret_cast<const uint32_t *>(vl + np), *exv = exg + 2u * n_eg;
        uint32_t carry = 0u;
        for (uint32_t base = 0u; base < n; base += (uint32_t)kTokThreads * 8u) {
            const uint32_t i0 = base + (uint32_t)tid * 8u;
            uint32_t dd[8], vv[8], run = 0u;
            if (i0 < n) {
                const uint2 dq = __ldg(reinterpret_cast<const uint2 *>(dl + i0)), vq = __ldg(reinterpret_cast<const uint2 *>(vl + i0));
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    dd[i] = ((i < 4 ? dq.x : dq.y) >> ((i & 3) * 8)) & 0xffu;
                    vv[i] = ((i < 4 ? vq.x : vq.y) >> ((i & 3) * 8)) & 0xffu;
                    if (vv[i] == 255u && i0 + i < n) vv[i] = packed_exception(exv, n_ev, i0 + i);
                    if (dd[i] == 255u && i0 + i < n) dd[i] = packed_exception(exg, n_eg, i0 + i);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) { run += (i0 + i < n) ? dd[i] + 1u : 0u; dd[i] = run; }
            }
            const uint32_t inc = warp_incl_scan(run);
            __syncthreads();
            if (lane == 31) s_ws[warp] = inc;
            __syncthreads();
            uint32_t wbase = 0u, tot = 0u;
#pragma unroll
            for (int w2 = 0; w2 < kTokThreads / 32; w2++) {
                const uint32_t t = s_ws[w2];
                if (w2 < warp) wbase += t;
                tot += t;
            }
            if (i0 < n) {
                const uint32_t g0 = base_gene + carry + wbase + inc - run - 1u;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (i0 + i < n) {
                        gene[s + i0 + i] = (GeneT)(g0 + dd[i]);
                        value[s + i0 + i] = (uint16_t)vv[i];
                    }
                }
            }
            carry += tot;
        }
        __syncthreads();
    }
}

}  // namespace slafb
