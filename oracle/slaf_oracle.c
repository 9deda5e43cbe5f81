/*
 * slaf_oracle.c -- CPU ORACLE for the SLAF ML-dataloader hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is a plain-C restatement of the
 * reference's algorithm (reference = slaf-project/slaf, package slafdb 0.5.2).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  The product path (slaf_b200/) never
 * imports, links or executes anything in oracle/.
 *
 * Parity status
 *   tokenizer half  : PINNED  -- tests/golden/tokenizer_golden.npz was produced by
 *                     executing the reference's own slaf/ml/tokenizers.py in the
 *                     build container (tests/golden/make_golden.py).
 *   window half     : pinned only by the known-answer vectors in the reference's
 *                     tests (tests/golden/window_kats.json cites them); polars is
 *                     not installable here, so a full-frame window output of the
 *                     reference could not be generated ("parity unpinned" beyond
 *                     those KATs).
 *   third-party math: the window arithmetic lives in polars 1.39.3 (Rust std ->
 *                     system libm).  This oracle calls glibc log1p()/log1pf(), the
 *                     same functions Rust's f64::ln_1p / f32::ln_1p resolve to on
 *                     x86_64-unknown-linux-gnu.
 *
 * Reference lines followed (paths relative to the reference checkout):
 *   grouping / order of cells : slaf/ml/samplers.py:93 (partition_by, first appearance)
 *                               slaf/ml/aggregators.py:111,190,208 (maintain_order=True)
 *   dense-rank filter         : slaf/ml/aggregators.py:88-95,120-127,198-207,240-247
 *   percentile filter         : slaf/ml/aggregators.py:167-189
 *   log1p binning             : slaf/ml/aggregators.py:96-110
 *   Geneformer tokens         : slaf/ml/tokenizers.py:666-722
 *   scGPT tokens              : slaf/ml/tokenizers.py:407-479, 502-508, 521-543
 *   cell id order             : slaf/ml/datasets.py:1087-1098
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* dtype codes shared with include/slaf_b200.h (kept numerically identical on purpose) */
enum { DT_U16 = 0, DT_I32 = 1, DT_U32 = 2, DT_F32 = 3 };
enum { MODE_GENEFORMER = 0, MODE_GENEFORMER_PCT = 1, MODE_SCGPT_RAW = 2, MODE_SCGPT_BINNED = 3 };

typedef struct {
    int32_t mode;
    int32_t max_items;     /* window max_items (= tokenizer.max_genes in the dataloader) */
    int32_t max_genes;     /* tokenizer.max_genes */
    int32_t n_bins;        /* n_expression_bins */
    int64_t vocab_size;    /* expr_bin_start */
    double  min_percentile;
} oracle_params;

int slaf_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline int64_t load_int(const void *p, int dt, int64_t i) {
    switch (dt) {
    case DT_U16: return ((const uint16_t *)p)[i];
    case DT_I32: return ((const int32_t *)p)[i];
    case DT_U32: return ((const uint32_t *)p)[i];
    default: return 0;
    }
}

/* value as double: exact, order preserving for u16 / i32 / u32 / f32 */
static inline double load_val(const void *p, int dt, int64_t i) {
    if (dt == DT_F32) return (double)((const float *)p)[i];
    return (double)load_int(p, dt, i);
}

/* ------------------------------------------------------------------------
 * Grouping: cells in order of first appearance, rows of a cell in row order
 * (polars partition_by / group_by(maintain_order=True)).  Handles cells whose
 * rows are NOT contiguous (slaf/ml/datasets.py:880-888 can produce that).
 *   order[R]      : row indices, grouped
 *   seg_start[C+1]: offsets into order
 *   seg_cell[C]   : the cell id of each group
 * returns C, or -1 on allocation failure / capacity overflow.
 * ---------------------------------------------------------------------- */
static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

int64_t slaf_oracle_group(const void *cell, int cell_dt, int64_t R, int64_t *order,
                          int64_t *seg_start, int64_t *seg_cell, int64_t cap_cells) {
    if (R == 0) { seg_start[0] = 0; return 0; }
    /* open addressing hash: cell id -> group index */
    int64_t hcap = 16;
    while (hcap < 4 * (cap_cells < R ? cap_cells : R) + 16) hcap <<= 1;
    int64_t *hkey = (int64_t *)malloc(sizeof(int64_t) * hcap);
    int64_t *hval = (int64_t *)malloc(sizeof(int64_t) * hcap);
    int64_t *grp = (int64_t *)malloc(sizeof(int64_t) * R);
    int64_t *cnt = (int64_t *)calloc((size_t)cap_cells + 1, sizeof(int64_t));
    if (!hkey || !hval || !grp || !cnt) { free(hkey); free(hval); free(grp); free(cnt); return -1; }
    for (int64_t i = 0; i < hcap; i++) hval[i] = -1;
    int64_t C = 0;
    int64_t prev_id = 0, prev_g = -1;
    for (int64_t i = 0; i < R; i++) {
        int64_t id = load_int(cell, cell_dt, i);
        int64_t g;
        if (prev_g >= 0 && id == prev_id) {
            g = prev_g;
        } else {
            uint64_t h = mix64((uint64_t)id) & (uint64_t)(hcap - 1);
            while (hval[h] >= 0 && hkey[h] != id) h = (h + 1) & (uint64_t)(hcap - 1);
            if (hval[h] < 0) {
                if (C >= cap_cells) { free(hkey); free(hval); free(grp); free(cnt); return -1; }
                hkey[h] = id; hval[h] = C; seg_cell[C] = id; C++;
            }
            g = hval[h];
            prev_id = id; prev_g = g;
        }
        grp[i] = g;
        cnt[g]++;
    }
    seg_start[0] = 0;
    for (int64_t g = 0; g < C; g++) seg_start[g + 1] = seg_start[g] + cnt[g];
    for (int64_t g = 0; g < C; g++) cnt[g] = seg_start[g];
    for (int64_t i = 0; i < R; i++) order[cnt[grp[i]]++] = i;
    free(hkey); free(hval); free(grp); free(cnt);
    return C;
}

/* ------------------------------------------------------------------------
 * Window for ONE cell (rows given through the index list idx[0..n)).
 *   keep[i]   : 1 if row i survives  rank(dense, desc) <= max_items  [and the
 *               percentile test]                     aggregators.py:199-207 / 181-189
 *   bins[i]   : (optional, may be NULL) the expr_bin column of the binned scGPT
 *               window, as double (f64 storage for integer values, f32 arithmetic
 *               widened exactly to double for f32 values)      aggregators.py:96-110
 * returns number of kept rows.
 * ---------------------------------------------------------------------- */
static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

int64_t slaf_oracle_window_cell(const void *value, int val_dt, const int64_t *idx, int64_t n,
                                int32_t max_items, int use_pct, double min_percentile,
                                int binned, int32_t n_bins, uint8_t *keep, double *bins) {
    if (n == 0) return 0;
    double *v = (double *)malloc(sizeof(double) * (size_t)n);
    double *d = (double *)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; i++) { v[i] = load_val(value, val_dt, idx[i]); d[i] = v[i]; }
    qsort(d, (size_t)n, sizeof(double), cmp_double);
    int64_t D = 0;                       /* distinct values, ascending in d[0..D) */
    for (int64_t i = 0; i < n; i++) if (i == 0 || d[i] != d[D - 1]) d[D++] = d[i];
    int64_t kept = 0;
    for (int64_t i = 0; i < n; i++) {
        /* ascending dense rank pr in 1..D ; descending dense rank r = D - pr + 1 */
        int64_t lo = 0, hi = D - 1;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (d[mid] < v[i]) lo = mid + 1; else hi = mid; }
        int64_t pr = lo + 1;
        int64_t r = D - pr + 1;
        int k = (r <= (int64_t)max_items);
        if (use_pct) {
            /* percentile_rank >= min_percentile * max(percentile_rank) / 100   (f64) */
            double t = min_percentile * (double)D / 100.0;
            k = k && ((double)pr >= t);
        }
        keep[i] = (uint8_t)k;
        kept += k;
    }
    if (binned && bins) {
        if (val_dt == DT_F32) {
            /* polars keeps Float32: log1pf, f32 mul/div/floor/clip */
            float m = -INFINITY;
            for (int64_t i = 0; i < n; i++) if (keep[i]) { float lv = log1pf((float)v[i]); if (lv > m) m = lv; }
            for (int64_t i = 0; i < n; i++) {
                if (!keep[i]) { bins[i] = 0.0; continue; }
                float lv = log1pf((float)v[i]);
                float b = 0.0f;
                if (lv > 0.0f) {
                    volatile float t = lv * (float)n_bins;
                    volatile float q = t / m;
                    b = floorf(q);
                    if (b < 0.0f) b = 0.0f;
                    if (b > (float)(n_bins - 1)) b = (float)(n_bins - 1);
                }
                bins[i] = (double)b;
            }
        } else {
            double m = -INFINITY;
            for (int64_t i = 0; i < n; i++) if (keep[i]) { double lv = log1p(v[i]); if (lv > m) m = lv; }
            for (int64_t i = 0; i < n; i++) {
                if (!keep[i]) { bins[i] = 0.0; continue; }
                double lv = log1p(v[i]);
                double b = 0.0;
                if (lv > 0.0) {
                    volatile double t = lv * (double)n_bins;
                    volatile double q = t / m;
                    b = floor(q);
                    if (b < 0.0) b = 0.0;
                    if (b > (double)(n_bins - 1)) b = (double)(n_bins - 1);
                }
                bins[i] = b;
            }
        }
    }
    free(v); free(d);
    return kept;
}

/* numpy `.astype(int)` of a float32 on x86-64 (cvttss2si): out-of-range / NaN give INT64_MIN */
static inline int64_t f32_to_i64_x86(float q) {
    if (!(q >= -9223372036854775808.0f && q < 9223372036854775808.0f)) return INT64_MIN;
    return (int64_t)q;
}

/* slaf/ml/tokenizers.py:521-543 with expression_values float32 and expr_bin_size a Python
 * float (weak scalar -> float32 under NEP 50) */
static inline int64_t expr_to_bin_token(float x, float bin_size_f32, int32_t n_bins, int64_t vocab) {
    volatile float q = x / bin_size_f32;
    int64_t b = f32_to_i64_x86(q);
    if (b < 0) b = 0;
    if (b > (int64_t)n_bins - 1) b = (int64_t)n_bins - 1;
    return (x > 0.0f) ? vocab + b : 0;
}

/* ------------------------------------------------------------------------
 * Tokenize ONE cell into row `out_row` of the output tensors.
 * gene ids g[0..m), expression items x[0..m) (already windowed, in row order).
 * ---------------------------------------------------------------------- */
static void tokenize_geneformer_row(const int64_t *g, int64_t m, int32_t max_genes, int64_t *ids,
                                    uint8_t *mask) {
    int64_t L = max_genes;
    for (int64_t j = 0; j < L; j++) ids[j] = 0;
    /* tokens = [CLS] ++ (g+4) ++ [SEP]; tokens[:max_genes]; right pad 0   tokenizers.py:684-716 */
    int64_t t = 0;
    if (t < L) ids[t] = 1;
    t++;
    for (int64_t i = 0; i < m; i++, t++) if (t < L) ids[t] = g[i] + 4;
    if (t < L) ids[t] = 2;
    for (int64_t j = 0; j < L; j++) mask[j] = (uint8_t)(ids[j] != 0);
}

static void tokenize_scgpt_row(const int64_t *g, const double *x, int x_is_int, int64_t m,
                               int32_t max_genes, int32_t n_bins, int64_t vocab, int64_t *ids,
                               int64_t *vals, uint8_t *mask) {
    int64_t L = (int64_t)max_genes + 2;
    for (int64_t j = 0; j < L; j++) { ids[j] = 0; vals[j] = 0; }
    int64_t n = m < max_genes ? m : max_genes;          /* tokenizers.py:436 */
    float bin_size = (float)(1.0 / (double)n_bins);     /* tokenizers.py:508, cast at :531 */
    ids[0] = 1;
    for (int64_t i = 0; i < n; i++) {
        ids[1 + i] = g[i] + 4;
        if (x_is_int) vals[1 + i] = (int64_t)x[i] + vocab;                           /* :441-444 */
        else vals[1 + i] = expr_to_bin_token((float)x[i], bin_size, n_bins, vocab);  /* :446-448 */
    }
    ids[1 + n] = 2;
    for (int64_t j = 0; j < L; j++) mask[j] = (uint8_t)(ids[j] != 0);
}

/* ------------------------------------------------------------------------
 * Whole prefetch batch: COO rows -> (input_ids, attention_mask, values, cell_ids).
 *   perm : permutation of the first-appearance cell order (random.shuffle of the
 *          chunk list, samplers.py:93-96), or NULL for identity.
 *   ids  : int64 [C * L] ; mask : uint8 [C * L] ; vals : int64 [C * L] or NULL (GF)
 *   cell_ids : int64 [C]
 *   L = max_genes (Geneformer) or max_genes + 2 (scGPT).
 * returns C (number of cells) or a negative error.
 * ---------------------------------------------------------------------- */
int64_t slaf_oracle_tokenize(const void *cell, int cell_dt, const void *gene, int gene_dt,
                             const void *value, int val_dt, int64_t R, const int64_t *perm,
                             const oracle_params *p, int64_t cap_cells, int64_t *ids, uint8_t *mask,
                             int64_t *vals, int64_t *cell_ids) {
    int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(R > 0 ? R : 1));
    int64_t *seg_start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cap_cells + 1));
    int64_t *seg_cell = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cap_cells > 0 ? cap_cells : 1));
    if (!order || !seg_start || !seg_cell) { free(order); free(seg_start); free(seg_cell); return -2; }
    int64_t C = slaf_oracle_group(cell, cell_dt, R, order, seg_start, seg_cell, cap_cells);
    if (C < 0) { free(order); free(seg_start); free(seg_cell); return -1; }
    const int scgpt = (p->mode == MODE_SCGPT_RAW || p->mode == MODE_SCGPT_BINNED);
    const int64_t L = scgpt ? (int64_t)p->max_genes + 2 : (int64_t)p->max_genes;

#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t j = 0; j < C; j++) {
        int64_t c = perm ? perm[j] : j;
        int64_t s = seg_start[c], n = seg_start[c + 1] - s;
        uint8_t *keep = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
        double *bins = (p->mode == MODE_SCGPT_BINNED) ? (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1)) : NULL;
        int64_t kept = slaf_oracle_window_cell(value, val_dt, order + s, n, p->max_items,
                                               p->mode == MODE_GENEFORMER_PCT, p->min_percentile,
                                               p->mode == MODE_SCGPT_BINNED, p->n_bins, keep, bins);
        int64_t *g = (int64_t *)malloc(sizeof(int64_t) * (size_t)(kept > 0 ? kept : 1));
        double *x = scgpt ? (double *)malloc(sizeof(double) * (size_t)(kept > 0 ? kept : 1)) : NULL;
        int64_t m = 0;
        for (int64_t i = 0; i < n; i++) {
            if (!keep[i]) continue;
            g[m] = load_int(gene, gene_dt, order[s + i]);
            if (scgpt) x[m] = (p->mode == MODE_SCGPT_BINNED) ? bins[i] : load_val(value, val_dt, order[s + i]);
            m++;
        }
        cell_ids[j] = seg_cell[c];
        if (scgpt) {
            /* integer branch only for the raw window on integer storage (Python ints from
             * polars .to_list()); binned window output is always float   tokenizers.py:441 */
            int x_is_int = (p->mode == MODE_SCGPT_RAW && val_dt != DT_F32);
            tokenize_scgpt_row(g, x, x_is_int, m, p->max_genes, p->n_bins, p->vocab_size,
                               ids + j * L, vals + j * L, mask + j * L);
        } else {
            tokenize_geneformer_row(g, m, p->max_genes, ids + j * L, mask + j * L);
        }
        free(keep); free(bins); free(g); free(x);
    }
    free(order); free(seg_start); free(seg_cell);
    return C;
}

/* ------------------------------------------------------------------------
 * Window facade output (list columns) for a whole frame, flattened:
 *   out_offsets[C+1], out_gene[sum kept] (int64), out_expr[sum kept] (double; raw value or bin)
 *   cell_ids[C]
 * Capacity of out_gene/out_expr must be >= R.  Returns C.
 * ---------------------------------------------------------------------- */
int64_t slaf_oracle_window(const void *cell, int cell_dt, const void *gene, int gene_dt,
                           const void *value, int val_dt, int64_t R, const int64_t *perm,
                           const oracle_params *p, int64_t cap_cells, int64_t *out_offsets,
                           int64_t *out_gene, double *out_expr, int64_t *cell_ids) {
    int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(R > 0 ? R : 1));
    int64_t *seg_start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cap_cells + 1));
    int64_t *seg_cell = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cap_cells > 0 ? cap_cells : 1));
    if (!order || !seg_start || !seg_cell) { free(order); free(seg_start); free(seg_cell); return -2; }
    int64_t C = slaf_oracle_group(cell, cell_dt, R, order, seg_start, seg_cell, cap_cells);
    if (C < 0) { free(order); free(seg_start); free(seg_cell); return -1; }
    int64_t w = 0;
    out_offsets[0] = 0;
    for (int64_t j = 0; j < C; j++) {
        int64_t c = perm ? perm[j] : j;
        int64_t s = seg_start[c], n = seg_start[c + 1] - s;
        uint8_t *keep = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
        double *bins = (p->mode == MODE_SCGPT_BINNED) ? (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1)) : NULL;
        slaf_oracle_window_cell(value, val_dt, order + s, n, p->max_items,
                                p->mode == MODE_GENEFORMER_PCT, p->min_percentile,
                                p->mode == MODE_SCGPT_BINNED, p->n_bins, keep, bins);
        for (int64_t i = 0; i < n; i++) {
            if (!keep[i]) continue;
            out_gene[w] = load_int(gene, gene_dt, order[s + i]);
            out_expr[w] = bins ? bins[i] : load_val(value, val_dt, order[s + i]);
            w++;
        }
        out_offsets[j + 1] = w;
        cell_ids[j] = seg_cell[c];
        free(keep); free(bins);
    }
    free(order); free(seg_start); free(seg_cell);
    return C;
}

/* ------------------------------------------------------------------------
 * Tokenizer facade on already-windowed flattened lists (tokenizers.py tokenize()).
 *   offsets[B+1], g[], x[] ; x_is_int selects the isinstance(int) branch.
 * ---------------------------------------------------------------------- */
int64_t slaf_oracle_tokenize_lists(int scgpt, const int64_t *offsets, const int64_t *g,
                                   const double *x, int x_is_int, int64_t B, int32_t max_genes,
                                   int32_t n_bins, int64_t vocab, int64_t *ids, uint8_t *mask,
                                   int64_t *vals) {
    const int64_t L = scgpt ? (int64_t)max_genes + 2 : (int64_t)max_genes;
    for (int64_t j = 0; j < B; j++) {
        int64_t s = offsets[j], m = offsets[j + 1] - s;
        if (scgpt) tokenize_scgpt_row(g + s, x + s, x_is_int, m, max_genes, n_bins, vocab, ids + j * L, vals + j * L, mask + j * L);
        else tokenize_geneformer_row(g + s, m, max_genes, ids + j * L, mask + j * L);
    }
    return B;
}

/* per-gene nonzero sum / count (extension, not in the reference) */
void slaf_oracle_gene_stats(const void *gene, int gene_dt, const void *value, int val_dt, int64_t R,
                            int64_t n_genes, double *sum, int64_t *cnt) {
    for (int64_t i = 0; i < R; i++) {
        int64_t gidx = load_int(gene, gene_dt, i);
        if (gidx < 0 || gidx >= n_genes) continue;
        double v = load_val(value, val_dt, i);
        if (v != 0.0) { sum[gidx] += v; cnt[gidx] += 1; }
    }
}
