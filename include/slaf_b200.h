/*
 * slaf_b200.h -- C ABI of libslaf_b200.so: the B200 (sm_100a) implementation of SLAF's
 * ML-dataloader hot path  (COO expression rows -> per-cell token tensors).
 *
 * The reference (slaf-project/slaf, package slafdb 0.5.2) is pure Python and has no FFI; the
 * path it computes with polars + numpy is replaced by the entry points below.  Each entry
 * point cites the reference interface it stands in for (paths relative to the reference
 * checkout).  INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 (SLAFB_OK) or a negative slafb_status; slafb_last_error()
 *     gives the text for the calling context.  Nothing aborts, nothing throws.
 *   - the caller owns every pointer it passes.  Output tensors are DEVICE memory allocated by
 *     the caller (PyTorch); the context owns only its scratch, staging buffers, side stream
 *     and events.  After the first batch of a given size no allocation happens per batch.
 *   - a context is single-threaded: one context per (feeder thread, device).
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); results are valid in
 *     stream order after the call that produced them returns.
 *   - there is NO CPU fallback: with no usable sm_100 device slafb_create fails.
 *   - HOST input columns (location = SLAFB_LOC_HOST) are read by the copy engine AFTER submit returns: they must stay valid and
 *     unmodified until the batch's H2D copy has finished -- i.e. until slafb_wait_loaded(slot) returns, or until anything
 *     ordered after the batch's outputs on `stream` has completed (collect returning is NOT enough: it only enqueues work
 *     behind the copy).  Small pageable pieces and the segment tables of submit_segments are copied before submit returns.
 */
#ifndef SLAF_B200_H
#define SLAF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLAFB_ABI_VERSION 3

typedef struct slafb_ctx slafb_ctx;

typedef enum {
    SLAFB_OK = 0,
    SLAFB_ERR_INVALID = -1,      /* bad argument (dtype, NULL, size, alignment)                      */
    SLAFB_ERR_CUDA = -2,         /* a CUDA runtime call failed; text in slafb_last_error             */
    SLAFB_ERR_CAPACITY = -3,     /* more rows / cells than the context or output tensors can hold    */
    SLAFB_ERR_BUSY = -4,         /* pipeline slot still in flight (collect it first)                  */
    SLAFB_ERR_NO_DEVICE = -5,    /* no CUDA device of compute capability 10.x                         */
    SLAFB_ERR_NONCONTIGUOUS = -6 /* a cell id occurs in two separated row runs (host must splice)     */
} slafb_status;

/* Column element types.  Input contract = expression.lance rows
 * (slaf/data/converter.py:2748-2799): cell_integer_id uint32|int32, gene_integer_id
 * uint16|int32|uint32, value uint16|float32.  No nulls, no NaN. */
typedef enum { SLAFB_U16 = 0, SLAFB_I32 = 1, SLAFB_U32 = 2, SLAFB_F32 = 3 } slafb_dtype;
/* Gene ids given as SLAFB_U32 are read as int32: they must be below 2^31 (SLAF's gene_integer_id is a dense index,
 * converter.py:2432-2434; a table with more than 2^31 genes does not exist). */

/* Which window + tokenizer pair runs (slaf/ml/aggregators.py, slaf/ml/tokenizers.py). */
typedef enum {
    SLAFB_GENEFORMER = 0,     /* GeneformerWindow plain branch (aggregators.py:197-214) + GeneformerTokenizer.tokenize (tokenizers.py:635-722) */
    SLAFB_GENEFORMER_PCT = 1, /* GeneformerWindow min_percentile branch (aggregators.py:167-196) + same tokenizer                           */
    SLAFB_SCGPT_RAW = 2,      /* ScGPTWindow use_binned_expressions=False / SimpleWindow (aggregators.py:119-134,226-256) + ScGPTTokenizer  */
    SLAFB_SCGPT_BINNED = 3    /* ScGPTWindow binned branch (aggregators.py:87-118) + ScGPTTokenizer.tokenize (tokenizers.py:375-479)        */
} slafb_mode;

typedef enum { SLAFB_LOC_DEVICE = 0, SLAFB_LOC_HOST = 1 } slafb_location;

/* A batch of COO rows = the `complete_df` of slaf/ml/datasets.py:956.  Rows of one cell must
 * be contiguous (the host feeder splices partial cells, datasets.py:880-888).  Column base
 * pointers must be 16-byte aligned. */
typedef struct {
    const void *cell;
    const void *gene;
    const void *value;
    int64_t n_rows;
    int32_t cell_dtype;  /* SLAFB_U32 | SLAFB_I32 */
    int32_t gene_dtype;  /* SLAFB_U16 | SLAFB_I32 | SLAFB_U32 */
    int32_t value_dtype; /* SLAFB_U16 | SLAFB_F32 */
    int32_t location;    /* slafb_location: device pointers, or host pointers (pinned for async H2D) */
} slafb_coo;

typedef enum {
    SLAFB_SHUFFLE_NONE = 0, /* cells in order of first appearance                                           */
    SLAFB_SHUFFLE_SEED = 1, /* CPython random.seed(seed); random.shuffle(chunks)  (slaf/ml/samplers.py:89-96) */
    SLAFB_SHUFFLE_PERM = 2  /* explicit order supplied in `perm` (host): output row j = segment perm[j]; n_perm may be
                               smaller than the number of segments (a selection: the other cells are not emitted)  */
} slafb_shuffle;

/* Order of the kept genes of a cell.  The reference keeps ROW order (aggregators.py:198-214, pinned by
 * tests/test_aggregators.py:146-174); SLAFB_ORDER_RANK is the textbook Geneformer rank-value encoding
 * (descending value, ties in row order) offered as a clearly separate extension. */
typedef enum { SLAFB_ORDER_ROW = 0, SLAFB_ORDER_RANK = 1 } slafb_order;

/* params.flags */
#define SLAFB_FLAG_CHECK_CONTIGUOUS 1 /* verify that no cell id occurs in two separated runs (one extra sync) */

typedef struct {
    int32_t mode;          /* slafb_mode */
    int32_t max_items;     /* Window.apply(max_items)  = tokenizer.max_genes in the dataloader (datasets.py:1041) */
    int32_t max_genes;     /* SLAFTokenizer.max_genes: L = max_genes (Geneformer) or max_genes + 2 (scGPT)          */
    int32_t n_bins;        /* n_expression_bins */
    int64_t vocab_size;    /* ScGPTTokenizer.expr_bin_start (tokenizers.py:507) */
    double min_percentile; /* SLAFB_GENEFORMER_PCT only; 0..100 */
    int32_t shuffle;       /* slafb_shuffle */
    int32_t flags;         /* SLAFB_FLAG_* */
    int64_t seed;          /* seed + batch_id + epoch*10000 (datasets.py:1016) */
    const int32_t *perm;   /* SLAFB_SHUFFLE_PERM */
    int64_t n_perm;
    int32_t window_n_bins; /* SLAFB_SCGPT_BINNED: n_expression_bins the WINDOW bins with (PrefetchBatchProcessor's,
                              datasets.py:1025-1028) when it differs from the tokenizer's n_bins; 0 = same as n_bins */
    int32_t order;         /* slafb_order; SLAFB_GENEFORMER only */
    /* EXTENSION (not in the reference, BASELINE.json config 4): per-gene normaliser, float32 [n_norm] on the
     * device.  Ranks are taken on float32(value) / norm[gene] (genes >= n_norm: 1.0).  NULL = the
     * reference's behaviour (rank on the raw value).  SLAFB_GENEFORMER only. */
    const float *norm;
    int64_t n_norm;
} slafb_params;

/* Output tensors (device, row-major, caller-allocated for `capacity_cells` rows):
 *   input_ids      int64 [capacity_cells, L]
 *   attention_mask uint8 [capacity_cells, L]   (torch.bool storage)
 *   values         int64 [capacity_cells, L]   scGPT modes only, else NULL
 *   cell_ids       int64 [capacity_cells]      grouped row order (datasets.py:1087-1089) */
typedef struct {
    int64_t *input_ids;
    uint8_t *attention_mask;
    int64_t *values;
    int64_t *cell_ids;
    int64_t capacity_cells;
} slafb_out;

/* Per-kernel device time accumulated since the last call (CUDA events on the launch stream). */
typedef struct {
    double csr_ms;      /* K1  slafb_csr_build                */
    double tokenize_ms; /* K2  slafb_tokenize_cells           */
    double slowpath_ms; /* K2s slafb_tokenize_cells_general   */
    double h2d_ms;      /* staged host->device copies         */
    int64_t batches;
    int64_t kernel_launches; /* kernels launched since the last call */
    int64_t cells;
    int64_t rows;
    int64_t slow_cells;      /* cells that needed the general (distinct-count) path */
    double host_submit_ms;   /* host wall time spent inside slafb_submit                      */
    double host_collect_ms;  /* host wall time spent inside slafb_collect (includes the two below) */
    double host_wait_ms;     /* ... of which: waiting for the cell count (H2D + K1 to finish)   */
    double host_shuffle_ms;  /* ... of which: building the CPython-exact permutation            */
    int64_t timed_batches;   /* batches that carried the per-kernel events (csr_ms .. h2d_ms sum over these) */
} slafb_timing;

int slafb_version(void);

/* Replaces the construction of the window/tokenizer compute state:
 * SLAFTokenizer.__init__ -> create_window() (slaf/ml/tokenizers.py:42-76) and
 * PrefetchBatchProcessor.__init__ (slaf/ml/datasets.py:312-435), compute part only.
 * max_rows / max_cells bound one batch (prefetch_batch_size * n_scanners rows). */
int slafb_create(slafb_ctx **ctx, int device, int64_t max_rows, int64_t max_cells);
int slafb_destroy(slafb_ctx *ctx);
const char *slafb_last_error(const slafb_ctx *ctx);

/* The fused hot path, replacing slaf/ml/datasets.py:1011-1098
 * (shuffle.apply -> tokenizer.window.apply -> .to_list() -> tokenizer.tokenize -> cell ids).
 * Synchronous with respect to *n_cells; token tensors are complete in `stream` order.
 * If out->capacity_cells is too small, returns SLAFB_ERR_CAPACITY with *n_cells set. */
int slafb_tokenize(slafb_ctx *ctx, const slafb_coo *in, const slafb_params *p, const slafb_out *out,
                   int64_t *n_cells, void *stream);

/* Same path, pipelined (double buffered): submit() enqueues the H2D copy (host input) and the
 * CSR build on the context's side stream and returns a slot; collect() waits for the cell
 * count, applies the shuffle and launches the tokenizing kernels on `stream`.  Eight slots:
 * up to eight batches may be in flight and they may be collected in any order (a ninth submit returns SLAFB_ERR_BUSY and
 * changes nothing).  A submit takes a free slot whose previous tokenization has finished and that already owns staging
 * columns, so a caller with one or two batches in flight touches two or three slots' worth of staging (12 B/row each),
 * a deep pipeline all eight.  Small host pieces in pageable memory are copied through a pinned scratch of the slot first:
 * a pageable async copy would make the calling thread wait for the copy stream to drain.  Replaces the AsyncPrefetcher
 * hand-off (slaf/ml/datasets.py:1242-1294) for the device-side part of the work. */
int slafb_submit(slafb_ctx *ctx, const slafb_coo *in, void *stream, int32_t *slot);
/* submit() for a batch that is the concatenation of n_pieces row blocks (all host or all device, same
 * dtypes): the partial cells carried over plus the fragment chunks just read (datasets.py:748-760,
 * 880-888).  Host blocks are copied into the slot's staging columns straight from where they live,
 * so with pinned / registered source memory the feeder does no host-side assembly at all. */
int slafb_submit_pieces(slafb_ctx *ctx, const slafb_coo *pieces, int32_t n_pieces, void *stream, int32_t *slot);
/* submit() for a feeder that already KNOWS the cell boundaries of the rows it hands over -- SLAF keeps the row offset of
 * every cell in SLAFArray._cell_start_index (slaf/core/slaf.py:645-702; the reference uses it for its completeness
 * check, slaf/ml/datasets.py:891-928) -- so the cell column neither crosses PCIe (half of the bytes) nor needs a CSR
 * build.  pieces[i].cell is ignored (may be NULL; cell_dtype still says how to widen the ids); seg_start[n_cells + 1]
 * (row offsets into the concatenated pieces, seg_start[0] = 0, seg_start[n_cells] = total rows) and seg_cell[n_cells]
 * are HOST arrays.  Everything downstream (collect, collect_csr, segments) is unchanged. */
int slafb_submit_segments(slafb_ctx *ctx, const slafb_coo *pieces, int32_t n_pieces, const uint32_t *seg_start,
                          const uint32_t *seg_cell, int64_t n_cells, void *stream, int32_t *slot);

/* submit() for a feeder that has the index itself: the segment table is derived INSIDE the library from
 * SLAFArray._cell_start_index (slaf/core/slaf.py:645-702) and the table position of each piece, replacing the host side of the
 * reference's completeness bookkeeping (slaf/ml/datasets.py:891-928) for the rows handed over.  Piece i holds the table rows
 * [row_lo[i], row_lo[i] + pieces[i].n_rows); csi[c] is the first table row of cell c, csi[n_csi - 1] the total row count (int64,
 * host).  Pieces that continue one cell back to back give ONE segment; a cell that occurs in two separated places gives two
 * segments with the same id (the caller splices, as after slafb_segments).  check: 0 = trust the index, 1 = compare every 16th
 * segment head (and each piece's last row) with pieces[i].cell, 2 = every head; a mismatch is SLAFB_ERR_INVALID and nothing is
 * submitted.  pieces[i].cell may be NULL (nothing to compare with); it never crosses the bus.  *n_cells = segments found; the table
 * itself is available through slafb_segments() without waiting for the GPU. */
int slafb_submit_csi(slafb_ctx *ctx, const slafb_coo *pieces, int32_t n_pieces, const int64_t *row_lo, const int64_t *csi,
                     int64_t n_csi, int32_t check, void *stream, int32_t *slot, int64_t *n_cells);

/* The host half of slafb_submit_csi on its own (no GPU involved, no context needed): the segment table the index implies for
 * the pieces, written to seg_start[capacity_cells + 1] / seg_cell[capacity_cells]. */
int slafb_csi_segments(const slafb_coo *pieces, int32_t n_pieces, const int64_t *row_lo, const int64_t *csi, int64_t n_csi,
                       int32_t check, uint32_t *seg_start, uint32_t *seg_cell, int64_t capacity_cells, int64_t *n_cells);

/* ---- packed wire format (bytes not sent) -----------------------------------------------------------------
 * End to end the path is bound by PCIe.  The expression table's rows are cell-major with gene ids ascending inside a cell and
 * small integer values (slaf/data/converter.py:2819,3025), so a cell's rows pack losslessly into ~2 bytes per row: one 16-byte
 * aligned block per cell = header {u32 n_rows, u32 gene id of row 0, u32 n_exc_g, u32 n_exc_v} | n_rows bytes min(delta, 255),
 * delta_i = gene_i - gene_{i-1} - 1, zero-padded to 16 | n_rows bytes min(value, 255), padded | {u32 row, u32 delta} per delta >= 255
 * | {u32 row, u32 value} per value >= 255 | padding to 16 (full description: csrc/pack_rows.h).  blk_off[n_cells + 1] holds the block
 * offsets in 16-byte units.  Packing costs host cycles once (at conversion time, or in a feeder with cores to spare); every epoch then
 * ships half of the cell-start-index route's bytes and a quarter of the COO route's, and the tokenizing kernel decodes the blocks inside
 * its shared-memory stage.  uint16 values only (float32 tables travel unpacked).  Token tensors are identical to the other routes. */
typedef struct {
    const void *bytes;         /* the blocks, 16-byte aligned (host: pinned for async DMA; or device) */
    int64_t n_bytes;           /* multiple of 16 */
    const uint32_t *blk_off;   /* HOST [n_cells + 1]: block offsets in 16-byte units, relative to `bytes` */
    const uint32_t *row_start; /* HOST [n_cells + 1]: row offsets of the cells (only the differences are used) */
    int64_t n_cells;
    int32_t location;          /* slafb_location of `bytes` */
    int32_t gene_dtype;        /* what the gene ids decode to: SLAFB_U16 (n_genes <= 65536) | SLAFB_I32 */
} slafb_packed;
/* Upper bound of the packed size for typical count data (room for one exception per 16 rows); pack_rows reports SLAFB_ERR_CAPACITY
 * if a batch needs more. */
int64_t slafb_pack_bound(int64_t n_rows, int64_t n_cells);
/* Host only, single threaded (callers pack disjoint cell ranges in parallel and submit them as pieces): packs the cells whose rows
 * are gene[row_start[c] .. row_start[c + 1]) / value[...] into `out`.  SLAFB_ERR_INVALID when some cell's gene ids do not ascend
 * strictly (such rows travel unpacked). */
int slafb_pack_rows(const void *gene, int32_t gene_dtype, const uint16_t *value, const uint32_t *row_start, int64_t n_cells, uint8_t *out,
                    int64_t out_capacity, uint32_t *blk_off, int64_t *out_bytes);
/* Host only: the exact inverse (tests; callers that want the columns back). */
int slafb_unpack_rows(const uint8_t *packed, const uint32_t *blk_off, const uint32_t *row_start, int64_t n_cells, void *gene, int32_t gene_dtype,
                      uint16_t *value);
/* submit() for packed rows: the pieces' blocks are copied back to back, seg_cell[sum of n_cells] (host) are the cell ids in block
 * order; the cell boundaries are the blocks themselves, so no cell column travels and no CSR build runs (as with submit_segments).
 * collect() tokenizes the slot exactly like any other; collect_csr / the rank-value extension take unpacked rows. */
int slafb_submit_packed(slafb_ctx *ctx, const slafb_packed *pieces, int32_t n_pieces, const uint32_t *seg_cell, int32_t cell_dtype,
                        void *stream, int32_t *slot);

/* Blocks until the H2D copy of the batch last submitted into `slot` has finished: its host columns may be reused. */
int slafb_wait_loaded(slafb_ctx *ctx, int32_t slot);

/* Optional hint between submit() and collect(): the shuffle seed the batch will be collected with
 * (seed + batch_id + epoch*10000, datasets.py:1016).  The context's helper thread then builds the
 * CPython-exact permutation as soon as the CSR build has produced the cell count, off the caller's
 * critical path; collect() with SLAFB_SHUFFLE_SEED and the same seed picks it up. */
int slafb_prepare_shuffle(slafb_ctx *ctx, int32_t slot, int64_t seed);
int slafb_collect(slafb_ctx *ctx, int32_t slot, const slafb_params *p, const slafb_out *out,
                  int64_t *n_cells, void *stream);

/* A steady pipeline over one context, for callers that tokenize batch after batch with fixed parameters (the AsyncPrefetcher loop,
 * slaf/ml/datasets.py:1242-1294): ONE call per step and no allocation per step.  `ring` holds n_ring sets of caller-owned output
 * tensors (each with capacity_cells rows); push() submits a batch (COO pieces, or gene / value pieces plus a host segment table when
 * seg_start != NULL) with its shuffle seed and, once more than `depth` batches are in flight, collects the OLDEST one into the next
 * ring entry (round robin) on `stream`; pop() collects the oldest without submitting.  done->ring_index is -1 when the call completed
 * no batch.  Ring entry i is rewritten n_ring completions later in `stream` order, so consumers on that stream need no further
 * synchronisation; consumers elsewhere order themselves with an event.  p->shuffle is SLAFB_SHUFFLE_SEED (seed per push) or
 * SLAFB_SHUFFLE_NONE.  A batch that does not fit a ring entry is dropped with SLAFB_ERR_CAPACITY. */
typedef struct slafb_pipe slafb_pipe;
typedef struct {
    int32_t ring_index; /* ring entry holding the completed batch, or -1 */
    int32_t reserved;
    int64_t n_cells;
    int64_t seq;        /* push order of the completed batch (0, 1, 2, ...) */
} slafb_pipe_result;
int slafb_pipe_create(slafb_ctx *ctx, const slafb_params *p, const slafb_out *ring, int32_t n_ring, int32_t depth, slafb_pipe **pipe);
int slafb_pipe_push(slafb_pipe *pipe, const slafb_coo *pieces, int32_t n_pieces, const uint32_t *seg_start, const uint32_t *seg_cell,
                    int64_t n_seg_cells, int64_t seed, void *stream, slafb_pipe_result *done);
int slafb_pipe_push_packed(slafb_pipe *pipe, const slafb_packed *pieces, int32_t n_pieces, const uint32_t *seg_cell, int32_t cell_dtype,
                           int64_t seed, void *stream, slafb_pipe_result *done);
int slafb_pipe_pop(slafb_pipe *pipe, void *stream, slafb_pipe_result *done);
int slafb_pipe_destroy(slafb_pipe *pipe);

/* Raw mode on the device, replacing RawPrefetchBatch (slaf/ml/datasets.py:288-296, 958-1008: shuffle only, no
 * window, no tokenizer): the submitted batch as a CSR matrix whose rows are the cells in the requested order
 * (p->shuffle / seed / perm as in collect(); a perm may be a selection).  Device outputs, caller allocated:
 * crow int64 [capacity_cells + 1], col int32 [capacity_rows] (gene ids), val [capacity_rows] in the batch's value
 * dtype, cell_ids int64 [capacity_cells].  Synchronous with respect to *n_cells and *n_rows; always releases the slot. */
int slafb_collect_csr(slafb_ctx *ctx, int32_t slot, const slafb_params *p, int64_t *crow, int32_t *col, void *val,
                      int64_t *cell_ids, int64_t capacity_cells, int64_t capacity_rows, int64_t *n_cells, int64_t *n_rows,
                      void *stream);

/* The segment table K1 built for a submitted batch, copied to the host (blocks until the H2D copy
 * and CSR build of that slot are done): n_cells segments in order of first appearance,
 * seg_start[n_cells + 1] row offsets, seg_cell[n_cells] cell ids (uint32 bits of the cell column).
 * The host feeder uses it for the reference's completeness check against _cell_start_index
 * (slaf/ml/datasets.py:891-928) without scanning the rows itself, then collects with a
 * SLAFB_SHUFFLE_PERM selection that leaves the incomplete boundary cells out. */
int slafb_segments(slafb_ctx *ctx, int32_t slot, int64_t *n_cells, uint32_t *seg_start, uint32_t *seg_cell,
                   int64_t capacity_cells);

/* Window facade: Window.apply(fragment_df, schema, max_items, **kwargs)
 * (slaf/ml/aggregators.py:26-50; slaf/distributed/window.py:25-59).  Produces the grouped
 * frame as flat list columns on the device:
 *   cell_ids int64 [C], offsets int64 [C+1], gene_list int64 [<= n_rows],
 *   expr_list float64 [<= n_rows] (raw value or expr_bin; NULL for Geneformer modes). */
int slafb_window(slafb_ctx *ctx, const slafb_coo *in, const slafb_params *p, int64_t *cell_ids,
                 int64_t *offsets, int64_t *gene_list, double *expr_list, int64_t capacity_cells,
                 int64_t *n_cells, int64_t *n_items, void *stream);

/* Tokenizer facade: SLAFTokenizer.tokenize(gene_sequences, expr_sequences)
 * (slaf/ml/tokenizers.py:262-273, 375-479, 635-722) on already-windowed, flattened lists
 * (device): offsets int64 [B+1], genes int64, exprs float64 (NULL for Geneformer).
 * exprs_are_int selects the isinstance(exprs[0], int) branch (tokenizers.py:441). */
int slafb_tokenize_lists(slafb_ctx *ctx, int32_t scgpt, const int64_t *offsets, const int64_t *genes,
                         const double *exprs, int32_t exprs_are_int, int64_t n_lists,
                         int32_t max_genes, int32_t n_bins, int64_t vocab_size, const slafb_out *out,
                         void *stream);

/* RandomShuffle.apply's chunk order (slaf/ml/samplers.py:89-96): bit-exact CPython
 * random.seed(seed); random.shuffle(list(range(n))) computed on the host in C. */
int slafb_shuffle_perm(int64_t seed, int64_t n, int32_t *perm);

/* Extension (not in the reference; BASELINE.json config 4): per-gene nonzero sum and count
 * accumulated INTO sum/cnt (uint64, device, length n_genes) so that ranks can all-reduce them. */
int slafb_gene_stats(slafb_ctx *ctx, const slafb_coo *in, int64_t n_genes, uint64_t *sum,
                     uint64_t *cnt, void *stream);

/* Extension (config 4): per-gene histogram of the uint16 values, accumulated INTO hist (uint32, device,
 * [n_genes x n_bins], values >= n_bins - 1 in the last bin) so that ranks can all-reduce it; and the exact
 * nonzero median of every gene from such a table (numpy.median semantics); exact[g] = 0 when a middle
 * value fell into the overflow bin. */
int slafb_gene_hist(slafb_ctx *ctx, const slafb_coo *in, int64_t n_genes, int32_t n_bins, uint32_t *hist, void *stream);
int slafb_gene_median(slafb_ctx *ctx, const uint32_t *hist, int64_t n_genes, int32_t n_bins, float *median,
                      uint8_t *exact, void *stream);

/* Pinned host memory for the feeder's staging buffers (so H2D copies are truly async). */
int slafb_host_alloc(void **ptr, size_t bytes);
int slafb_host_free(void *ptr);
/* Page-lock memory the caller already owns (e.g. the column buffers of an in-memory / memory-mapped
 * expression table), so slices of it are DMA sources without a staging copy.  Read-only mappings are
 * retried with cudaHostRegisterReadOnly; SLAFB_ERR_CUDA when the driver refuses both. */
int slafb_host_register(void *ptr, size_t bytes);
int slafb_host_unregister(void *ptr);

/* Context options.
 *   SLAFB_OPT_TIMING_INTERVAL  n >= 1: the per-kernel CUDA-event timings are taken on every n-th batch
 *                              (default 1); 0: never (no timing events in any stream: bench.py's throughput passes).
 *   SLAFB_OPT_FRONT_ON_CALLER  0|1: run the CSR build of DEVICE-resident batches on the caller's stream
 *                              instead of the side stream (default 0).  Kernels then never overlap, which
 *                              is how bench.py measures each kernel's own duration. */
typedef enum { SLAFB_OPT_TIMING_INTERVAL = 1, SLAFB_OPT_FRONT_ON_CALLER = 2 } slafb_option;
int slafb_set_option(slafb_ctx *ctx, int32_t option, int64_t value);

/* Checked builds only (nvcc -DSLAFB_CHECKED: every stage / table / output index is compared with its bound on the device): the
 * number of violated checks on the current device since the library was loaded, info[8] = {count, code, value, CTA, thread, ...} of
 * the first.  -1 = the library was not built with checks. */
int slafb_debug_check_failures(uint32_t *info);

/* The float32 log1p of the binned window on float32 storage (polars keeps Float32 and calls Rust's f32::ln_1p = the platform's
 * log1pf, slaf/ml/aggregators.py:96-105).  The device carries two evaluations -- a restatement of the fdlibm float algorithm
 * (what glibc <= 2.39 ships; kind 1) and the correctly rounded value (glibc >= 2.40; kind 0) -- and uses the one this host's libm
 * matches, so its float32 bins equal the reference's on the same machine bit for bit.  slafb_debug_log1pf evaluates one of the two
 * on the HOST for n values (no GPU needed; the device code is the same source with round-to-nearest intrinsics) and returns the
 * kind the library selects on this host (1 or 0); kind < 0 only asks. */
int slafb_debug_log1pf(int32_t kind, const float *x, float *y, int64_t n);

/* Timing / counters since the previous call (synchronises the context's streams). */
int slafb_get_timing(slafb_ctx *ctx, slafb_timing *t);

#ifdef __cplusplus
}
#endif
#endif /* SLAF_B200_H */
