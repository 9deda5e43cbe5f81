"""numpy / pure-Python restatement of the reference's hot path (TEST INFRASTRUCTURE ONLY).

This module is a *checker* and the `cpu_baseline` "port": only tests/, bench.py's
cpu_baseline / `--impl reference` legs and __graft_entry__.smoke() may import it.
Nothing under slaf_b200/ imports it.

It keeps the reference's *structure* (shuffle -> window -> list hand-off -> per-cell
tokenizer loop) so that timing it says something about the reference's CPU path:

    shuffle   slaf/ml/samplers.py:62-106      (CPython random.seed / random.shuffle on the
                                               list of per-cell chunks, first-appearance order)
    window    slaf/ml/aggregators.py:62-256   (polars rank(dense, desc).over(cell), filter,
                                               group_by(maintain_order=True).agg -> restated
                                               with vectorised numpy; polars is not installable
                                               in this image)
    hand-off  slaf/ml/datasets.py:1052-1059   (.to_list(): Python lists of Python ints/floats)
    tokenize  slaf/ml/tokenizers.py:375-479, 635-722 (per-cell Python loop, numpy rows)

Parity status: the tokenizer functions are pinned against the reference's own
tokenizers.py (tests/golden/tokenizer_golden.npz); the window functions are pinned by the
reference tests' known-answer vectors plus hand-worked vectors for the percentile and
integer-storage binned branches (tests/golden/window_kats.json) and cross-checked against a
pandas transcription of the reference's expressions (tests/test_oracle_crosscheck.py).
"""

from __future__ import annotations

import math
import random
from typing import Any

import numpy as np

PAD, CLS, SEP, MASK = 0, 1, 2, 3


# --------------------------------------------------------------------------- grouping
def first_appearance_groups(cell: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """(group index per row, cell id per group); groups numbered by first appearance."""
    if len(cell) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    uniq, first, inv = np.unique(cell, return_index=True, return_inverse=True)
    rank_of_sorted = np.empty(len(uniq), np.int64)
    rank_of_sorted[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    return rank_of_sorted[inv], uniq[np.argsort(first, kind="stable")].astype(np.int64)


def shuffle_permutation(seed: int, n_cells: int) -> list[int]:
    """Order in which RandomShuffle.apply emits the per-cell chunks (samplers.py:89-96).

    random.shuffle's draws depend only on the list length, so shuffling indices equals
    shuffling the chunk list."""
    random.seed(seed)
    idx = list(range(n_cells))
    random.shuffle(idx)
    return idx


def apply_block_shuffle(cell, gene, value, seed: int):
    """RandomShuffle.apply(df, seed) on column arrays: returns the re-ordered columns."""
    if len(cell) == 0:
        return cell, gene, value
    g, _ = first_appearance_groups(cell)
    n_cells = int(g.max()) + 1
    perm = shuffle_permutation(seed, n_cells)
    pos = np.empty(n_cells, np.int64)
    pos[np.asarray(perm, np.int64)] = np.arange(n_cells)
    order = np.argsort(pos[g], kind="stable")
    return cell[order], gene[order], value[order]


# --------------------------------------------------------------------------- window
def _dense_rank_desc(g: np.ndarray, v: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """per-group dense rank of v, descending (1 = largest) and number of distinct per row's group."""
    n = len(v)
    order = np.lexsort((v, g))            # by group, then value ascending
    gs, vs = g[order], v[order]
    new_grp = np.ones(n, bool)
    new_grp[1:] = gs[1:] != gs[:-1]
    new_val = new_grp.copy()
    new_val[1:] |= vs[1:] != vs[:-1]
    asc_cum = np.cumsum(new_val)
    grp_first = np.maximum.accumulate(np.where(new_grp, np.arange(n), 0))
    asc_rank_sorted = asc_cum - asc_cum[grp_first] + 1      # ascending dense rank
    # distinct count per group = ascending rank of the group's last row
    grp_last = np.ones(n, bool)
    grp_last[:-1] = new_grp[1:]
    last_idx = np.flatnonzero(grp_last)
    seg_id = np.cumsum(new_grp) - 1
    distinct_sorted = asc_rank_sorted[last_idx][seg_id]
    asc = np.empty(n, np.int64)
    dist = np.empty(n, np.int64)
    asc[order] = asc_rank_sorted
    dist[order] = distinct_sorted
    return dist - asc + 1, dist


def window_apply(
    cell: np.ndarray,
    gene: np.ndarray,
    value: np.ndarray,
    kind: str,
    max_items: int,
    **kwargs: Any,
) -> dict[str, Any]:
    """Restates {Geneformer,ScGPT,Simple}Window.apply on column arrays.

    Returns {"cell_integer_id": [...], "gene_sequence": [[...]], "expr_sequence": [[...]] | None}
    with Python ints / floats exactly as polars' .to_list() would hand them over."""
    kind = kind.lower()
    n_bins = kwargs.get("n_expression_bins", 10)
    use_binned = kwargs.get("use_binned_expressions", True)
    min_percentile = kwargs.get("min_percentile", None)
    if len(cell) == 0:
        return {"cell_integer_id": [], "gene_sequence": [], "expr_sequence": None if kind == "geneformer" else []}
    g, group_cell = first_appearance_groups(cell)
    rank, distinct = _dense_rank_desc(g, value)
    keep = rank <= max_items
    if kind == "geneformer" and min_percentile is not None:
        pr = distinct - rank + 1
        keep &= pr.astype(np.float64) >= (min_percentile * distinct.astype(np.float64)) / 100
    expr = None
    if kind in ("scgpt", "simple"):
        if kind == "scgpt" and use_binned:
            if value.dtype == np.float32:
                uniq, inv = np.unique(value, return_inverse=True)
                lv = _log1pf(uniq)[inv]          # polars keeps Float32 end to end
                ftype = np.float32
            else:
                lv = _log1p_lut()[value.astype(np.int64)]   # glibc log1p on the f64 cast, like Rust std
                ftype = np.float64
            n_groups = len(group_cell)
            m = np.full(n_groups, -np.inf, ftype)
            np.maximum.at(m, g[keep], lv[keep])
            with np.errstate(all="ignore"):
                q = (lv * ftype(n_bins)) / m[g]
                b = np.clip(np.floor(q), 0, n_bins - 1)
            expr = np.where(lv > 0, b, ftype(0)).astype(ftype)
        else:
            expr = value
    # group_by(maintain_order=True).agg(list): rows of a cell in row order, cells by first appearance
    rows = np.flatnonzero(keep)
    rows = rows[np.argsort(g[rows], kind="stable")]
    gk = g[rows]
    bounds = np.flatnonzero(np.r_[True, gk[1:] != gk[:-1], True]) if len(rows) else np.array([0])
    present = gk[bounds[:-1]] if len(rows) else np.zeros(0, np.int64)
    gene_py = gene[rows].tolist()
    expr_py = expr[rows].tolist() if expr is not None else None
    out: dict[str, Any] = {
        "cell_integer_id": group_cell[present].tolist(),
        "gene_sequence": [gene_py[bounds[i] : bounds[i + 1]] for i in range(len(present))],
        "expr_sequence": (
            [expr_py[bounds[i] : bounds[i + 1]] for i in range(len(present))] if expr_py is not None else None
        ),
    }
    return out


_LUT = None


def _log1p_lut() -> np.ndarray:
    """log1p(double(v)) for every uint16 v through CPython's math.log1p (= glibc log1p)."""
    global _LUT
    if _LUT is None:
        _LUT = np.array([math.log1p(float(i)) for i in range(65536)], np.float64)
    return _LUT


def _log1pf(x: np.ndarray) -> np.ndarray:
    """glibc log1pf on a float32 array (what Rust's f32::ln_1p calls on linux-gnu)."""
    import ctypes

    libm = ctypes.CDLL("libm.so.6")
    libm.log1pf.restype = ctypes.c_float
    libm.log1pf.argtypes = [ctypes.c_float]
    return np.array([libm.log1pf(float(v)) for v in x], np.float32)


# --------------------------------------------------------------------------- tokenizers
def geneformer_tokenize(gene_sequences: list[list[int]], max_genes: int = 2048):
    """GeneformerTokenizer.tokenize (tokenizers.py:635-722): (ids int64 [B,L], mask bool [B,L], None)."""
    if not gene_sequences:
        raise ValueError("Gene sequences cannot be empty")
    width = max_genes
    ids = np.zeros((len(gene_sequences), width), np.int64)
    for row, seq in enumerate(gene_sequences):
        body = np.asarray(seq, np.int64) + 4
        framed = np.empty(len(body) + 2, np.int64)
        framed[0] = CLS
        framed[1:-1] = body
        framed[-1] = SEP
        framed = framed[:width]
        ids[row, : len(framed)] = framed
    return ids, ids != PAD, None


def scgpt_tokenize(
    gene_sequences: list[list[int]],
    expr_sequences: list[list[float]] | None,
    max_genes: int = 1024,
    vocab_size: int = 50000,
    n_expression_bins: int = 10,
):
    """ScGPTTokenizer.tokenize (tokenizers.py:375-479): dual stream, L = max_genes + 2."""
    if not gene_sequences:
        raise ValueError("Gene sequences cannot be empty")
    width = max_genes + 2
    bin_size = 1.0 / n_expression_bins
    ids = np.zeros((len(gene_sequences), width), np.int64)
    vals = np.zeros((len(gene_sequences), width), np.int64)
    for row, seq in enumerate(gene_sequences):
        ex = expr_sequences[row] if expr_sequences is not None else []
        genes = list(seq) if seq else []
        exprs = list(ex) if ex else []
        n = min(len(genes), len(exprs), max_genes)
        ids[row, 0] = CLS
        if n > 0:
            ids[row, 1 : 1 + n] = np.asarray(genes[:n], np.int64) + 4
            if isinstance(exprs[0], (int, np.integer)):
                vals[row, 1 : 1 + n] = np.asarray(exprs[:n], np.int64) + vocab_size
            else:
                x = np.asarray(exprs[:n], np.float32)
                with np.errstate(all="ignore"):
                    b = np.clip((x / bin_size).astype(int), 0, n_expression_bins - 1)
                vals[row, 1 : 1 + n] = np.where(x > 0, vocab_size + b, PAD)
        ids[row, 1 + n] = SEP
    return ids, ids != PAD, vals


# --------------------------------------------------------------------------- whole path
def hot_path(
    cell: np.ndarray,
    gene: np.ndarray,
    value: np.ndarray,
    tokenizer_type: str,
    max_genes: int,
    seed: int | None = None,
    vocab_size: int = 50000,
    n_expression_bins: int = 10,
    use_binned_expressions: bool = True,
    **window_kwargs: Any,
):
    """complete_df -> (input_ids, attention_mask, values|None, cell_ids) as
    PrefetchBatchProcessor.load_prefetch_batch does it (datasets.py:1011-1098).
    seed=None skips the shuffle (identity order)."""
    if seed is not None:
        cell, gene, value = apply_block_shuffle(cell, gene, value, seed)
    kind = "geneformer" if tokenizer_type == "geneformer" else "scgpt"
    grouped = window_apply(
        cell, gene, value, kind, max_genes,
        n_expression_bins=n_expression_bins, use_binned_expressions=use_binned_expressions, **window_kwargs,
    )
    if kind == "geneformer":
        ids, mask, vals = geneformer_tokenize(grouped["gene_sequence"], max_genes)
    else:
        ids, mask, vals = scgpt_tokenize(
            grouped["gene_sequence"], grouped["expr_sequence"], max_genes, vocab_size, n_expression_bins
        )
    return ids, mask, vals, np.asarray(grouped["cell_integer_id"], np.int64)


# --------------------------------------------------------------------------- extension oracle (NOT in the reference)
def geneformer_rank_value(cell, gene, value, max_genes: int, max_items: int | None = None, norm=None, order: str = "rank",
                          seed: int | None = None):
    """Oracle of the library's rank-value EXTENSION (BASELINE.json config 4); the reference has neither per-gene
    normalisation nor rank order, so this pins the kernels to a written definition, not to
    the reference ("parity unpinned").  Per cell, rows in arrival order:
        x    = float32(value) / float32(norm[gene])      (norm None: x = value)
        keep = rows whose dense rank of x, descending, is <= max_items             (the reference's filter)
        seq  = kept genes by descending x, ties in row order (order="rank"), or in row order (order="row")
    then GeneformerTokenizer.tokenize(seq).  Returns (ids, mask, cell_ids)."""
    k = max_items if max_items is not None else max_genes
    if seed is not None:
        cell, gene, value = apply_block_shuffle(cell, gene, value, seed)
    g, ids = first_appearance_groups(np.asarray(cell))
    order_rows = np.argsort(g, kind="stable")
    bounds = np.concatenate(([0], np.cumsum(np.bincount(g, minlength=len(ids)))))
    seqs = []
    for c in range(len(ids)):
        rows = order_rows[bounds[c]:bounds[c + 1]]
        gg = np.asarray(gene)[rows].astype(np.int64)
        x = np.asarray(value)[rows].astype(np.float32)
        if norm is not None:
            d = np.ones(len(rows), np.float32)
            inside = gg < len(norm)
            d[inside] = np.asarray(norm, np.float32)[gg[inside]]
            x = (x / d).astype(np.float32)
        distinct = np.unique(x)[::-1]
        thr = distinct[min(k, len(distinct)) - 1]
        keep = x >= thr
        if order == "rank":
            idx = np.flatnonzero(keep)
            idx = idx[np.lexsort((idx, -x[idx].astype(np.float64)))]
        else:
            idx = np.flatnonzero(keep)
        seqs.append(gg[idx].tolist())
    tok, mask, _ = geneformer_tokenize(seqs, max_genes)
    return tok, mask, ids
