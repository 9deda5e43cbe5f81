"""Shared builders for parity tests: adversarial COO batches for the window/tokenizer path."""

from __future__ import annotations

import numpy as np


def adversarial_batch(seed: int, value_dtype="uint16", gene_dtype="uint16", cell_dtype="uint32", k: int = 64,
                      big: bool = True):
    """Cells chosen to hit every branch: n<=k, n>k with few distinct values, value range > k
    (uint16 bitmap path), more than k distinct values (real filtering), ties at the k-th rank,
    single-row cells, multi-chunk cells (> 2048 rows), > 12288 rows (global-memory sort for
    float32), non-monotone cell ids."""
    rng = np.random.default_rng(seed)
    G = 60000 if gene_dtype == "uint16" else 70000
    cells = []

    def add(n, vals):
        genes = np.sort(rng.choice(G, size=n, replace=False))
        cells.append((genes, np.asarray(vals)))

    add(1, [3])                                                   # single row
    add(k, rng.integers(1, 6, k))                                 # n == k
    add(k + 1, rng.integers(1, 6, k + 1))                         # n > k, few distinct
    add(3 * k, rng.integers(1, 4 * k, 3 * k))                     # > k distinct: real filter, ties
    add(2 * k, np.r_[rng.integers(1, 5, 2 * k - 1), 60000])       # range > k but few distinct
    add(k + 5, np.arange(1, k + 6))                               # exactly k+5 distinct, drops the 5 smallest
    add(k + 5, np.r_[np.arange(1, k + 1), [1, 1, 2, 2, 3]])       # ties below the threshold
    add(2 * k + 7, np.repeat(np.arange(1, k + 5), 2)[: 2 * k + 7])  # every value twice
    for _ in range(6):
        n = int(rng.integers(2, 5 * k))
        add(n, rng.geometric(0.3, n))
    for _ in range(4):
        n = int(rng.integers(k + 1, 6 * k))
        add(n, rng.integers(1, 3 * k, n))
    if big:
        add(2048 + 300, rng.integers(1, 40, 2348))                # two chunks, keep-all by range
        add(5000, rng.integers(1, 9000, 5000))                    # three chunks, real filter
        add(13000, rng.integers(1, 30000, 13000))                 # beyond the shared-memory sort capacity
    order = rng.permutation(len(cells))
    ids = rng.choice(1_000_000, size=len(cells), replace=False)   # non-monotone cell ids
    cell_col, gene_col, val_col = [], [], []
    for j in order:
        g, v = cells[j]
        cell_col.append(np.full(len(g), ids[j]))
        gene_col.append(g)
        val_col.append(v)
    cell = np.concatenate(cell_col).astype(cell_dtype)
    gene = np.concatenate(gene_col).astype(gene_dtype)
    val = np.concatenate(val_col)
    if value_dtype == "uint16":
        value = np.minimum(val, 65535).astype(np.uint16)
    else:
        # float values with ties: quantised logs (many equal floats) plus a few irrationals
        value = (np.log1p(val.astype(np.float64)) * 1.7).astype(np.float32)
    return cell, gene, value


def assert_same(got, want, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        first = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} mismatches, first at {first}: got {got[first]} want {want[first]}")
