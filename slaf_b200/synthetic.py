"""Deterministic synthetic COO expression data in the reference's on-disk row contract.

Row contract (reference slaf/data/converter.py:2748-2799, 2809-3027): three columns
``cell_integer_id`` (uint32 | int32), ``gene_integer_id`` (uint16 when n_genes <= 65535 else
int32), ``value`` (uint16 counts | float32); rows are cell-major, gene-ascending within a cell
(``tocoo()`` order), no nulls.

Generator: nnz per cell ~ clip(round(lognormal(ln mu, 0.35)), 16, G);
genes = a sorted sample without replacement from [0, G) (spacing method, uniform popularity);
counts ~ geometric(p=0.35) (heavy ties, like UMI counts) with a small heavy tail: a fraction
``tail_frac`` of the cells carries one count in the thousands, as highly expressed genes do in
real data, so that the exact k-th-distinct-value path is exercised and not dead code.
``value_dtype=float32`` gives log1p(count / total * 1e4) (library-size normalised values).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class CooBatch:
    cell: np.ndarray
    gene: np.ndarray
    value: np.ndarray
    n_cells: int
    cell_start: np.ndarray  # int64 [n_cells + 1] row offsets (the _cell_start_index slice of this batch)

    @property
    def n_rows(self) -> int:
        return int(self.cell.shape[0])


def make_batch(
    n_cells: int,
    n_genes: int,
    mean_nnz: float,
    seed: int,
    value_dtype: str = "uint16",
    cell_dtype: str = "uint32",
    gene_dtype: str | None = None,
    first_cell_id: int = 0,
    tail_frac: float = 0.02,
    sigma: float = 0.35,
    min_nnz: int = 16,
) -> CooBatch:
    rng = np.random.default_rng(seed)
    if gene_dtype is None:
        gene_dtype = "uint16" if n_genes <= 65535 else "int32"
    nnz = np.clip(np.rint(rng.lognormal(np.log(mean_nnz), sigma, n_cells)), min(min_nnz, n_genes), n_genes).astype(np.int64)
    start = np.zeros(n_cells + 1, np.int64)
    np.cumsum(nnz, out=start[1:])
    R = int(start[-1])
    cell_of_row = np.repeat(np.arange(n_cells, dtype=np.int64), nnz)
    idx_in_cell = np.arange(R, dtype=np.int64) - start[:-1][cell_of_row]
    # sorted distinct genes per cell: n+1 exponential spacings -> floor(frac * (G - n + 1)) + rank
    gaps = rng.exponential(1.0, R + n_cells)
    gstart = start[:-1] + np.arange(n_cells)            # each cell owns nnz+1 gaps
    csum = np.cumsum(gaps)
    before = np.concatenate(([0.0], csum))[gstart]      # cumulative sum before the cell's first gap
    total = csum[gstart + nnz] - before                 # sum of the cell's nnz+1 gaps
    row_gap_idx = gstart[cell_of_row] + idx_in_cell
    frac = (csum[row_gap_idx] - before[cell_of_row]) / total[cell_of_row]
    span = (n_genes - nnz + 1)[cell_of_row]
    y = np.minimum(np.floor(frac * span).astype(np.int64), span - 1)
    gene = (y + idx_in_cell).astype(gene_dtype)
    counts = np.minimum(rng.geometric(0.35, R), 65535).astype(np.int64)
    if tail_frac > 0 and n_cells > 0:
        heavy = np.flatnonzero(rng.random(n_cells) < tail_frac)
        if heavy.size:
            pos = start[heavy] + rng.integers(0, nnz[heavy])
            counts[pos] = np.minimum(np.rint(rng.lognormal(np.log(2500.0), 0.6, heavy.size)), 65535).astype(np.int64)
    if value_dtype == "uint16":
        value = counts.astype(np.uint16)
    elif value_dtype == "float32":
        tot = np.add.reduceat(counts, start[:-1]).astype(np.float64) if R else np.zeros(0)
        value = np.log1p(counts / tot[cell_of_row] * 1e4).astype(np.float32)
    else:
        raise ValueError(value_dtype)
    cell = (cell_of_row + first_cell_id).astype(cell_dtype)
    return CooBatch(cell, gene, value, n_cells, start)


def batch_nnz(n_cells: int, n_genes: int, mean_nnz: float, seed: int, sigma: float = 0.35, min_nnz: int = 16) -> np.ndarray:
    """Rows per cell of ``make_batch(n_cells, n_genes, mean_nnz, seed, ...)`` without generating the rows (the generator's first
    draw replayed): lets every rank of a sharded run know the row offsets of the whole table."""
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(rng.lognormal(np.log(mean_nnz), sigma, n_cells)), min(min_nnz, n_genes), n_genes).astype(np.int64)


def split_fragments(batch: CooBatch, max_rows_per_file: int) -> list[tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """Contiguous row ranges NOT aligned to cells (cells straddle fragments, converter.py:121)."""
    out = []
    for lo in range(0, batch.n_rows, max_rows_per_file):
        hi = min(lo + max_rows_per_file, batch.n_rows)
        out.append((batch.cell[lo:hi], batch.gene[lo:hi], batch.value[lo:hi]))
    return out
