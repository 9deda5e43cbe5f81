"""A third, independent opinion on the window half of the oracle (CPU only).

polars is not installable in this image, so ``slaf/ml/aggregators.py`` cannot be executed; the C oracle and the numpy
restatement were both written from a reading of it.  pandas IS installed and has the same relational vocabulary the reference's
expressions use, so this file transcribes those expressions into pandas one to one and compares the result with
``c_oracle.window`` on adversarial frames (ties at the k-th rank, more than k distinct values, > k rows with few distinct values,
single-row cells, non-monotone cell ids), all four modes x both storage dtypes:

    pl.col(v).rank(method="dense", descending=True).over(g)   ->  df.groupby(g, sort=False)[v].rank(method="dense", ascending=False)
    .filter(rank <= max_items [& pct_rank >= p * max(pct_rank).over(g) / 100])   ->  boolean mask, f64 compare
    log1p / * n / max().over(g) / floor / clip / when(>0).otherwise(0)           ->  the same chain, same evaluation order, in the
                                                                                   dtype polars keeps (f64 for integer storage,
                                                                                   f32 for Float32 storage)
    .group_by(g, maintain_order=True).agg(list)                ->  groupby(g, sort=False).agg(list)   (first-appearance order,
                                                                                                     rows in row order)

Only the ranking / filtering / grouping logic comes from pandas; log1p goes through glibc (math.log1p, libm log1pf), the function
Rust's std calls on linux-gnu, so the comparison does not depend on numpy's SIMD log1p.
"""

import ctypes
import math

import numpy as np
import pandas as pd
import pytest

from oracle import c_oracle
from tests.util import adversarial_batch

MODES = [("geneformer", None), ("geneformer_pct", 37.5), ("geneformer_pct", 10), ("geneformer_pct", 100.0), ("scgpt_raw", None),
         ("scgpt_binned", None)]


def _log1p_like_polars(v: np.ndarray) -> np.ndarray:
    uniq, inv = np.unique(v, return_inverse=True)
    if v.dtype == np.float32:
        libm = ctypes.CDLL("libm.so.6")
        libm.log1pf.restype = ctypes.c_float
        libm.log1pf.argtypes = [ctypes.c_float]
        return np.array([libm.log1pf(float(x)) for x in uniq], np.float32)[inv]
    return np.array([math.log1p(float(x)) for x in uniq], np.float64)[inv]


def pandas_window(cell, gene, value, mode, max_items, n_bins=10, min_percentile=None):
    """slaf/ml/aggregators.py:87-134,167-214 transcribed into pandas."""
    df = pd.DataFrame({"cell": cell, "gene": gene, "value": value})
    by = df.groupby("cell", sort=False)["value"]
    df["gene_rank"] = by.rank(method="dense", ascending=False)
    keep = df["gene_rank"] <= max_items
    if mode == "geneformer_pct":
        df["percentile_rank"] = by.rank(method="dense", ascending=True)
        top = df.groupby("cell", sort=False)["percentile_rank"].transform("max")
        keep &= df["percentile_rank"].astype(np.float64) >= (min_percentile * top.astype(np.float64)) / 100
    out = df[keep].copy()
    if mode == "scgpt_binned":
        ft = np.float32 if value.dtype == np.float32 else np.float64
        lv = _log1p_like_polars(out["value"].to_numpy())
        out["lv"] = lv
        m = out.groupby("cell", sort=False)["lv"].transform("max").to_numpy().astype(ft)
        with np.errstate(all="ignore"):
            b = np.clip(np.floor((lv * ft(n_bins)) / m), 0, n_bins - 1)
        out["expr"] = np.where(lv > 0, b, ft(0)).astype(np.float64)
    elif mode == "scgpt_raw":
        out["expr"] = out["value"].astype(np.float64)
    agg = {"gene": list} | ({"expr": list} if mode.startswith("scgpt") else {})
    grouped = out.groupby("cell", sort=False).agg(agg)
    return grouped.index.to_list(), grouped["gene"].to_list(), (grouped["expr"].to_list() if mode.startswith("scgpt") else None)


def _lists(offs, flat):
    return [flat[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]


@pytest.mark.parametrize("seed", [3, 11])
@pytest.mark.parametrize("value_dtype", ["uint16", "float32"])
@pytest.mark.parametrize("mode,pct", MODES)
def test_c_oracle_window_equals_pandas_transcription(mode, pct, value_dtype, seed):
    k = 24
    cell, gene, value = adversarial_batch(seed, value_dtype=value_dtype, gene_dtype="int32", cell_dtype="int32", k=k, big=False)
    cids, offs, genes, exprs = c_oracle.window(cell, gene, value, mode, k, n_bins=51, min_percentile=pct)
    p_cells, p_genes, p_exprs = pandas_window(cell, gene, value, mode, k, n_bins=51, min_percentile=pct)
    assert cids.tolist() == [int(c) for c in p_cells]
    assert _lists(offs, genes) == [[int(g) for g in row] for row in p_genes]
    if p_exprs is not None:
        assert _lists(offs, exprs) == [[float(x) for x in row] for row in p_exprs]


@pytest.mark.parametrize("value_dtype", ["uint16", "float32"])
def test_pandas_transcription_on_count_like_cells_at_scale(value_dtype):
    """Shapes like the benchmark's: a few hundred cells of ~600 rows, geometric counts (heavy ties), k below the number of
    distinct values for some cells, min_percentile filter included."""
    from slaf_b200 import synthetic

    b = synthetic.make_batch(192, 20000, 600, seed=5, value_dtype=value_dtype)
    for mode, pct, k in (("geneformer", None, 6), ("geneformer_pct", 25.0, 512), ("scgpt_binned", None, 8), ("scgpt_raw", None, 3)):
        cids, offs, genes, exprs = c_oracle.window(b.cell, b.gene, b.value, mode, k, n_bins=51, min_percentile=pct)
        p_cells, p_genes, p_exprs = pandas_window(b.cell, b.gene, b.value, mode, k, n_bins=51, min_percentile=pct)
        assert cids.tolist() == [int(c) for c in p_cells]
        assert _lists(offs, genes) == [[int(g) for g in row] for row in p_genes]
        if p_exprs is not None:
            assert _lists(offs, exprs) == [[float(x) for x in row] for row in p_exprs]
