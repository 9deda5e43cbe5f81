"""GPU tests of the EXTENSIONS (BASELINE.json config 4; not in the reference, "parity unpinned"): Geneformer
rank-value encoding in descending-value order, per-gene normalisation, per-gene statistics / exact medians.
Checked against the written definition in oracle/py_restatement.geneformer_rank_value and numpy."""

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import py_restatement as pr
from slaf_b200 import _native, synthetic
from slaf_b200.distributed import gene_medians, gene_stats, normalisation_vector
from tests.util import adversarial_batch, assert_same

pytestmark = pytest.mark.gpu


def _norm(n_genes, seed=0):
    return np.random.default_rng(seed).uniform(0.25, 8.0, n_genes).astype(np.float32)


@pytest.mark.parametrize("order", ["rank", "row"])
@pytest.mark.parametrize("with_norm", [False, True])
@pytest.mark.parametrize("value_dtype,gene_dtype,cell_dtype", [("uint16", "uint16", "uint32"), ("float32", "int32", "int32")])
def test_rank_value_extension_vs_written_definition(engine, order, with_norm, value_dtype, gene_dtype, cell_dtype):
    k = 64
    cell, gene, value = adversarial_batch(41, value_dtype, gene_dtype, cell_dtype, k=k)          # includes a 13,000-row cell (global scratch)
    norm = _norm(70000, 3) if with_norm else None
    for max_genes, max_items, seed in ((k, k, 42), (40, 100, None), (2048, 2048, 7)):
        ids, mask, cids = pr.geneformer_rank_value(cell, gene, value, max_genes, max_items, norm=norm, order=order, seed=seed)
        got = engine.tokenize(cell, gene, value, mode="geneformer", max_genes=max_genes, max_items=max_items, order=order,
                              norm=torch.from_numpy(norm).cuda() if with_norm else None, shuffle_seed=seed)
        assert got.values is None
        assert_same(got.cell_ids.cpu().numpy(), cids, "cell ids")
        assert_same(got.input_ids.cpu().numpy(), ids, f"ids {order} norm={with_norm} L={max_genes}")
        assert_same(got.attention_mask.cpu().numpy(), mask, "mask")


def test_row_order_without_norm_is_the_reference_mode(engine):
    """order='row', norm=None through the extension kernel == the reference-exact Geneformer path."""
    cell, gene, value = adversarial_batch(43, "uint16", "uint16", "uint32", k=32, big=False)
    ones = torch.ones(70000, dtype=torch.float32, device="cuda")
    want = c_oracle.tokenize(cell, gene, value, "geneformer", 32)
    got = engine.tokenize(cell, gene, value, mode="geneformer", max_genes=32, norm=ones)     # forces the extension kernel, identity normaliser
    assert_same(got.input_ids.cpu().numpy(), want[0], "ids")
    assert_same(got.attention_mask.cpu().numpy(), want[1], "mask")
    assert_same(got.cell_ids.cpu().numpy(), want[3], "cells")


def test_rank_order_properties_at_scale(engine):
    """A full-size prefetch batch: every row is CLS + genes in non-increasing normalised value + SEP/PAD."""
    b = synthetic.make_batch(3000, 62000, 2000, seed=8)
    norm = _norm(62000, 5)
    got = engine.tokenize(torch.from_numpy(b.cell).cuda(), torch.from_numpy(b.gene).cuda(), torch.from_numpy(b.value).cuda(),
                          mode="geneformer", max_genes=2048, order="rank", norm=torch.from_numpy(norm).cuda(), shuffle_seed=1)
    ids = got.input_ids.cpu().numpy()
    assert (ids[:, 0] == 1).all() and sorted(got.cell_ids.cpu().tolist()) == list(range(3000))
    for r in (0, 17, 2999):
        c = int(got.cell_ids[r])
        lo, hi = b.cell_start[c], b.cell_start[c + 1]
        x = dict(zip(b.gene[lo:hi].tolist(), (b.value[lo:hi].astype(np.float32) / norm[b.gene[lo:hi]]).tolist()))
        toks = ids[r][(ids[r] > 3)] - 4
        xs = [x[int(g)] for g in toks]
        assert all(xs[i] >= xs[i + 1] for i in range(len(xs) - 1))
        assert len(toks) == min(hi - lo, 2047)
    want = pr.geneformer_rank_value(b.cell[: b.cell_start[64]], b.gene[: b.cell_start[64]], b.value[: b.cell_start[64]], 2048, norm=norm)
    sub = engine.tokenize(b.cell[: b.cell_start[64]], b.gene[: b.cell_start[64]], b.value[: b.cell_start[64]], mode="geneformer",
                          max_genes=2048, order="rank", norm=torch.from_numpy(norm).cuda())
    assert_same(sub.input_ids.cpu().numpy(), want[0], "ids")


def test_extension_argument_errors(engine):
    b = synthetic.make_batch(50, 3000, 40, seed=2)
    with pytest.raises(_native.SlafB200Error) as ei:
        engine.tokenize(b.cell, b.gene, b.value, mode="scgpt_raw", max_genes=16, order="rank")
    assert ei.value.code == _native.ERR_INVALID
    with pytest.raises(ValueError):
        engine.tokenize(b.cell, b.gene, b.value, mode="geneformer", max_genes=16, order="sorted")
    with pytest.raises(TypeError):
        engine.tokenize(b.cell, b.gene, b.value, mode="geneformer", max_genes=16, norm=torch.ones(10))     # not on the device


def test_gene_statistics_means_and_exact_medians(engine):
    n_genes = 500
    batches = [synthetic.make_batch(600, n_genes, 60, seed=30 + i, min_nnz=4) for i in range(3)]
    dev = [(torch.from_numpy(b.gene).cuda(), torch.from_numpy(b.value).cuda()) for b in batches]
    gene = np.concatenate([b.gene for b in batches]).astype(np.int64)
    value = np.concatenate([b.value for b in batches]).astype(np.int64)
    total, count = gene_stats(engine, dev, n_genes)
    assert_same(total.cpu().numpy(), np.bincount(gene, weights=value, minlength=n_genes).astype(np.int64), "sum")
    assert_same(count.cpu().numpy(), np.bincount(gene, minlength=n_genes), "count")
    norm = normalisation_vector(total, count).cpu().numpy()
    want_mean = np.where(np.bincount(gene, minlength=n_genes) > 0,
                         np.bincount(gene, weights=value, minlength=n_genes) / np.maximum(np.bincount(gene, minlength=n_genes), 1), 1.0)
    assert_same(norm, want_mean.astype(np.float32), "mean")
    med, exact = gene_medians(engine, dev, n_genes, n_bins=64)
    med, exact = med.cpu().numpy(), exact.cpu().numpy()
    for g in range(n_genes):
        v = value[gene == g]
        if len(v) == 0:
            assert med[g] == 0 and exact[g]
            continue
        srt = np.sort(v)
        mids = (srt[(len(v) - 1) // 2], srt[len(v) // 2])
        if max(mids) < 63:
            assert exact[g] and med[g] == np.float32(np.median(v)), g
        else:
            assert not exact[g]
