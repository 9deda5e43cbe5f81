"""CPU tests of the oracle itself: it must reproduce every golden vector the reference's own
tests hold for this path, the outputs of the reference's tokenizers.py (tests/golden), and the
independent numpy/Python restatement must agree with the C restatement."""

import json
import os
import random

import numpy as np
import pytest

from oracle import c_oracle
from oracle import py_restatement as pr
from tests.util import adversarial_batch, assert_same

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "tokenizer_golden.npz"))
KATS = json.load(open(os.path.join(HERE, "golden", "window_kats.json")))["cases"]


def _lists(offs, flat):
    return [flat[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]


@pytest.mark.parametrize("case", KATS, ids=[c["cite"][:48] for c in KATS])
def test_window_kats_c_and_py(case):
    cell = np.asarray(case["cell"], np.int32)
    gene = np.asarray(case["gene"], np.int32)
    value = np.asarray(case["value"], case["value_dtype"])
    cids, offs, genes, exprs = c_oracle.window(cell, gene, value, case["mode"], case["max_items"], n_bins=case["n_bins"],
                                               min_percentile=case.get("min_percentile"))
    assert cids.tolist() == case["expect_cells"]
    got_genes = _lists(offs, genes)
    assert got_genes == case["expect_genes"]
    if "expect_exprs" in case:
        assert _lists(offs, exprs) == [[float(x) for x in row] for row in case["expect_exprs"]]
    if "expect_bin_range" in case:
        lo, hi = case["expect_bin_range"]
        assert all(lo <= b < hi for b in exprs.tolist())
        assert all(float(b).is_integer() for b in exprs.tolist())
    if "expect_len" in case:
        assert [len(g) for g in got_genes] == case["expect_len"]
    # the independent numpy restatement gives the same frame
    kind = {"geneformer": "geneformer", "geneformer_pct": "geneformer", "scgpt_raw": "scgpt", "scgpt_binned": "scgpt"}[case["mode"]]
    extra = {"min_percentile": case["min_percentile"]} if "min_percentile" in case else {}
    w = pr.window_apply(cell, gene, value, kind, case["max_items"], n_expression_bins=case["n_bins"],
                        use_binned_expressions=case["mode"] == "scgpt_binned", **extra)
    assert w["cell_integer_id"] == case["expect_cells"]
    assert w["gene_sequence"] == case["expect_genes"]
    if kind == "scgpt" and len(cell):
        assert [list(map(float, r)) for r in w["expr_sequence"]] == _lists(offs, exprs)


def test_window_bins_documented_values():
    # SURVEY appendix A worked example: vals [5,3,1] -> n=10 [9,7,3]; n=5 [4,3,1]; n=3 [2,2,1]
    cell = np.zeros(3, np.int32)
    gene = np.array([10, 20, 30], np.int32)
    for n, want in ((10, [9, 7, 3]), (5, [4, 3, 1]), (3, [2, 2, 1])):
        for v in (np.array([5, 3, 1], np.uint16), np.array([5, 3, 1], np.float32)):
            _, _, _, e = c_oracle.window(cell, gene, v, "scgpt_binned", 3, n_bins=n)
            assert e.tolist() == [float(x) for x in want]


@pytest.mark.parametrize("tag", ["gf8", "gf64", "gf2048"])
def test_tokenizer_golden_geneformer(tag):
    offs, genes = GOLD[f"{tag}_offsets"], GOLD[f"{tag}_genes"]
    mg = int(GOLD[f"{tag}_max_genes"])
    ids, mask, vals = c_oracle.tokenize_lists(False, offs, genes, None, True, mg)
    assert vals is None
    assert_same(ids, GOLD[f"{tag}_ids"], "ids")
    assert_same(mask, GOLD[f"{tag}_mask"], "mask")
    ids2, mask2, _ = pr.geneformer_tokenize(_lists(offs, genes), mg)
    assert_same(ids2, GOLD[f"{tag}_ids"], "py ids")
    assert_same(mask2, GOLD[f"{tag}_mask"], "py mask")


@pytest.mark.parametrize("tag,is_int", [("sgi4", True), ("sgi1200", True), ("sgf_bins", False), ("sgf_bins10", False),
                                        ("sgf_raw", False), ("sgf_raw5", False)])
def test_tokenizer_golden_scgpt(tag, is_int):
    offs, genes, exprs = GOLD[f"{tag}_offsets"], GOLD[f"{tag}_genes"], GOLD[f"{tag}_exprs"]
    mg, vocab, nb = (int(x) for x in GOLD[f"{tag}_cfg"])
    ids, mask, vals = c_oracle.tokenize_lists(True, offs, genes, exprs, is_int, mg, n_bins=nb, vocab_size=vocab)
    assert_same(ids, GOLD[f"{tag}_ids"], "ids")
    assert_same(mask, GOLD[f"{tag}_mask"], "mask")
    assert_same(vals, GOLD[f"{tag}_vals"], "vals")
    ex_lists = _lists(offs, exprs.astype(np.int64) if is_int else exprs)
    ids2, mask2, vals2 = pr.scgpt_tokenize(_lists(offs, genes), ex_lists, mg, vocab, nb)
    assert_same(ids2, GOLD[f"{tag}_ids"], "py ids")
    assert_same(vals2, GOLD[f"{tag}_vals"], "py vals")
    assert_same(mask2, GOLD[f"{tag}_mask"], "py mask")


def test_binning_kat():
    # tests/test_tokenizers.py:199-213 (re-executed on the reference for float32 inputs)
    x = GOLD["binkat_x"]
    offs = np.array([0, len(x)], np.int64)
    _, _, vals = c_oracle.tokenize_lists(True, offs, np.zeros(len(x), np.int64), x.astype(np.float64), False, len(x),
                                         n_bins=10, vocab_size=1000)
    assert vals[0, 1:1 + len(x)].tolist() == GOLD["binkat_tok"].tolist()
    assert GOLD["binkat_tok"][:5].tolist() == [0, 1001, 1005, 1009, 0]


def test_empty_inputs():
    with pytest.raises(ValueError, match="Gene sequences cannot be empty"):
        pr.geneformer_tokenize([], 8)
    with pytest.raises(ValueError, match="Gene sequences cannot be empty"):
        pr.scgpt_tokenize([], [], 8)
    ids, mask, _ = pr.geneformer_tokenize([[]], 8)            # tests/test_tokenizers.py:174-182
    assert ids.tolist() == [[1, 2, 0, 0, 0, 0, 0, 0]] and mask.sum() == 2


MODES = [("geneformer", {}), ("geneformer_pct", {"min_percentile": 37.5}), ("scgpt_raw", {}), ("scgpt_binned", {})]


@pytest.mark.parametrize("value_dtype", ["uint16", "float32"])
@pytest.mark.parametrize("mode,extra", MODES)
def test_c_oracle_equals_py_restatement(mode, extra, value_dtype):
    k = 24
    cell, gene, value = adversarial_batch(7, value_dtype=value_dtype, gene_dtype="int32", cell_dtype="int32", k=k, big=False)
    seed = 42 + 3
    C = len(np.unique(cell))
    random.seed(seed)
    perm = list(range(C))
    random.shuffle(perm)
    ids, mask, vals, cids = c_oracle.tokenize(cell, gene, value, mode, k, perm=perm, n_bins=51, vocab_size=65000,
                                              min_percentile=extra.get("min_percentile"))
    tk = "geneformer" if mode.startswith("geneformer") else "scgpt"
    ids2, mask2, vals2, cids2 = pr.hot_path(cell, gene, value, tk, k, seed=seed, vocab_size=65000, n_expression_bins=51,
                                            use_binned_expressions=(mode == "scgpt_binned"), **extra)
    assert_same(cids, cids2, "cell ids")
    assert_same(ids, ids2, "ids")
    assert_same(mask, mask2, "mask")
    if vals is not None:
        assert_same(vals, vals2, "vals")


def test_noncontiguous_cells_grouped_like_partition_by():
    # a cell split in two separated runs (slaf/ml/datasets.py:880-888 can produce this)
    cell = np.array([5, 5, 9, 9, 5, 7], np.int32)
    gene = np.array([1, 2, 3, 4, 5, 6], np.int32)
    value = np.array([1, 2, 3, 4, 5, 6], np.uint16)
    cids, offs, genes, _ = c_oracle.window(cell, gene, value, "geneformer", 10)
    assert cids.tolist() == [5, 9, 7]
    assert _lists(offs, genes) == [[1, 2, 5], [3, 4], [6]]
