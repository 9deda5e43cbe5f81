"""GPU tests of the drop-in layer (slaf_b200.ml): the reference's Window / SLAFTokenizer interfaces,
written the way the reference's own tests/test_aggregators.py, tests/test_distributed_window.py
and tests/test_tokenizers.py are, against the reference's golden vectors and the CPU oracle."""

import json
import os
from unittest.mock import Mock

import numpy as np
import pyarrow as pa
import pytest
import torch

from oracle import c_oracle
from slaf_b200.core.tabular_schema import SLAF_LANCE_COO_SCHEMA, DataSchema
from slaf_b200.ml import (GeneformerTokenizer, GeneformerWindow, ScGPTTokenizer, ScGPTWindow, SimpleWindow, create_window)
from slaf_b200.ml.aggregators import GenericWindow
from slaf_b200.ml import frames
from tests.util import adversarial_batch, assert_same

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "tokenizer_golden.npz"))
KATS = json.load(open(os.path.join(HERE, "golden", "window_kats.json")))["cases"]


def _frame(case, kind):
    cols = {"cell_integer_id": np.asarray(case["cell"], np.int32), "gene_integer_id": np.asarray(case["gene"], np.int32),
            "value": np.asarray(case["value"], case["value_dtype"])}
    if kind == "dict":
        return cols
    if kind == "arrow":
        return pa.table(cols)
    import pandas as pd

    return pd.DataFrame(cols)


def _window_for(case):
    if case["mode"] == "geneformer":
        return GeneformerWindow(), {}
    if case["mode"] == "geneformer_pct":
        return GeneformerWindow(), {"min_percentile": case["min_percentile"]}
    if case["mode"] == "scgpt_binned":
        return ScGPTWindow(), {"n_expression_bins": case["n_bins"], "use_binned_expressions": True}
    return ScGPTWindow(), {"n_expression_bins": case["n_bins"], "use_binned_expressions": False}


@pytest.mark.parametrize("kind", ["dict", "arrow", "pandas"])
@pytest.mark.parametrize("case", KATS, ids=[c["cite"][:48] for c in KATS])
def test_window_apply_reference_kats(case, kind):
    win, kw = _window_for(case)
    out = win.apply(_frame(case, kind), SLAF_LANCE_COO_SCHEMA, case["max_items"], **kw)
    assert frames.frame_len(out) == len(case["expect_cells"])
    if not case["expect_cells"]:
        return
    assert [int(x) for x in frames.column(out, "cell_integer_id")] == case["expect_cells"]
    assert frames.list_column(out, "gene_sequence") == case["expect_genes"]
    if "expect_exprs" in case:
        assert frames.list_column(out, "expr_sequence") == [[float(x) for x in r] for r in case["expect_exprs"]]
    if "expect_bin_range" in case:
        lo, hi = case["expect_bin_range"]
        for row in frames.list_column(out, "expr_sequence"):
            assert all(lo <= b < hi for b in row)


def test_window_empty_frame_gives_zero_rows():
    # reference tests/test_aggregators.py:133-144,257-268
    empty = {"cell_integer_id": np.zeros(0, np.int32), "gene_integer_id": np.zeros(0, np.int32), "value": np.zeros(0, np.float32)}
    for win in (ScGPTWindow(), GeneformerWindow(), SimpleWindow()):
        out = win.apply(empty, SLAF_LANCE_COO_SCHEMA, 3)
        assert frames.frame_len(out) == 0
    out = ScGPTWindow().apply(pa.table(empty), SLAF_LANCE_COO_SCHEMA, 3)
    assert out.num_rows == 0 and set(out.column_names) == {"cell_integer_id", "gene_sequence", "expr_sequence"}


def test_generic_window_follows_schema_names():
    # reference tests/test_distributed_window.py:39-61: generic keys, item list only when no value_list_key
    schema = DataSchema(group_key="g", item_key="i", value_key="v", group_key_out="g", item_list_key="items", value_list_key=None)
    df = {"g": np.repeat(np.arange(3, dtype=np.int32), 5), "i": np.tile(np.arange(5, dtype=np.int32), 3),
          "v": np.tile(np.array([5, 4, 3, 2, 1], np.float32), 3)}
    out = GenericWindow().apply(df, schema, 3)
    assert out.columns == ["g", "items"]
    assert [list(x) for x in out["items"]] == [[0, 1, 2]] * 3
    schema2 = DataSchema(group_key="g", item_key="i", value_key="v", group_key_out="g", item_list_key="items", value_list_key="vals")
    out2 = GenericWindow().apply(df, schema2, 3)
    assert [list(x) for x in out2["vals"]] == [[5.0, 4.0, 3.0]] * 3


def test_window_noncontiguous_cells_regrouped_like_partition_by():
    cell = np.array([5, 5, 9, 9, 5, 7], np.int32)
    gene = np.array([1, 2, 3, 4, 5, 6], np.int32)
    value = np.array([1, 2, 3, 4, 5, 6], np.uint16)
    out = GeneformerWindow().apply({"cell_integer_id": cell, "gene_integer_id": gene, "value": value}, SLAF_LANCE_COO_SCHEMA, 10)
    assert list(out["cell_integer_id"]) == [5, 9, 7]
    assert [list(x) for x in out["gene_sequence"]] == [[1, 2, 5], [3, 4], [6]]


@pytest.mark.parametrize("value_dtype", ["uint16", "float32"])
def test_window_then_tokenize_equals_fused_path(engine, value_dtype):
    """window.apply -> .to_list() -> tokenizer.tokenize (the reference's three calls, datasets.py:1038-1059)
    gives the same tensors as the fused kernel path and as the oracle."""
    k = 40
    cell, gene, value = adversarial_batch(17, value_dtype, "int32", "int32", k=k, big=False)
    df = {"cell_integer_id": cell, "gene_integer_id": gene, "value": value}
    arr = Mock()
    for binned in (True, False):
        tok = ScGPTTokenizer(arr, vocab_size=65000, n_expression_bins=51, max_genes=k)
        grouped = tok.window.apply(df, SLAF_LANCE_COO_SCHEMA, tok.max_genes, n_expression_bins=51, use_binned_expressions=binned)
        ids, mask, vals = tok.tokenize(frames.list_column(grouped, "gene_sequence"), frames.list_column(grouped, "expr_sequence"))
        mode = "scgpt_binned" if binned else "scgpt_raw"
        want = c_oracle.tokenize(cell, gene, value, mode, k, n_bins=51, vocab_size=65000)
        assert ids.is_cuda and ids.dtype == torch.int64 and mask.dtype == torch.bool and vals.dtype == torch.int64
        assert_same(ids.cpu().numpy(), want[0], f"{mode} ids")
        assert_same(mask.cpu().numpy(), want[1], f"{mode} mask")
        assert_same(vals.cpu().numpy(), want[2], f"{mode} vals")
        fused = engine.tokenize(cell, gene, value, **tok.hot_path_params(use_binned_expressions=binned))
        assert_same(fused.input_ids.cpu().numpy(), want[0], "fused ids")
        assert_same(fused.values.cpu().numpy(), want[2], "fused vals")
    tok = GeneformerTokenizer(arr, max_genes=k)
    grouped = tok.window.apply(df, SLAF_LANCE_COO_SCHEMA, tok.max_genes)
    ids, mask, vals = tok.tokenize(frames.list_column(grouped, "gene_sequence"))
    want = c_oracle.tokenize(cell, gene, value, "geneformer", k)
    assert vals is None
    assert_same(ids.cpu().numpy(), want[0], "gf ids")
    assert_same(mask.cpu().numpy(), want[1], "gf mask")


@pytest.mark.parametrize("n_window,n_tok", [(10, 51), (51, 10), (3, 1000), (1, 7)])
def test_window_bins_differ_from_tokenizer_bins(engine, n_window, n_tok):
    """PrefetchBatchProcessor's n_expression_bins (window) may differ from the tokenizer's
    (datasets.py:1025-1028 vs tokenizers.py:507-508): oracle = window(n_window) then tokenize_lists(n_tok)."""
    k = 24
    for value_dtype in ("uint16", "float32"):
        cell, gene, value = adversarial_batch(29, value_dtype, "uint16", "uint32", k=k, big=False)
        cids, offs, genes, exprs = c_oracle.window(cell, gene, value, "scgpt_binned", k, n_bins=n_window)
        if value_dtype == "float32":
            exprs = exprs.astype(np.float32).astype(np.float64)
        ids, mask, vals = c_oracle.tokenize_lists(True, offs, genes, exprs, False, k, n_bins=n_tok, vocab_size=65000)
        got = engine.tokenize(cell, gene, value, mode="scgpt_binned", max_genes=k, n_bins=n_tok, window_n_bins=n_window, vocab_size=65000)
        assert_same(got.cell_ids.cpu().numpy(), cids, "cell ids")
        assert_same(got.input_ids.cpu().numpy(), ids, "ids")
        assert_same(got.values.cpu().numpy(), vals, "vals")
        assert_same(got.attention_mask.cpu().numpy(), mask, "mask")


# ---------------------------------------------------------------- tokenizer interface (reference tests/test_tokenizers.py)
def _lists(offs, flat, cast=int):
    return [[cast(x) for x in flat[offs[i]:offs[i + 1]]] for i in range(len(offs) - 1)]


@pytest.mark.parametrize("tag", ["gf8", "gf64", "gf2048"])
def test_geneformer_tokenizer_reference_golden(tag):
    mg = int(GOLD[f"{tag}_max_genes"])
    tok = GeneformerTokenizer(Mock(), vocab_size=50000, max_genes=mg)
    ids, mask, vals = tok.tokenize(_lists(GOLD[f"{tag}_offsets"], GOLD[f"{tag}_genes"]))
    assert vals is None and tuple(ids.shape) == (24, mg)
    assert_same(ids.cpu().numpy(), GOLD[f"{tag}_ids"], "ids")
    assert_same(mask.cpu().numpy(), GOLD[f"{tag}_mask"], "mask")


@pytest.mark.parametrize("tag,is_int", [("sgi4", True), ("sgi1200", True), ("sgf_bins", False), ("sgf_bins10", False),
                                        ("sgf_raw", False), ("sgf_raw5", False)])
def test_scgpt_tokenizer_reference_golden(tag, is_int):
    mg, vocab, nb = (int(x) for x in GOLD[f"{tag}_cfg"])
    tok = ScGPTTokenizer(Mock(), vocab_size=vocab, n_expression_bins=nb, max_genes=mg)
    offs = GOLD[f"{tag}_offsets"]
    exprs = _lists(offs, GOLD[f"{tag}_exprs"], int if is_int else float)
    ids, mask, vals = tok.tokenize(_lists(offs, GOLD[f"{tag}_genes"]), exprs)
    assert tuple(ids.shape) == (len(offs) - 1, mg + 2)
    assert_same(ids.cpu().numpy(), GOLD[f"{tag}_ids"], "ids")
    assert_same(mask.cpu().numpy(), GOLD[f"{tag}_mask"], "mask")
    assert_same(vals.cpu().numpy(), GOLD[f"{tag}_vals"], "vals")


def test_tokenizer_contract_like_reference_tests():
    arr = Mock()
    # tests/test_tokenizers.py:66-82
    tok = GeneformerTokenizer(arr, vocab_size=1000)
    ids, mask, vals = tok.tokenize([[1, 2, 3], [4, 5, 6]])
    assert tuple(ids.shape) == (2, 2048) and ids.dtype == torch.int64 and mask.dtype == torch.bool and vals is None
    assert ids[0, 0].item() == 1 and ids[0, -1].item() == 0
    # :102-129 scGPT (2, 1026), CLS / SEP positions of the value stream are PAD
    st = ScGPTTokenizer(arr, vocab_size=1000, n_expression_bins=5)
    assert st.expr_bin_start == 1000 and st.expr_bin_size == 0.2           # :48-50
    ids, mask, vals = st.tokenize([[1, 2, 3], [4, 5, 6]], [[0.5, 0.8, 0.2], [0.9, 0.1, 0.7]])
    assert tuple(ids.shape) == (2, 1026) and tuple(vals.shape) == (2, 1026)
    assert vals[0, 0].item() == 0 and vals[0, 4].item() == 0 and ids[0, 4].item() == 2
    # :147-154,170-182 empty input raises, empty cell is fine
    with pytest.raises(ValueError, match="Gene sequences cannot be empty"):
        tok.tokenize([])
    with pytest.raises(ValueError, match="Gene sequences cannot be empty"):
        st.tokenize([], [])
    ids, mask, _ = GeneformerTokenizer(arr, max_genes=8).tokenize([[]])
    assert ids.cpu().tolist() == [[1, 2, 0, 0, 0, 0, 0, 0]]
    # :199-213 binning KAT
    s10 = ScGPTTokenizer(arr, vocab_size=1000, n_expression_bins=10)
    assert s10._expression_to_bin_vectorized(np.array([0.0, 0.1, 0.5, 0.9, -1.0])).tolist() == [0, 1001, 1005, 1009, 0]
    assert s10._expression_to_bin_vectorized(GOLD["binkat_x"]).tolist() == GOLD["binkat_tok"].tolist()
    # :229-245 +4 offset including ids >= vocab_size
    ids, _, _ = GeneformerTokenizer(arr, vocab_size=1000, max_genes=8).tokenize([[0, 999, 1000]])
    assert ids[0, :5].cpu().tolist() == [1, 4, 1003, 1004, 2]
    with pytest.raises(ValueError, match="max_genes must be >= 1"):
        GeneformerTokenizer(arr, max_genes=0)
    # mixed python int / float cells in one call take the branch per cell (tokenizers.py:441)
    ids, mask, vals = ScGPTTokenizer(arr, vocab_size=1000, n_expression_bins=51, max_genes=4).tokenize(
        [[5, 6, 7], [5, 6, 7], [1, 2, 3, 4, 5, 6]], [[3, 1, 9], [3.0, 0.0, 50.0], [1.0, 2.0, 0.0, 4.0, 5.0, 6.0]])
    assert vals.cpu().tolist() == [[0, 1003, 1001, 1009, 0, 0], [0, 1050, 0, 1050, 0, 0], [0, 1050, 1050, 0, 1050, 0]]
    assert ids.cpu().tolist()[2] == [1, 5, 6, 7, 8, 2]
    # explicit CPU placement is a copy of the finished tensors
    cpu_tok = GeneformerTokenizer(arr, max_genes=8, output_device="cpu")
    ids, mask, _ = cpu_tok.tokenize([[1, 2]])
    assert ids.device.type == "cpu" and ids.tolist() == [[1, 5, 6, 2, 0, 0, 0, 0]]


def test_create_window_factory():
    assert isinstance(create_window("geneformer"), GeneformerWindow)
    assert isinstance(create_window("scgpt"), ScGPTWindow)
    assert isinstance(create_window("simple"), SimpleWindow)
    with pytest.raises(ValueError):
        create_window("nope")


# ---------------------------------------------------------------- generic (Modal-side) pipeline: second caller of the path
def test_generic_batch_processor_and_format_batch(engine):
    """slaf/distributed/processor.py:105-301 + dataloader.py:357-475: boundary handling -> shuffle -> window ->
    tokenize -> per-group samples -> collated batch; fused route and three-call route agree with the oracle."""
    from oracle import py_restatement as pr
    from slaf_b200.ml.distributed import BatchProcessor, format_batch, tokenize_grouped
    from slaf_b200.ml.samplers import GenericShuffle

    rng = np.random.default_rng(4)
    lens = rng.integers(3, 40, 60)
    cell = np.repeat(np.arange(100, 160, dtype=np.uint32), lens)
    gene = np.concatenate([np.sort(rng.choice(5000, n, replace=False)) for n in lens]).astype(np.uint16)
    value = rng.geometric(0.3, cell.size).astype(np.uint16)
    cut = int(np.cumsum(lens)[29] + 2)                      # the batch boundary falls inside cell 130
    raw = [{"cell_integer_id": cell[:cut], "gene_integer_id": gene[:cut], "value": value[:cut]},
           {"cell_integer_id": cell[cut:], "gene_integer_id": gene[cut:], "value": value[cut:]}]

    class Custom(GenericShuffle):       # a user-supplied strategy object forces the three-call route
        pass

    for shuffle in (GenericShuffle(), Custom()):
        tok = ScGPTTokenizer(Mock(), vocab_size=6000, n_expression_bins=10, max_genes=16)
        proc = BatchProcessor(SLAF_LANCE_COO_SCHEMA, window=ScGPTWindow(), shuffle=shuffle, tokenizer=tokenize_grouped(tok),
                              max_items=16, seed=42, n_expression_bins=10, use_binned_expressions=False)
        s1 = proc.process_batch([raw[0]], epoch=0, partition_id=0)
        s2 = proc.process_batch([raw[1]], epoch=0, partition_id=0, is_partition_exhausted=True)
        assert [x["group_key"] for x in s1 + s2] != sorted(x["group_key"] for x in s1 + s2)      # shuffled
        assert sorted(x["group_key"] for x in s1) == list(range(100, 130))                      # cell 130 waits for its tail
        assert sorted(x["group_key"] for x in s2) == list(range(130, 160))
        # expected: complete rows of each call through the restated reference path, seed 42 + batch_id
        first = cell < 130
        for samples, rows, seed in ((s1, first, 42), (s2, ~first, 43)):
            ids, mask, vals, cids = pr.hot_path(cell[rows], gene[rows], value[rows], "scgpt", 16, seed=seed, vocab_size=6000,
                                                n_expression_bins=10, use_binned_expressions=False)
            batch = next(format_batch(samples))
            assert set(batch) == {"input_ids", "attention_mask", "cell_ids", "values"}
            assert_same(batch["cell_ids"].cpu().numpy(), cids, "cell ids")
            assert_same(batch["input_ids"].cpu().numpy(), ids, "ids")
            assert_same(batch["values"].cpu().numpy(), vals, "values")
            assert_same(batch["attention_mask"].cpu().numpy(), mask, "mask")
    # no tokenizer: grouped samples
    proc = BatchProcessor(SLAF_LANCE_COO_SCHEMA, window=GeneformerWindow(), shuffle=None, tokenizer=None, max_items=8)
    out = proc.process_batch(raw, is_partition_exhausted=True)
    assert len(out) == 60 and out[0]["group_key"] == 100 and set(out[0]["grouped"]) == {"cell_integer_id", "gene_sequence"}
    g = next(format_batch(out[:3]))
    assert g["group_keys"] == [100, 101, 102] and len(g["grouped"]) == 3
    assert list(format_batch([])) == []
