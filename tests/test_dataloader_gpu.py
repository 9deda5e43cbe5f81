"""GPU tests of the dataloader mirror (slaf_b200.ml.datasets / dataloaders), modelled on the
reference's tests/test_dataloaders.py, tests/test_pytorch_datasets.py and
tests/test_cell_boundary_reassembly.py: shapes / dtypes / keys, every cell exactly once with all
of its rows despite fragment boundaries, and every emitted row bit-equal to the CPU oracle."""

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import py_restatement as pr
from slaf_b200 import synthetic
from slaf_b200.ml import (GeneformerTokenizer, PrefetchBatchProcessor, RandomShuffle, ScGPTTokenizer, SLAFDataLoader,
                          TokenizedPrefetchBatch)
from slaf_b200.ml.sources import ArrowExpressionSource, SyntheticSLAFArray, write_arrow_dataset
from tests.util import assert_same

pytestmark = pytest.mark.gpu

N_CELLS, N_GENES = 300, 2000


@pytest.fixture(scope="module")
def data():
    b = synthetic.make_batch(N_CELLS, N_GENES, 60, seed=5, min_nnz=4, first_cell_id=0)
    return b


def _array(b, rows_per_file=1777):
    return SyntheticSLAFArray.from_batch(b, N_GENES, max_rows_per_file=rows_per_file)


def _expected_rows(b, mode, **kw):
    ids, mask, vals, cids = c_oracle.tokenize(b.cell, b.gene, b.value, mode, **kw)
    return {int(c): (ids[i], mask[i], None if vals is None else vals[i]) for i, c in enumerate(cids)}


def _spanning_cells(b, rows_per_file):
    """cells whose rows cross a fragment boundary: under mixture-of-scanners their two parts can meet in
    either order (the reference's partition_by keeps arrival order too), so only their token multiset is fixed"""
    cuts = np.arange(rows_per_file, b.n_rows, rows_per_file)
    return {int(b.cell[c]) for c in cuts if b.cell[c] == b.cell[c - 1]}


def _check_rows(batches, expected, scgpt, loose=()):
    seen = []
    for batch in batches:
        ids, mask, cids = batch["input_ids"].cpu().numpy(), batch["attention_mask"].cpu().numpy(), batch["cell_ids"].cpu().numpy()
        vals = batch["values"].cpu().numpy() if scgpt else None
        for r, c in enumerate(cids):
            e = expected[int(c)]
            if int(c) in loose:
                assert np.array_equal(np.sort(ids[r]), np.sort(e[0])) and np.array_equal(mask[r], e[1]), f"spanning cell {c}"
                continue
            assert np.array_equal(ids[r], e[0]) and np.array_equal(mask[r], e[1]), f"cell {c}"
            if scgpt:
                assert np.array_equal(vals[r], e[2]), f"cell {c} values"
        seen.extend(int(c) for c in cids)
    return seen


@pytest.mark.parametrize("mode", ["mos", "fragment", "sequential"])
def test_geneformer_loader_every_cell_once_and_bit_exact(data, mode):
    kw = dict(use_mixture_of_scanners=(mode == "mos"), by_fragment=(mode != "sequential"))
    np.random.seed(0)
    loader = SLAFDataLoader(_array(data), tokenizer_type="geneformer", batch_size=32, max_genes=256, prefetch_batch_size=1000,
                            n_scanners=3, batches_per_chunk=2 if mode == "sequential" else 1, verbose=False, **kw)
    batches = list(loader)
    b0 = batches[0]
    assert set(b0.keys()) == {"input_ids", "attention_mask", "cell_ids"}                 # tests/test_dataloaders.py:76-83
    assert b0["input_ids"].dtype == torch.int64 and b0["attention_mask"].dtype == torch.bool and b0["cell_ids"].dtype == torch.int64
    assert b0["input_ids"].is_cuda and b0["input_ids"].shape[1] == 256
    assert all(b["input_ids"].shape[0] <= 32 for b in batches)
    for b in batches:                                                                    # :147-161 unique ids within a batch, in range
        ids = b["cell_ids"].cpu().tolist()
        assert len(set(ids)) == len(ids) and all(0 <= i < N_CELLS for i in ids)
    # max_genes=256 exceeds every cell's row count here, so no row is filtered and a spanning cell keeps its token multiset
    seen = _check_rows(batches, _expected_rows(data, "geneformer", max_genes=256), scgpt=False,
                       loose=_spanning_cells(data, 1777) if mode == "mos" else ())
    assert sorted(seen) == list(range(N_CELLS))                                          # test_cell_boundary_reassembly.py:33-73
    assert len(loader) == 0


@pytest.mark.parametrize("binned", [True, False])
def test_scgpt_loader_values_stream(data, binned):
    arr = _array(data)
    tok = ScGPTTokenizer(arr, vocab_size=5000, n_expression_bins=51, max_genes=48)
    proc = PrefetchBatchProcessor(arr, RandomShuffle(), tok, seed=42, n_expression_bins=51, use_binned_expressions=binned,
                                  use_mixture_of_scanners=False, by_fragment=True, verbose=False)
    assert proc.use_binned_expressions == binned and proc.batch_id == 0 and proc.current_epoch == 0
    out = []
    with pytest.raises(StopIteration, match="No more epochs available"):
        while True:
            out.append(proc.load_prefetch_batch())
    assert all(isinstance(o, TokenizedPrefetchBatch) for o in out)
    mode = "scgpt_binned" if binned else "scgpt_raw"
    exp = _expected_rows(data, mode, max_genes=48, n_bins=51, vocab_size=5000)
    seen = _check_rows([{"input_ids": o.input_ids, "attention_mask": o.attention_mask, "values": o.values, "cell_ids": o.cell_ids}
                        for o in out], exp, scgpt=True)
    assert sorted(seen) == list(range(N_CELLS))
    assert out[0].input_ids.shape[1] == 50 and out[0].values.shape == out[0].input_ids.shape       # L = max_genes + 2
    assert out[0].cell_integer_ids == out[0].cell_ids.cpu().tolist()
    proc.close()


def test_fragment_mode_matches_reference_batch_composition_and_order(data):
    """by_fragment, no MoS is deterministic in the reference: emulate datasets.py:812-953 + 1011-1098 with
    the numpy restatement and compare batch by batch, including the shuffled cell order."""
    arr = _array(data, rows_per_file=2500)
    tok = GeneformerTokenizer(arr, max_genes=32)
    proc = PrefetchBatchProcessor(arr, RandomShuffle(), tok, seed=42, n_epochs=2, use_mixture_of_scanners=False, by_fragment=True,
                                  verbose=False)
    got = []
    try:
        while True:
            got.append(proc.load_prefetch_batch())
    except StopIteration:
        pass
    want = []
    for epoch in range(2):
        batch_id, partial = 0, None
        frags = [(data.cell[lo:lo + 2500], data.gene[lo:lo + 2500], data.value[lo:lo + 2500]) for lo in range(0, data.n_rows, 2500)]
        for frag in frags + [None]:
            if frag is None:
                if partial is None:
                    break
                comb, complete_mask = partial, np.ones(len(partial[0]), bool)
            else:
                comb = frag if partial is None else tuple(np.concatenate([p, f]) for p, f in zip(partial, frag))
                complete_mask = comb[0] < comb[0].max()
            complete = tuple(c[complete_mask] for c in comb)
            partial = tuple(c[~complete_mask] for c in comb) if (~complete_mask).any() else None
            if len(complete[0]):
                want.append((epoch, batch_id, pr.hot_path(*complete, "geneformer", 32, seed=42 + batch_id + epoch * 10000)))
            batch_id += 1
    assert len(got) == len(want)
    for g, (epoch, batch_id, (ids, mask, _vals, cids)) in zip(got, want):
        assert (g.epoch, g.batch_id) == (epoch, batch_id)
        assert_same(g.cell_ids.cpu().numpy(), cids, "cell order")
        assert_same(g.input_ids.cpu().numpy(), ids, "ids")
        assert_same(g.attention_mask.cpu().numpy(), mask, "mask")
    proc.close()


def test_multi_epoch_and_cpu_placement(data):
    loader = SLAFDataLoader(_array(data), tokenizer_type="scgpt", batch_size=64, max_genes=16, n_expression_bins=10, n_epochs=3,
                            use_mixture_of_scanners=False, by_fragment=True, verbose=False, output_device="cpu")
    epochs, cells = [], {0: [], 1: [], 2: []}
    for b in loader:
        assert set(b.keys()) == {"input_ids", "attention_mask", "cell_ids", "values", "epoch"}
        assert b["input_ids"].device.type == "cpu" and b["input_ids"].shape[1] == 18               # :114-122
        epochs.append(b["epoch"])
        cells[b["epoch"]].extend(b["cell_ids"].tolist())
    assert sorted(set(epochs)) == [0, 1, 2] and epochs == sorted(epochs)
    for e in range(3):
        assert sorted(cells[e]) == list(range(N_CELLS))
    assert cells[0] != cells[1]                                                                   # seed + epoch * 10000


def test_arrow_directory_source_and_torch_dataloader(tmp_path, data):
    write_arrow_dataset(str(tmp_path / "expression.arrow"), data.cell, data.gene, data.value, max_rows_per_file=3001, rows_per_batch=700)
    src = ArrowExpressionSource(str(tmp_path / "expression.arrow"))
    assert src.count_rows() == data.n_rows and len(src.get_fragments()) == -(-data.n_rows // 3001)
    csi = np.asarray(data.cell_start)
    arr = SyntheticSLAFArray(src, csi, N_GENES, slaf_path=str(tmp_path))
    from slaf_b200.ml import SLAFIterableDataset

    np.random.seed(1)
    ds = SLAFIterableDataset(arr, GeneformerTokenizer(arr, max_genes=400), batch_size=50, n_scanners=2, prefetch_batch_size=1000, verbose=False)
    dl = torch.utils.data.DataLoader(ds, batch_size=None)                                        # tests/test_pytorch_datasets.py:430-469
    seen = _check_rows(list(dl), _expected_rows(data, "geneformer", max_genes=400), scgpt=False, loose=_spanning_cells(data, 3001))
    assert sorted(seen) == list(range(N_CELLS))


def test_loader_validation_errors(data):
    arr = _array(data)
    with pytest.raises(ValueError, match="n_scanners must be at least 1"):
        SLAFDataLoader(arr, n_scanners=0)
    with pytest.raises(ValueError, match="n_scanners cannot exceed 100"):
        SLAFDataLoader(arr, n_scanners=101)
    with pytest.raises(ValueError, match="prefetch_batch_size must be at least 1,000"):
        SLAFDataLoader(arr, prefetch_batch_size=10)
    with pytest.raises(ValueError, match="prefetch_batch_size cannot exceed 10,000,000"):
        SLAFDataLoader(arr, prefetch_batch_size=10_000_001)
    with pytest.raises(ValueError, match="is not supported"):
        SLAFDataLoader(arr, tokenizer_type="bert", use_mixture_of_scanners=False)
    proc = PrefetchBatchProcessor(arr, RandomShuffle(), None, use_mixture_of_scanners=False, verbose=False)
    with pytest.raises(RuntimeError, match="Tokenizer is required for tokenized mode"):
        proc.load_prefetch_batch()
    with pytest.raises(ValueError, match="Invalid epoch 5"):
        proc.reset_for_epoch(5)


def test_collect_csr_raw_mode_on_device(engine, data):
    """slafb_collect_csr: the batch as CSR in shuffled cell order (reference RawPrefetchBatch semantics: shuffle only)."""
    import random

    b = data
    slot = engine.submit(b.cell, b.gene, b.value)
    csr = engine.collect_csr(slot, shuffle_seed=91)
    random.seed(91)
    order = list(range(b.n_cells))
    random.shuffle(order)
    assert csr.cell_ids.cpu().tolist() == order and csr.n_cells == b.n_cells
    crow = csr.crow_indices.cpu().numpy()
    assert crow[0] == 0 and crow[-1] == b.n_rows and csr.col_indices.dtype == torch.int32 and csr.values.dtype == torch.uint16
    col, val = csr.col_indices.cpu().numpy(), csr.values.cpu().numpy()
    for j in (0, 1, 57, b.n_cells - 1):
        c = order[j]
        lo, hi = b.cell_start[c], b.cell_start[c + 1]
        assert np.array_equal(col[crow[j]:crow[j + 1]], b.gene[lo:hi].astype(np.int32))
        assert np.array_equal(val[crow[j]:crow[j + 1]], b.value[lo:hi])
    dense = csr.to_sparse_csr(N_GENES).to_dense()
    assert dense.shape == (b.n_cells, N_GENES) and float(dense.sum()) == float(b.value.sum())
    # a selection emits only those cells
    slot = engine.submit(b.cell, b.gene, b.value)
    sub = engine.collect_csr(slot, perm=np.array([5, 0, 7], np.int32))
    assert sub.cell_ids.cpu().tolist() == [5, 0, 7] and int(sub.crow_indices[-1]) == sum(int(b.cell_start[c + 1] - b.cell_start[c]) for c in (5, 0, 7))


@pytest.mark.parametrize("mode", ["mos", "fragment"])
def test_raw_mode_on_device_loader(data, mode):
    np.random.seed(2)
    loader = SLAFDataLoader(_array(data), raw_mode=True, device_raw=True, batch_size=24, use_mixture_of_scanners=(mode == "mos"),
                            by_fragment=True, n_scanners=3, prefetch_batch_size=1000, verbose=False)
    seen = {}
    for batch in loader:
        x = batch["x"]
        crow = x["crow_indices"].cpu().numpy()
        ids = batch["cell_ids"].cpu().tolist()
        assert x["col_indices"].is_cuda and crow[0] == 0 and len(crow) == len(ids) + 1 and len(ids) <= 24
        assert crow[-1] == x["col_indices"].numel() == x["values"].numel()
        col, val = x["col_indices"].cpu().numpy(), x["values"].cpu().numpy()
        for r, c in enumerate(ids):
            assert c not in seen
            seen[c] = (np.sort(col[crow[r]:crow[r + 1]]), int(val[crow[r]:crow[r + 1]].sum()))
    assert sorted(seen) == list(range(N_CELLS))
    for c in (0, 33, N_CELLS - 1):
        lo, hi = data.cell_start[c], data.cell_start[c + 1]
        assert np.array_equal(seen[c][0], np.sort(data.gene[lo:hi]).astype(np.int32)) and seen[c][1] == int(data.value[lo:hi].sum())


def test_reference_cross_fragment_fixture_tokenized():
    """The reference's own boundary fixture (tests/conftest.py:478-481: cell 1 split across two fragments, all values 1.0)
    through the tokenized loader in both loading modes: every cell once, tokens = all of the cell's genes in row order."""
    from tests.test_host_logic import _reference_boundary_fixture

    want = {0: [0, 1], 1: [0, 1, 2, 3, 4, 5], 2: [0, 1], 3: [0, 1, 2], 4: [0, 1], 5: [0]}
    for mos in (True, False):
        arr, _ = _reference_boundary_fixture()
        np.random.seed(1)
        loader = SLAFDataLoader(arr, tokenizer_type="scgpt", batch_size=4, max_genes=8, n_expression_bins=10, vocab_size=100,
                                use_mixture_of_scanners=mos, by_fragment=True, n_scanners=2, prefetch_batch_size=1000, verbose=False)
        seen = {}
        for batch in loader:
            for row, c in zip(batch["input_ids"].cpu().tolist(), batch["cell_ids"].cpu().tolist()):
                assert c not in seen
                seen[c] = row
        assert sorted(seen) == list(range(6))
        for c, genes in want.items():
            assert seen[c] == [1] + [g + 4 for g in genes] + [2] + [0] * (10 - len(genes) - 2), (mos, c)


def test_reference_tiny_adata_matrix_all_tokenizers(engine):
    """The reference's `tiny_adata` fixture matrix (tests/conftest.py:298-324: seed 42, 100 x 50, density 0.3, uniform(1, 5)
    float32, every row and column non-empty), in converter row order, through every mode against the oracle."""
    from scipy.sparse import csr_matrix

    np.random.seed(42)
    n_cells, n_genes = 100, 50
    n_nonzero = int(n_cells * n_genes * 0.3)
    data = np.random.uniform(1.0, 5.0, n_nonzero).astype(np.float32)
    rows = np.random.randint(0, n_cells, n_nonzero)
    cols = np.random.randint(0, n_genes, n_nonzero)
    X = csr_matrix((data, (rows, cols)), shape=(n_cells, n_genes)).tolil()
    for i in range(n_cells):
        if X[i, :].nnz == 0:
            X[i, np.random.randint(0, n_genes)] = np.random.uniform(1.0, 5.0)
    for j in range(n_genes):
        if X[:, j].nnz == 0:
            X[np.random.randint(0, n_cells), j] = np.random.uniform(1.0, 5.0)
    coo = X.tocsr().tocoo()                                        # cell-major, gene-ascending: the converter's row order
    cell, gene, value = coo.row.astype(np.int32), coo.col.astype(np.int32), coo.data.astype(np.float32)
    assert len(np.unique(cell)) == 100
    for mode, pct in (("geneformer", None), ("geneformer_pct", 25.0), ("scgpt_raw", None), ("scgpt_binned", None)):
        for k in (3, 10, 64):
            w = c_oracle.tokenize(cell, gene, value, mode, k, n_bins=10, vocab_size=50000, min_percentile=pct)
            g = engine.tokenize(cell, gene, value, mode=mode, max_genes=k, n_bins=10, vocab_size=50000, min_percentile=pct)
            assert_same(g.input_ids.cpu().numpy(), w[0], f"{mode} k={k} ids")
            assert_same(g.attention_mask.cpu().numpy(), w[1], f"{mode} k={k} mask")
            if w[2] is not None:
                assert_same(g.values.cpu().numpy(), w[2], f"{mode} k={k} values")


@pytest.mark.parametrize("mode", ["mos", "fragment", "sequential"])
def test_loader_with_cell_start_index_gives_identical_batches(data, mode):
    """use_cell_start_index=True: cell boundaries from _cell_start_index, the cell column never reaches the GPU;
    batch composition, order and every token are the same as on the default path."""
    kw = dict(use_mixture_of_scanners=(mode == "mos"), by_fragment=(mode != "sequential"), tokenizer_type="scgpt", batch_size=40,
              max_genes=48, n_expression_bins=51, vocab_size=5000, prefetch_batch_size=1000, n_scanners=3, verbose=False)
    runs = []
    for flag in (False, True):
        np.random.seed(11)                              # same mixture-of-scanners draws
        loader = SLAFDataLoader(_array(data), use_cell_start_index=flag, **kw)
        runs.append([{k: v.cpu().numpy() for k, v in b.items()} for b in loader])
        if flag:
            t = loader._dataset.batch_processor._engine.timing()
            assert t["csr_ms"] == 0.0                     # no CSR build ran
    assert len(runs[0]) == len(runs[1])
    for a, b in zip(*runs):
        for k in ("cell_ids", "input_ids", "values", "attention_mask"):
            assert np.array_equal(a[k], b[k]), k
    assert sorted(int(c) for b in runs[1] for c in b["cell_ids"]) == list(range(N_CELLS))


def test_cell_start_index_mismatch_is_detected(data):
    arr = _array(data)
    arr._cell_start_index = arr._cell_start_index.copy()
    arr._cell_start_index[1:] += 3                         # wrong offsets: the boundary probes must notice
    tok = GeneformerTokenizer(arr, max_genes=32)
    proc = PrefetchBatchProcessor(arr, RandomShuffle(), tok, use_mixture_of_scanners=False, by_fragment=True, verbose=False,
                                  use_cell_start_index=True)
    with pytest.raises(ValueError, match="_cell_start_index does not describe"):
        proc.load_prefetch_batch()
    proc.close()


def _emulate_fragment_mode(cell, gene, value, rows_per_file, n_epochs, max_genes, seed=42):
    """datasets.py:812-953 + 1011-1098 for by_fragment without mixture-of-scanners, with the numpy restatement: the rows with the
    largest cell id of [partial + fragment] wait for the next load, the rest is tokenized in first-appearance order, shuffled."""
    want = []
    for epoch in range(n_epochs):
        batch_id, partial = 0, None
        frags = [(cell[lo:lo + rows_per_file], gene[lo:lo + rows_per_file], value[lo:lo + rows_per_file]) for lo in range(0, len(cell), rows_per_file)]
        for frag in frags + [None]:
            if frag is None:
                if partial is None:
                    break
                comb, complete_mask = partial, np.ones(len(partial[0]), bool)
            else:
                comb = frag if partial is None else tuple(np.concatenate([p, f]) for p, f in zip(partial, frag))
                complete_mask = comb[0] < comb[0].max()
            complete = tuple(c[complete_mask] for c in comb)
            partial = tuple(c[~complete_mask] for c in comb) if (~complete_mask).any() else None
            if len(complete[0]):
                want.append((epoch, batch_id, pr.hot_path(*complete, "geneformer", max_genes, seed=seed + batch_id + epoch * 10000)))
            batch_id += 1
    return want


def _drain(proc):
    got = []
    try:
        while True:
            got.append(proc.load_prefetch_batch())
    except StopIteration:
        pass
    return got


def test_look_ahead_falls_back_when_the_table_is_not_sorted_by_cell(data):
    """The look-ahead predicts the waiting cell as the last run of the load; with cell ids that are NOT ascending the cell with
    the largest id sits somewhere else, the segment table contradicts the prediction and the next load is rebuilt from the
    table -- batch composition, order, ids and epochs stay those of the reference's rule."""
    rng = np.random.default_rng(3)
    relabel = rng.permutation(N_CELLS).astype(data.cell.dtype)
    cell = relabel[data.cell]                                   # runs stay contiguous, ids are shuffled
    from slaf_b200.ml.sources import ArrayExpressionSource

    arr = SyntheticSLAFArray(ArrayExpressionSource(cell, data.gene, data.value, 2500), np.zeros(N_CELLS + 1, np.int64), N_GENES)
    tok = GeneformerTokenizer(arr, max_genes=32)
    proc = PrefetchBatchProcessor(arr, RandomShuffle(), tok, seed=42, n_epochs=2, use_mixture_of_scanners=False, by_fragment=True, verbose=False, use_cell_start_index=False)
    assert proc._lookahead_ok()
    got = _drain(proc)
    want = _emulate_fragment_mode(cell, data.gene, data.value, 2500, 2, 32)
    assert len(got) == len(want) and len(got) > 4
    for g, (epoch, batch_id, (ids, mask, _vals, cids)) in zip(got, want):
        assert (g.epoch, g.batch_id) == (epoch, batch_id)
        assert_same(g.cell_ids.cpu().numpy(), cids, "cell order")
        assert_same(g.input_ids.cpu().numpy(), ids, "ids")
        assert_same(g.attention_mask.cpu().numpy(), mask, "mask")
    proc.close()


def test_look_ahead_defers_the_submit_when_the_engine_has_to_grow(data):
    from slaf_b200.engine import HotPathEngine

    arr = _array(data, rows_per_file=2500)
    tok = GeneformerTokenizer(arr, max_genes=32)
    proc = PrefetchBatchProcessor(arr, RandomShuffle(), tok, seed=42, use_mixture_of_scanners=False, by_fragment=True, verbose=False, use_cell_start_index=False)
    proc._engine = small = HotPathEngine(0, max_rows=2510, max_cells=2510)     # the first load fits, partial + second fragment does not
    got = _drain(proc)
    assert proc._engine is not small and proc._engine.max_rows > 2510
    want = _emulate_fragment_mode(data.cell, data.gene, data.value, 2500, 1, 32)
    assert len(got) == len(want)
    for g, (_epoch, batch_id, (ids, _mask, _vals, cids)) in zip(got, want):
        assert g.batch_id == batch_id
        assert_same(g.cell_ids.cpu().numpy(), cids, "cell order")
        assert_same(g.input_ids.cpu().numpy(), ids, "ids")
    assert [sorted(g.partial_cell_data) for g in got[:-1]] == [[int(g.cell_ids.max().item()) + 1] for g in got[:-1]]   # the waiting cell
    assert got[-1].partial_cell_data == {}
    proc.close()


def test_look_ahead_reset_for_epoch_drops_the_submitted_load(data):
    arr = _array(data, rows_per_file=2500)
    tok = GeneformerTokenizer(arr, max_genes=32)
    proc = PrefetchBatchProcessor(arr, RandomShuffle(), tok, seed=42, n_epochs=2, use_mixture_of_scanners=False, by_fragment=True, verbose=False, use_cell_start_index=False)
    first = proc.load_prefetch_batch()
    second = proc.load_prefetch_batch()
    assert proc._ahead is not None and (first.batch_id, second.batch_id) == (0, 1)
    proc.reset_for_epoch(1)
    assert proc._ahead is None and proc.batch_id == 0
    again = _drain(proc)
    want = [w for w in _emulate_fragment_mode(data.cell, data.gene, data.value, 2500, 2, 32) if w[0] == 1]
    assert len(again) == len(want)
    for g, (epoch, batch_id, (ids, _mask, _vals, cids)) in zip(again, want):
        assert (g.epoch, g.batch_id) == (epoch, batch_id)
        assert_same(g.cell_ids.cpu().numpy(), cids, "cell order")
        assert_same(g.input_ids.cpu().numpy(), ids, "ids")
    proc.close()


def _stream(proc):
    return [(g.epoch, g.batch_id, g.cell_ids.cpu().numpy(), g.input_ids.cpu().numpy(), sorted(g.partial_cell_data)) for g in _drain(proc)]


@pytest.mark.parametrize("case", ["consistent", "index_overstates_one_cell"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_mixture_of_scanners_look_ahead_equals_the_synchronous_feeder(data, seed, case, monkeypatch):
    """Under mixture-of-scanners the look-ahead predicts the incomplete cells from the runs at the piece boundaries and
    _cell_start_index.  With the same scanner draws the stream of prefetch batches (ids, order, tokens, pending cells) must
    be the one the synchronous feeder produces -- also when the prediction is wrong because the index claims more rows for
    a cell in the middle of a chunk than the table has (that cell then waits until the final flush in both)."""
    def run(lookahead):
        arr = _array(data, rows_per_file=1777)
        if case == "index_overstates_one_cell":
            # cell 150 lies inside a chunk: claim 5 more rows for it without moving any other cell's count
            csi = arr._cell_start_index.copy()
            counts = np.diff(csi)
            counts[150] += 5
            arr._cell_start_index = np.concatenate(([0], np.cumsum(counts)))
        tok = GeneformerTokenizer(arr, max_genes=32)
        proc = PrefetchBatchProcessor(arr, RandomShuffle(), tok, seed=42, n_epochs=2, use_mixture_of_scanners=True, n_scanners=3,
                                      prefetch_batch_size=1000, verbose=False, use_cell_start_index=False)
        if not lookahead:
            monkeypatch.setattr(proc, "_lookahead_ok", lambda: False)
        else:
            assert proc._lookahead_ok()
        np.random.seed(seed)
        out = _stream(proc)
        proc.close()
        return out

    a, b = run(True), run(False)
    assert len(a) == len(b) and len(a) > 6
    for x, y in zip(a, b):
        assert x[:2] == y[:2] and x[4] == y[4]
        assert np.array_equal(x[2], y[2]) and np.array_equal(x[3], y[3])
    for epoch in (0, 1):
        seen = np.concatenate([x[2] for x in a if x[0] == epoch])
        assert sorted(seen.tolist()) == list(range(N_CELLS))


@pytest.mark.parametrize("csi_route", [None, False], ids=["cell_start_index", "cell_column"])
def test_arrow_source_page_locked_maps_feed_the_copy_engine_directly(tmp_path, data, csi_route):
    """ArrowExpressionSource.pin(): the files' read-only mappings are registered with the driver, the record-batch views are
    DMA sources and the feeder's pinned staging ring is never allocated; tokens are the same either way.  (A driver that refuses
    to register file mappings leaves the staged path in place: pin() then returns False and only the equality is checked.)"""
    write_arrow_dataset(str(tmp_path / "expression.arrow"), data.cell, data.gene, data.value, max_rows_per_file=3001, rows_per_batch=700)
    expected = _expected_rows(data, "geneformer", max_genes=400)
    for pin in (False, True):
        src = ArrowExpressionSource(str(tmp_path / "expression.arrow"))
        pinned = src.pin() if pin else False
        arr = SyntheticSLAFArray(src, np.asarray(data.cell_start), N_GENES)
        proc = PrefetchBatchProcessor(arr, RandomShuffle(), GeneformerTokenizer(arr, max_genes=400), seed=42, use_mixture_of_scanners=False,
                                      by_fragment=True, verbose=False, use_cell_start_index=csi_route)
        proc.SMALL_PIECE_ROWS = 650                    # the 700-row record batches of this small table must be page-locked or staged; pending cells stay small
        out = _drain(proc)
        seen = _check_rows([{"input_ids": o.input_ids, "attention_mask": o.attention_mask, "cell_ids": o.cell_ids} for o in out], expected, scgpt=False)
        assert sorted(seen) == list(range(N_CELLS))
        if pinned:
            assert proc._staging.capacity == 0, "page-locked views must not go through the staging ring"
        elif not pin:
            assert proc._staging.capacity > 0
        torch.cuda.synchronize()
        proc.close()
        src.unpin()
        print("arrow mappings page-locked:", pinned)
