"""CPU tests of the host-side control plane of the drop-in layer: frame adapters, the block
shuffle (CPython-exact permutation applied to frames), schema objects, factories.  No kernel runs
here; the window / tokenizer arithmetic is covered by the -m gpu tests."""

import random

import numpy as np
import pyarrow as pa
import pytest

from oracle import py_restatement as pr
from slaf_b200.core.tabular_schema import SLAF_LANCE_COO_SCHEMA, DataSchema
from slaf_b200.engine import regroup_contiguous
from slaf_b200.ml import frames
from slaf_b200.ml.samplers import GenericShuffle, RandomShuffle, ShuffleType, StratifiedShuffle, create_shuffle


def _coo(seed=0, n_cells=37, split=False):
    rng = np.random.default_rng(seed)
    ids = rng.choice(10_000, n_cells, replace=False)
    lens = rng.integers(1, 9, n_cells)
    cell = np.repeat(ids, lens).astype(np.uint32)
    gene = rng.integers(0, 60000, cell.size).astype(np.uint16)
    value = rng.integers(1, 50, cell.size).astype(np.uint16)
    if split:  # move the tail of the first cell to the end (a cell in two separated runs)
        k = int(lens[0])
        order = np.r_[np.arange(0, max(1, k // 2)), np.arange(k, cell.size), np.arange(max(1, k // 2), k)]
        cell, gene, value = cell[order], gene[order], value[order]
    return {"cell_integer_id": cell, "gene_integer_id": gene, "value": value}


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("kind", ["dict", "arrow", "pandas"])
@pytest.mark.parametrize("seed", [42, 43, 10042])
def test_random_shuffle_matches_reference_semantics(seed, kind, split):
    cols = _coo(seed, split=split)
    df = cols if kind == "dict" else (pa.table(cols) if kind == "arrow" else __import__("pandas").DataFrame(cols))
    out = RandomShuffle().apply(df, seed)
    want = pr.apply_block_shuffle(cols["cell_integer_id"], cols["gene_integer_id"], cols["value"], seed)
    for name, w in zip(("cell_integer_id", "gene_integer_id", "value"), want):
        assert np.array_equal(frames.column(out, name), w), name


def test_random_shuffle_chunked_mode():
    cols = _coo(7, n_cells=23)
    chunks = RandomShuffle().apply(cols, 99, batch_size=5)
    assert len(chunks) == 5                                   # ceil(23 / 5)
    per = [len(np.unique(c["cell_integer_id"])) for c in chunks]
    assert per == [5, 5, 5, 5, 3]
    whole = RandomShuffle().apply(cols, 99)
    assert np.array_equal(np.concatenate([c["cell_integer_id"] for c in chunks]), whole["cell_integer_id"])
    # chunk order is the CPython shuffle of the first-appearance order
    ids = cols["cell_integer_id"][np.sort(np.unique(cols["cell_integer_id"], return_index=True)[1])]
    random.seed(99)
    order = list(ids)
    random.shuffle(order)
    first_seen = whole["cell_integer_id"][np.sort(np.unique(whole["cell_integer_id"], return_index=True)[1])]
    assert first_seen.tolist() == [int(x) for x in order]


def test_shuffle_empty_and_factories():
    empty = {"cell_integer_id": np.zeros(0, np.uint32), "gene_integer_id": np.zeros(0, np.uint16), "value": np.zeros(0, np.uint16)}
    assert RandomShuffle().apply(empty, 1) is empty
    assert RandomShuffle().apply(empty, 1, batch_size=4) == []
    assert isinstance(create_shuffle("random"), RandomShuffle)
    assert isinstance(create_shuffle(ShuffleType.STRATIFIED), StratifiedShuffle)
    with pytest.raises(ValueError, match="Unsupported shuffle type"):
        create_shuffle("sorted")
    # without a cell-type column the stratified strategy is the random one (samplers.py:140-142)
    cols = _coo(3)
    a = StratifiedShuffle().apply(cols, 5)
    b = RandomShuffle().apply(cols, 5)
    assert np.array_equal(a["cell_integer_id"], b["cell_integer_id"])


def test_generic_shuffle_uses_schema_group_key():
    cols = _coo(11)
    df = {"g": cols["cell_integer_id"], "i": cols["gene_integer_id"], "v": cols["value"]}
    schema = DataSchema(group_key="g", item_key="i", value_key="v")
    out = GenericShuffle().apply(df, schema, 77)
    want = pr.apply_block_shuffle(cols["cell_integer_id"], cols["gene_integer_id"], cols["value"], 77)
    assert np.array_equal(out["g"], want[0]) and np.array_equal(out["i"], want[1])


def test_regroup_contiguous_is_partition_by_order():
    cell = np.array([5, 5, 9, 9, 5, 7, 9], np.int32)
    gene = np.arange(7, dtype=np.int32)
    value = np.arange(7, dtype=np.uint16)
    c, g, v = regroup_contiguous(cell, gene, value)
    assert c.tolist() == [5, 5, 5, 9, 9, 9, 7]
    assert g.tolist() == [0, 1, 4, 2, 3, 6, 5]
    assert v.tolist() == g.tolist()


def test_schema_mirror():
    s = SLAF_LANCE_COO_SCHEMA
    assert (s.group_key, s.item_key, s.value_key, s.item_list_key, s.value_list_key) == (
        "cell_integer_id", "gene_integer_id", "value", "gene_sequence", "expr_sequence")
    d = DataSchema(group_key="a", item_key="b", value_key="c")
    assert d.item_list_key == "item_list" and d.value_list_key is None and d.group_key_out is None


def test_frame_extract_dtype_rules():
    base = _coo(5)
    # the storage dtypes pass through untouched (zero copy)
    c = frames.extract(base, SLAF_LANCE_COO_SCHEMA)
    assert c.cell.dtype == np.uint32 and c.gene.dtype == np.uint16 and c.value.dtype == np.uint16 and not c.value_is_float
    # wide integer ids narrow exactly
    wide = {"cell_integer_id": base["cell_integer_id"].astype(np.int64), "gene_integer_id": base["gene_integer_id"].astype(np.int64),
            "value": base["value"].astype(np.int64)}
    c = frames.extract(wide, SLAF_LANCE_COO_SCHEMA)
    assert c.cell.dtype == np.uint32 and c.gene.dtype == np.int32 and c.value.dtype == np.uint16
    assert c.cell_dtype == np.int64                                     # answers go back in the caller's dtype
    # float64 integral counts take the exact uint16 path but stay "float" for the bin dtype
    f = dict(base, value=base["value"].astype(np.float64))
    c = frames.extract(f, SLAF_LANCE_COO_SCHEMA)
    assert c.value.dtype == np.uint16 and c.value_is_float
    # non-representable float64 is refused rather than silently rounded
    bad = dict(base, value=np.full(base["value"].shape, 0.1, np.float64))
    with pytest.raises(TypeError):
        frames.extract(bad, SLAF_LANCE_COO_SCHEMA)
    with pytest.raises(ValueError, match="NaN"):
        frames.extract(dict(base, value=np.full(base["value"].shape, np.nan, np.float64)), SLAF_LANCE_COO_SCHEMA)
    with pytest.raises(ValueError):
        frames.extract(dict(base, value=np.full(base["value"].shape, 70000, np.int64)), SLAF_LANCE_COO_SCHEMA)
    with pytest.raises(TypeError):
        frames.frame_kind([1, 2, 3])


def test_build_grouped_arrow_and_dict():
    cols = frames.extract(_coo(2, n_cells=3), SLAF_LANCE_COO_SCHEMA)
    cell_ids = np.array([7, 8, 9], np.int64)
    offs = np.array([0, 2, 2, 5], np.int64)
    genes = np.arange(5, dtype=np.int64)
    exprs = np.array([1, 2, 3, 4, 5], np.float64)
    out = frames.build_grouped(cols, SLAF_LANCE_COO_SCHEMA, cell_ids, offs, genes, exprs, expr_is_bins=True)
    assert out.columns == ["cell_integer_id", "gene_sequence", "expr_sequence"] and len(out) == 3
    assert [list(x) for x in out["gene_sequence"]] == [[0, 1], [], [2, 3, 4]]
    assert out["expr_sequence"][2].dtype == np.float64                  # integer storage -> Float64 bins
    cols.kind = "arrow"
    t = frames.build_grouped(cols, SLAF_LANCE_COO_SCHEMA, cell_ids, offs, genes, None, expr_is_bins=False)
    assert t.column_names == ["cell_integer_id", "gene_sequence"]
    assert t.column("gene_sequence").to_pylist() == [[0, 1], [], [2, 3, 4]]
    assert frames.list_column(t, "gene_sequence") == [[0, 1], [], [2, 3, 4]]


# ---------------------------------------------------------------- host feeder (no GPU: raw mode)
def _raw_dataset(mode, n_cells=120, rows_per_file=611):
    from slaf_b200 import synthetic
    from slaf_b200.ml.sources import SyntheticSLAFArray

    b = synthetic.make_batch(n_cells, 500, 30, seed=3, min_nnz=2)
    return b, SyntheticSLAFArray.from_batch(b, 500, max_rows_per_file=rows_per_file)


@pytest.mark.parametrize("mode", ["mos", "fragment", "sequential"])
def test_raw_mode_boundary_reassembly_every_cell_once_with_all_rows(mode):
    """reference tests/test_cell_boundary_reassembly.py:33-73: despite fragments that cut cells, every
    cell comes out exactly once and with its full row count."""
    from slaf_b200.ml import SLAFDataLoader

    b, arr = _raw_dataset(mode)
    np.random.seed(3)
    loader = SLAFDataLoader(arr, raw_mode=True, batch_size=16, use_mixture_of_scanners=(mode == "mos"),
                            by_fragment=(mode != "sequential"), n_scanners=2, prefetch_batch_size=1000, verbose=False)
    assert loader.tokenizer is None and loader.special_tokens is None
    counts: dict[int, int] = {}
    for batch in loader:
        assert set(batch.keys()) == {"x", "cell_ids"}
        ids = np.asarray(batch["x"]["cell_integer_id"])
        assert len(batch["cell_ids"]) <= 16 and len(set(batch["cell_ids"])) == len(batch["cell_ids"])
        for c, n in zip(*np.unique(ids, return_counts=True)):
            assert int(c) not in counts, f"cell {c} emitted twice"
            counts[int(c)] = int(n)
    expected = np.diff(b.cell_start)
    assert sorted(counts) == list(range(b.n_cells))
    assert [counts[i] for i in range(b.n_cells)] == expected.tolist()


def test_processor_attributes_and_epoch_protocol():
    from slaf_b200.ml import PrefetchBatchProcessor, RandomShuffle, RawPrefetchBatch

    b, arr = _raw_dataset("fragment")
    proc = PrefetchBatchProcessor(arr, RandomShuffle(), None, raw_mode=True, n_epochs=2, use_mixture_of_scanners=False,
                                  by_fragment=True, batch_size=8, verbose=False, log_metrics=True)
    assert (proc.batch_id, proc.current_epoch, proc.n_epochs, proc.partial_cell_data) == (0, 0, 2, {})
    epochs = []
    with pytest.raises(StopIteration, match="No more epochs available"):
        while True:
            out = proc.load_prefetch_batch()
            assert isinstance(out, RawPrefetchBatch) and len(out.batch_dfs) == -(-len(out.cell_integer_ids) // 8)
            epochs.append(proc.current_epoch)
    assert sorted(set(epochs)) == [0, 1]
    m = proc.get_timing_metrics()
    assert set(m) == {"lance_loading", "window", "shuffle", "tokenize", "total", "cells_processed"} and m["cells_processed"] > 0
    with pytest.raises(ValueError, match="Invalid epoch -1"):
        proc.reset_for_epoch(-1)
    proc.reset_for_epoch(1)
    assert proc.batch_id == 0 and proc.current_epoch == 1 and proc.partial_cell_data == {}


def test_arrow_source_roundtrip(tmp_path):
    from slaf_b200.ml.sources import ArrowExpressionSource, open_expression, write_arrow_dataset

    b, _ = _raw_dataset("fragment")
    write_arrow_dataset(str(tmp_path / "expression.arrow"), b.cell, b.gene, b.value, max_rows_per_file=700, rows_per_batch=256)
    src = open_expression(str(tmp_path))
    assert isinstance(src, ArrowExpressionSource) and src.count_rows() == b.n_rows
    got = np.concatenate([c.cell for f in src.get_fragments() for c in f.to_batches(300)])
    assert np.array_equal(got, b.cell) and got.dtype == b.cell.dtype
    whole = np.concatenate([f.to_table().value for f in src.get_fragments()])
    assert np.array_equal(whole, b.value) and whole.dtype == np.uint16
    assert np.array_equal(np.concatenate([c.gene for c in src.to_batches(8192)]), b.gene)


def test_splice_partials_and_segment_regroup_equal_partition_by():
    """A boundary cell whose parts arrive through different scanners: the feeder's piece splice (views only)
    and the segment-level safety net both give what partition_by(maintain_order) + concat would."""
    from slaf_b200.ml.datasets import _regroup_segments, _runs, _splice_partials
    from slaf_b200.ml.sources import CooChunk, concat_chunks

    def chunk(ids, lens, base):
        cell = np.repeat(np.asarray(ids, np.uint32), lens)
        return CooChunk(cell, (np.arange(len(cell)) + base).astype(np.uint16), np.ones(len(cell), np.uint16))

    partials = {7: chunk([7], [3], 0), 40: chunk([40], [2], 100)}
    body = [chunk([20, 21, 40], [4, 5, 6], 200),        # ends with rows of cell 40 (its first part arrives last)
            chunk([7, 8, 9], [300, 2, 2], 1000),         # starts with the continuation of cell 7 (longer than the gallop window)
            chunk([30], [5], 3000)]
    naive = concat_chunks(list(partials.values()) + body)
    want = regroup_contiguous(naive.cell, naive.gene, naive.value)
    got = concat_chunks(_splice_partials(partials, body))
    for a, b in zip((got.cell, got.gene, got.value), want):
        assert np.array_equal(a, b)
    # the safety net on the un-spliced order, driven by the segment table K1 would report
    starts, ends, ids = _runs(naive.cell)
    seg_start = np.concatenate((starts, [len(naive)])).astype(np.uint32)
    pieces = _regroup_segments(list(partials.values()) + body, seg_start, ids)
    got2 = concat_chunks(pieces)
    for a, b in zip((got2.cell, got2.gene, got2.value), want):
        assert np.array_equal(a, b)
    assert all(p.cell.base is not None for p in pieces)           # views, not copies


def test_group_boundary_handler_local_merge():
    """slaf/distributed/boundary.py:66-153,285-357 (local part): the last group of a batch waits for the next
    batch of the same partition; continuity = same group or the next integer."""
    from slaf_b200.ml.distributed import GroupBoundaryHandler

    h = GroupBoundaryHandler(SLAF_LANCE_COO_SCHEMA, continuity_check="sequential")
    assert h.check_continuity(5, 5) and h.check_continuity(5, 6) and not h.check_continuity(5, 7)
    assert GroupBoundaryHandler(SLAF_LANCE_COO_SCHEMA, "ordered").check_continuity(5, 9)
    assert not GroupBoundaryHandler(SLAF_LANCE_COO_SCHEMA, "none").check_continuity(5, 6)
    mk = lambda ids: {"cell_integer_id": np.asarray(ids, np.uint32), "gene_integer_id": np.arange(len(ids), dtype=np.uint16),
                      "value": np.ones(len(ids), np.uint16)}
    complete, partial = h.merge_partial_data({}, mk([1, 1, 2, 2, 3]), partition_id=0)
    assert complete["cell_integer_id"].tolist() == [1, 1, 2, 2] and list(partial) == [3]
    complete, partial = h.merge_partial_data(partial, mk([3, 3, 4]), partition_id=0)
    assert complete["cell_integer_id"].tolist() == [3, 3, 3] and complete["gene_integer_id"].tolist() == [4, 0, 1] and list(partial) == [4]
    complete, partial = h.merge_partial_data(partial, mk([5, 6]), partition_id=0, is_partition_exhausted=True)
    assert complete["cell_integer_id"].tolist() == [4, 5, 6] and partial == {}
    with pytest.raises(NotImplementedError):
        GroupBoundaryHandler(SLAF_LANCE_COO_SCHEMA, partial_groups_kv={})


# ---------------------------------------------------------------- the reference's own boundary fixture
def _reference_boundary_fixture():
    """tests/conftest.py:478-481 of the reference: two fragments, cell 1 split across them, all values 1.0 (int32/int32/float32)."""
    from slaf_b200.ml.sources import CooChunk, ExpressionSource, Fragment, SyntheticSLAFArray

    frag1_cell, frag1_gene = [0, 0, 1, 1, 1], [0, 1, 0, 1, 2]
    frag2_cell, frag2_gene = [1, 1, 1, 2, 2, 3, 3, 3, 4, 4, 5], [3, 4, 5, 0, 1, 0, 1, 2, 0, 1, 0]

    class _Frag(Fragment):
        def __init__(self, c, g):
            self._t = CooChunk(np.asarray(c, np.int32), np.asarray(g, np.int32), np.ones(len(c), np.float32))

        def count_rows(self):
            return len(self._t)

        def to_table(self):
            return self._t

    class _Src(ExpressionSource):
        def get_fragments(self):
            return [_Frag(frag1_cell, frag1_gene), _Frag(frag2_cell, frag2_gene)]

    csi = np.array([0, 2, 8, 10, 13, 15, 16], np.int64)
    return SyntheticSLAFArray(_Src(), csi, 10), csi


@pytest.mark.parametrize("mos", [True, False])
def test_reference_cross_fragment_fixture_raw_epoch(mos):
    """tests/test_cell_boundary_reassembly.py:33-73 on the reference's fixture: every cell exactly once, observed rows per
    cell == _cell_start_index, no duplicate (cell, gene) pair."""
    from slaf_b200.ml import PrefetchBatchProcessor, RandomShuffle, RawPrefetchBatch

    arr, csi = _reference_boundary_fixture()
    np.random.seed(0)
    proc = PrefetchBatchProcessor(arr, RandomShuffle(), None, raw_mode=True, use_mixture_of_scanners=mos, by_fragment=True,
                                  n_scanners=2, prefetch_batch_size=1000, batch_size=4, verbose=False)
    batches = []
    with pytest.raises(StopIteration):
        while True:
            batches.append(proc.load_prefetch_batch())
    assert all(isinstance(b, RawPrefetchBatch) for b in batches)
    cells = np.concatenate([np.asarray(df["cell_integer_id"]) for b in batches for df in b.batch_dfs])
    genes = np.concatenate([np.asarray(df["gene_integer_id"]) for b in batches for df in b.batch_dfs])
    ids, counts = np.unique(cells, return_counts=True)
    assert ids.tolist() == list(range(6)) and counts.tolist() == np.diff(csi).tolist()
    assert len({(int(c), int(g)) for c, g in zip(cells, genes)}) == len(cells) == 16
    emitted = [c for b in batches for c in b.cell_integer_ids]
    assert sorted(emitted) == list(range(6))


def test_segments_from_cell_start_index_match_the_column():
    """The CSI-driven segment table (no look at the cell column) equals the runs of the column itself: whole fragments,
    consecutive pieces that continue a cell, unrelated pieces, cells without any row."""
    from slaf_b200 import synthetic
    from slaf_b200.engine import csi_segments
    from slaf_b200.ml.datasets import _runs
    from slaf_b200.ml.sources import ArrayExpressionSource

    def _segments_from_csi(pieces, csi):     # the library's host-only C routine (slafb_csi_segments), every head probed
        return csi_segments([(p.cell, p.gene, p.value) for p in pieces], [p.row_lo for p in pieces], csi, check=2)

    b = synthetic.make_batch(200, 500, 30, seed=3, min_nnz=2)
    # insert cells without rows: ids shift, CSI gets repeated offsets
    empty_after = {17, 18, 90}
    new_id, csi = [], [0]
    nid = 0
    for c in range(b.n_cells):
        new_id.append(nid)
        csi.append(int(b.cell_start[c + 1]))
        nid += 1
        if c in empty_after:
            csi.append(int(b.cell_start[c + 1]))
            nid += 1
    cell = np.asarray(new_id, np.uint32)[b.cell]
    csi = np.asarray(csi, np.int64)
    src = ArrayExpressionSource(cell, b.gene, b.value, 611)
    frs = [f.to_table() for f in src.get_fragments()]
    for pieces in ([frs[0]], [frs[2], frs[3]], [frs[4].slice(5, 300), frs[1]], frs, [frs[0].slice(0, 1)], list(frs[5].slice(0, 40).slice(i, i + 8) for i in range(0, 40, 8))):
        st, ids = _segments_from_csi(pieces, csi)
        cat = np.concatenate([p.cell for p in pieces])
        s0, _e0, i0 = _runs(cat)
        assert np.array_equal(st[:-1], s0) and st[-1] == len(cat) and np.array_equal(ids, i0)


# ---------------------------------------------------------------- property tests (hypothesis)
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 7), min_size=1, max_size=40), st.integers(1, 30), st.data())
def test_property_csi_segments_equal_column_runs(lens, rows_per_file, data):
    """For any table (cells of 0..7 rows), any fragment size and any selection of consecutive fragment slices, the segment
    table computed from _cell_start_index alone equals the runs of the cell column."""
    from slaf_b200.engine import csi_segments
    from slaf_b200.ml.datasets import _runs
    from slaf_b200.ml.sources import ArrayExpressionSource

    def _segments_from_csi(pieces, csi):
        return csi_segments([(p.cell, p.gene, p.value) for p in pieces], [p.row_lo for p in pieces], np.ascontiguousarray(csi, np.int64), check=2)

    lens = np.asarray(lens, np.int64)
    if lens.sum() == 0:
        lens[0] = 1
    csi = np.concatenate(([0], np.cumsum(lens)))
    cell = np.repeat(np.arange(len(lens), dtype=np.uint32), lens)
    src = ArrayExpressionSource(cell, np.zeros(len(cell), np.uint16), np.ones(len(cell), np.uint16), rows_per_file)
    frs = [f.to_table() for f in src.get_fragments()]
    k0 = data.draw(st.integers(0, len(frs) - 1))
    k1 = data.draw(st.integers(k0, len(frs) - 1))
    pieces = frs[k0:k1 + 1]
    cut = data.draw(st.integers(0, len(pieces[0]) - 1))
    pieces = [pieces[0].slice(cut, len(pieces[0]))] + pieces[1:]
    stt, ids = _segments_from_csi(pieces, csi)
    cat = np.concatenate([p.cell for p in pieces])
    s0, _e0, i0 = _runs(cat)
    assert np.array_equal(stt[:-1], s0) and stt[-1] == len(cat) and np.array_equal(ids, i0)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 5), st.integers(1, 6)), min_size=1, max_size=25), st.integers(0, 2**31 - 1))
def test_property_block_shuffle_keeps_cells_whole_and_matches_cpython(runs, seed):
    """RandomShuffle.apply on arbitrary (even non-contiguous) cell runs: every cell's rows stay together in arrival order,
    and the cell order is CPython's shuffle of the first-appearance order."""
    cell = np.concatenate([np.full(n, c, np.uint32) for c, n in runs])
    gene = np.arange(len(cell), dtype=np.uint16)
    out = RandomShuffle().apply({"cell_integer_id": cell, "gene_integer_id": gene, "value": gene}, seed)
    want = pr.apply_block_shuffle(cell, gene, gene, seed)
    assert np.array_equal(out["cell_integer_id"], want[0]) and np.array_equal(out["gene_integer_id"], want[1])
    oc = np.asarray(out["cell_integer_id"])
    assert len(np.flatnonzero(oc[1:] != oc[:-1])) + 1 == len(np.unique(cell))        # one run per cell
