"""CPU tests of the host feeder's control flow (slaf_b200/ml/datasets.py) with a stand-in engine: which rows are submitted when,
which cells are emitted in which load, what waits in partial_cell_data -- the look-ahead must produce exactly the stream of the
synchronous feeder, which in turn is the reference's loop (slaf/ml/datasets.py:812-953, 1011-1098).  The stand-in keeps the
submitted cell columns, answers ``segments`` with their runs and ``collect`` with the selected cell ids; the tokens themselves are
the GPU tests' business (tests/test_dataloader_gpu.py runs the same comparisons through the real library)."""

import numpy as np
import pytest
import torch

from slaf_b200 import synthetic
from slaf_b200.ml import GeneformerTokenizer, PrefetchBatchProcessor, RandomShuffle, SLAFIterableDataset
from slaf_b200.ml import datasets as D
from slaf_b200.ml.sources import ArrayExpressionSource, SyntheticSLAFArray

N_CELLS, N_GENES = 400, 3000


def _segments_from_csi(pieces: list, csi: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Segment table of the concatenated pieces from SLAFArray._cell_start_index alone (no look at the cell column):
    piece p covers table rows [row_lo, row_lo + len); cell c owns rows [csi[c], csi[c + 1]).  Consecutive pieces that
    continue the same cell give ONE segment (their rows are adjacent in the staging columns); a cell that shows up in
    two separated places gives two segments with the same id, which the caller's splice safety net handles.
    Returns (seg_start uint32 [C + 1], seg_cell int64 [C]).  O(cells), vectorised."""
    starts, ids = [], []
    pos = 0
    last_cell = None
    for p in pieces:
        n = len(p)
        if n == 0:
            continue
        lo, hi = int(p.row_lo), int(p.row_lo) + n
        c0 = int(np.searchsorted(csi, lo, side="right")) - 1
        c1 = int(np.searchsorted(csi, hi - 1, side="right")) - 1
        cells = np.arange(c0, c1 + 1, dtype=np.int64)
        first_row = np.maximum(csi[cells], lo)
        nonempty = csi[cells + 1] > first_row                      # cells without rows (or not reaching into this piece) have no segment
        cells, first_row = cells[nonempty], first_row[nonempty]
        st = first_row - lo + pos
        if len(cells):
            new_last = int(cells[-1])
            if last_cell is not None and last_cell == int(cells[0]) and st[0] == pos:
                cells, st = cells[1:], st[1:]                        # the previous piece's last cell continues here: one segment
            last_cell = new_last
        starts.append(st)
        ids.append(cells)
        pos += n
    ids = [i for i in ids if len(i)]
    starts = [s_ for s_ in starts if len(s_)]
    seg_cell = np.concatenate(ids) if ids else np.zeros(0, np.int64)
    seg_start = np.concatenate(starts + [np.array([pos], np.int64)]) if starts else np.array([0], np.int64)
    return seg_start.astype(np.uint32), seg_cell


class _Out:
    def __init__(self, cids, width):
        n = len(cids)
        self.input_ids = torch.zeros((n, width), dtype=torch.int64)
        self.attention_mask = torch.zeros((n, width), dtype=torch.bool)
        self.values = None
        self.cell_ids = torch.from_numpy(np.asarray(cids, np.int64))
        self.n_cells = n


class StandInEngine:
    """submit / segments / collect / release of HotPathEngine on numpy; eight slots like the library, and a log of the calls."""

    device = torch.device("cpu")

    def __init__(self, max_rows, max_cells):
        self.max_rows, self.max_cells, self.slots, self.next, self.log = max_rows, max_cells, {}, 0, []

    def _put(self, cell):
        assert len(cell) <= self.max_rows, "submitted more rows than the engine was sized for"
        assert len(self.slots) < 8, "more than eight batches in flight"
        slot, self.next = self.next, self.next + 1
        self.slots[slot] = np.array(cell)            # a copy: the feeder may recycle its buffers once submit has returned
        self.log.append(("submit", slot))
        return slot

    def submit_pieces(self, pieces):
        return self._put(np.concatenate([np.asarray(p[0]) for p in pieces]))

    def submit(self, cell, gene, value, shuffle_seed=None):
        return self._put(np.asarray(cell))

    def submit_segments(self, pieces, seg_start, seg_cell, cell_signed=False, shuffle_seed=None):
        # the cell-start-index route: no cell column; the table comes from the host and must cover the rows handed over
        n_rows = sum(len(np.asarray(g)) for g, _v in pieces)
        seg_start, seg_cell = np.asarray(seg_start), np.asarray(seg_cell)
        assert all(len(np.asarray(g)) == len(np.asarray(v)) for g, v in pieces)
        assert seg_start[0] == 0 and seg_start[-1] == n_rows and len(seg_start) == len(seg_cell) + 1 and np.all(np.diff(seg_start.astype(np.int64)) > 0)
        return self._put(np.repeat(seg_cell, np.diff(seg_start.astype(np.int64))))

    def submit_csi(self, pieces, row_lo, csi, check=1, shuffle_seed=None):
        # the cell-start-index route: the library's host-only C routine derives (and probes) the table; checked here against the
        # numpy restatement below
        from slaf_b200.engine import csi_segments
        from slaf_b200.ml.sources import CooChunk

        seg_start, seg_cell = csi_segments(pieces, row_lo, csi, check=check)
        want = _segments_from_csi([CooChunk(c, g, v, lo) for (c, g, v), lo in zip(pieces, row_lo)], csi)
        assert np.array_equal(seg_start, want[0]) and np.array_equal(seg_cell, want[1].astype(np.uint32))
        return self.submit_segments([(g, v) for _c, g, v in pieces], seg_start, seg_cell), len(seg_cell)

    def wait_loaded(self, slot):
        pass

    def segments(self, slot):
        cell = self.slots[slot]
        starts, _ends, ids = D._runs(cell)
        self.log.append(("segments", slot))
        return np.concatenate((starts, [len(cell)])).astype(np.uint32), ids.copy()

    def collect(self, slot, perm=None, capacity_cells=None, **kw):
        ids = D._runs(self.slots.pop(slot))[2]
        self.log.append(("collect", slot))
        return _Out(ids[np.asarray(perm)], kw["max_genes"])

    def release(self, slot):
        self.slots.pop(slot, None)
        self.log.append(("release", slot))

    def close(self):
        pass


@pytest.fixture()
def stand_in(monkeypatch):
    engines = []

    def engine_for(self, rows, all_cells=False):
        if self._engine is None or self._engine.max_rows < rows:
            assert self._engine is None or not self._engine.slots, "engine re-created with batches in flight"
            self._engine = StandInEngine(rows + rows // 4, rows)
            engines.append(self._engine)
        return self._engine

    class _Event:
        def record(self, *a):
            pass

        def synchronize(self):
            pass

    monkeypatch.setattr(D.PrefetchBatchProcessor, "_engine_for", engine_for)
    monkeypatch.setattr(D, "pinned_empty", lambda n, dt=np.uint8: np.empty(n, dt))
    monkeypatch.setattr(D, "is_page_locked", lambda a: True)           # "direct" route: pieces are handed over as they are
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)      # also on a GPU box these tests stay on the stand-in
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: None)
    return engines


def _data(seed=5):
    return synthetic.make_batch(N_CELLS, N_GENES, 50, seed=seed, min_nnz=3, first_cell_id=0)


def _proc(b, cell=None, csi=None, rows_per_file=1777, **kw):
    cell = b.cell if cell is None else cell
    arr = SyntheticSLAFArray(ArrayExpressionSource(cell, b.gene, b.value, rows_per_file),
                             b.cell_start if csi is None else csi, N_GENES)
    tok = GeneformerTokenizer(arr, max_genes=16)
    kw.setdefault("use_cell_start_index", False)      # most tests here are about the cell-column route and its look-ahead
    return PrefetchBatchProcessor(arr, RandomShuffle(), tok, seed=42, verbose=False, **kw)


def _drain(proc):
    out = []
    try:
        while True:
            g = proc.load_prefetch_batch()
            out.append((g.epoch, g.batch_id, g.cell_ids.numpy().copy(), sorted(g.partial_cell_data)))
    except StopIteration:
        pass
    return out


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x[:2] == y[:2] and x[3] == y[3] and np.array_equal(x[2], y[2])


@pytest.mark.parametrize("mode", ["fragment", "sequential", "mos"])
@pytest.mark.parametrize("seed", [0, 1])
def test_look_ahead_stream_equals_synchronous_stream(stand_in, monkeypatch, mode, seed):
    b = _data()
    kw = dict(use_mixture_of_scanners=(mode == "mos"), by_fragment=(mode != "sequential"), n_scanners=3, prefetch_batch_size=1000,
              batches_per_chunk=2 if mode == "sequential" else 1, n_epochs=2)
    runs = []
    for lookahead in (True, False):
        proc = _proc(b, **kw)
        if lookahead:
            assert proc._lookahead_ok()
        else:
            monkeypatch.setattr(proc, "_lookahead_ok", lambda: False)
        np.random.seed(seed)
        runs.append(_drain(proc))
        log = proc._engine.log
        if lookahead:
            # the point of it: a later load's rows are submitted before the current load's table is asked for
            first_seg = next(i for i, e in enumerate(log) if e[0] == "segments")
            assert sum(1 for e in log[:first_seg] if e[0] == "submit") == 2
        assert not proc._engine.slots                                  # nothing left in flight at the end
    _same(*runs)
    for epoch in (0, 1):
        seen = np.concatenate([x[2] for x in runs[0] if x[0] == epoch])
        assert sorted(seen.tolist()) == list(range(N_CELLS))


def test_fragment_stream_is_the_reference_rule(stand_in):
    """by_fragment without mixture-of-scanners: the cell with the largest id of [pending + fragment] waits (datasets.py:931-953)."""
    b = _data()
    got = _drain(_proc(b, rows_per_file=2500, use_mixture_of_scanners=False, by_fragment=True))
    import random

    pending = np.zeros(0, b.cell.dtype)
    want, batch_id = [], 0
    for lo in list(range(0, b.n_rows, 2500)) + [None]:
        comb = pending if lo is None else np.concatenate([pending, b.cell[lo:lo + 2500]])
        if lo is None and not len(pending):
            break
        keep = np.ones(len(comb), bool) if lo is None else comb < comb.max()
        ids = comb[keep][np.sort(np.unique(comb[keep], return_index=True)[1])]
        pending = comb[~keep]
        if len(ids):
            random.seed(42 + batch_id)
            order = list(range(len(ids)))
            random.shuffle(order)
            want.append((batch_id, ids[order], sorted(set(pending.tolist()))))
        batch_id += 1
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g[1] == w[0] and np.array_equal(g[2], w[1]) and g[3] == w[2]


def test_wrong_prediction_is_rebuilt_from_the_table(stand_in, monkeypatch):
    """Cell ids that do not ascend: the waiting cell is not the last run, the look-ahead is released and rebuilt."""
    b = _data()
    relabel = np.random.default_rng(3).permutation(N_CELLS).astype(b.cell.dtype)
    cell = relabel[b.cell]
    kw = dict(rows_per_file=2500, use_mixture_of_scanners=False, by_fragment=True, n_epochs=2)
    ahead = _proc(b, cell=cell, **kw)
    a = _drain(ahead)
    assert any(e[0] == "release" for e in ahead._engine.log)          # the fallback really ran
    sync = _proc(b, cell=cell, **kw)
    monkeypatch.setattr(sync, "_lookahead_ok", lambda: False)
    _same(a, _drain(sync))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_mos_index_that_overstates_a_cell(stand_in, monkeypatch, seed):
    """_cell_start_index claims five rows more for a cell in the middle of a chunk: the boundary-run prediction misses it, the
    table shows it incomplete; both feeders carry it to the final flush and emit it there."""
    b = _data()
    counts = np.diff(b.cell_start)
    counts[150] += 5
    csi = np.concatenate(([0], np.cumsum(counts)))
    kw = dict(csi=csi, use_mixture_of_scanners=True, n_scanners=3, prefetch_batch_size=1000, n_epochs=1)
    runs = []
    for lookahead in (True, False):
        proc = _proc(b, **kw)
        if not lookahead:
            monkeypatch.setattr(proc, "_lookahead_ok", lambda: False)
        np.random.seed(seed)
        runs.append(_drain(proc))
    _same(*runs)
    assert 150 in runs[0][-1][2].tolist() and all(150 not in x[2].tolist() for x in runs[0][:-1])
    assert sorted(np.concatenate([x[2] for x in runs[0]]).tolist()) == list(range(N_CELLS))


def test_engine_growth_is_deferred_until_nothing_is_in_flight(stand_in):
    b = _data()
    proc = _proc(b, rows_per_file=2500, use_mixture_of_scanners=False, by_fragment=True)
    proc._engine = small = StandInEngine(2510, 2510)                   # the first fragment fits, pending cell + second does not
    got = _drain(proc)
    assert proc._engine is not small and len(stand_in) == 1            # re-created once, by the fixture's engine_for (which asserts idleness)
    assert sorted(np.concatenate([x[2] for x in got]).tolist()) == list(range(N_CELLS))


def test_reset_for_epoch_releases_the_look_ahead(stand_in):
    b = _data()
    proc = _proc(b, rows_per_file=2500, use_mixture_of_scanners=False, by_fragment=True, n_epochs=2)
    proc.load_prefetch_batch()
    assert proc._ahead is not None and len(proc._engine.slots) == 1
    proc.reset_for_epoch(1)
    assert proc._ahead is None and not proc._engine.slots and proc.batch_id == 0
    rest = _drain(proc)
    assert {x[0] for x in rest} == {1} and sorted(np.concatenate([x[2] for x in rest]).tolist()) == list(range(N_CELLS))


@pytest.mark.parametrize("mos", [False, True])
@pytest.mark.parametrize("batch_size", [7, 64])
def test_iterable_dataset_batches_over_the_stand_in(stand_in, mos, batch_size):
    """Producer thread + consumer: batch_size-cell views, every cell once per epoch, the epoch key, a clean end."""
    b = _data()
    arr = SyntheticSLAFArray(ArrayExpressionSource(b.cell, b.gene, b.value, 1777), b.cell_start, N_GENES)
    tok = GeneformerTokenizer(arr, max_genes=16)
    np.random.seed(4)
    ds = SLAFIterableDataset(arr, tok, batch_size=batch_size, use_mixture_of_scanners=mos, by_fragment=True, verbose=False,
                             n_epochs=3, n_scanners=3, prefetch_batch_size=1000, max_queue_size=3)
    per_epoch: dict = {}
    for batch in ds:
        assert set(batch) == {"input_ids", "attention_mask", "cell_ids", "epoch"}
        assert 0 < batch["input_ids"].shape[0] <= batch_size and batch["input_ids"].shape[0] == batch["cell_ids"].shape[0]
        per_epoch.setdefault(batch["epoch"], []).extend(batch["cell_ids"].tolist())
    assert sorted(per_epoch) == [0, 1, 2]
    for ids in per_epoch.values():
        assert sorted(ids) == list(range(N_CELLS))


# ---------------------------------------------------------------------------------------------- properties
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(lens=st.lists(st.integers(0, 40), min_size=2, max_size=60), rows_per_file=st.integers(5, 300), mos=st.booleans(),
       n_scanners=st.integers(1, 5), chunk=st.integers(1000, 1400), draw_seed=st.integers(0, 1000), n_epochs=st.integers(1, 2))
def test_property_look_ahead_equals_synchronous_feeder(stand_in, lens, rows_per_file, mos, n_scanners, chunk, draw_seed, n_epochs):
    """Any table sorted by cell (cells without rows included), any fragment size, either loading mode: same stream of loads."""
    lens = np.asarray(lens, np.int64)
    if lens.sum() == 0:
        lens[0] = 3
    cell = np.repeat(np.arange(len(lens), dtype=np.uint32), lens)
    rng = np.random.default_rng(draw_seed)
    gene = rng.integers(0, 100, len(cell)).astype(np.uint16)
    value = rng.integers(1, 50, len(cell)).astype(np.uint16)
    csi = np.concatenate(([0], np.cumsum(lens)))
    runs = []
    for lookahead in (True, False):
        arr = SyntheticSLAFArray(ArrayExpressionSource(cell, gene, value, rows_per_file), csi, 100)
        proc = PrefetchBatchProcessor(arr, RandomShuffle(), GeneformerTokenizer(arr, max_genes=8), seed=7, verbose=False, n_epochs=n_epochs,
                                      use_mixture_of_scanners=mos, by_fragment=True, n_scanners=n_scanners, prefetch_batch_size=chunk)
        if not lookahead:
            proc._lookahead_ok = lambda: False
        np.random.seed(draw_seed)
        runs.append(_drain(proc))
        assert not proc._engine.slots
    _same(*runs)
    present = np.flatnonzero(lens > 0).tolist()
    for epoch in range(n_epochs):
        seen = [int(c) for x in runs[0] if x[0] == epoch for c in x[2]]
        assert sorted(seen) == present


def test_error_while_reading_ahead_surfaces_after_the_current_load(stand_in):
    """A reader error met while staging load i+1 must not lose load i (already on the device): it is raised by the next call."""
    b = _data()
    proc = _proc(b, rows_per_file=2500, use_mixture_of_scanners=False, by_fragment=True)
    real = proc._next_pieces
    calls = {"n": 0}

    def flaky():
        calls["n"] += 1
        if calls["n"] == 3:
            raise OSError("fragment unreadable")
        return real()

    proc._next_pieces = flaky
    first = proc.load_prefetch_batch()            # stages loads 0 and 1
    second = proc.load_prefetch_batch()           # staging load 2 fails; load 1 is still delivered
    assert (first.batch_id, second.batch_id) == (0, 1) and not proc._engine.slots
    with pytest.raises(OSError, match="fragment unreadable"):
        proc.load_prefetch_batch()


def _chunk(ids):
    ids = np.asarray(ids, np.uint32)
    from slaf_b200.ml.sources import CooChunk

    return CooChunk(ids, np.arange(len(ids), dtype=np.uint16), np.ones(len(ids), np.uint16))


def test_boundary_runs_merge_across_adjacent_pieces():
    runs = PrefetchBatchProcessor._boundary_runs([_chunk([4, 4]), _chunk([4, 5, 5, 6]), _chunk([6, 6]), _chunk([6, 7, 8, 8]), _chunk([]), _chunk([9])])
    assert [(r[0], r[1]) for r in runs] == [(4, 3), (6, 4), (8, 2), (9, 1)]          # 5 and 7 are interior: never candidates
    assert [sum(len(v) for v in r[2]) for r in runs] == [3, 4, 2, 1]
    assert [int(v.gene[0]) for v in runs[1][2]] == [3, 0, 0]                        # views in row order: tail of piece 2, piece 3, head of piece 4
    assert PrefetchBatchProcessor._boundary_runs([]) == []
    one = PrefetchBatchProcessor._boundary_runs([_chunk([3, 3, 3])])
    assert [(r[0], r[1]) for r in one] == [(3, 3)]


def test_mos_build_puts_a_generators_pending_cell_in_front_of_its_chunk():
    """datasets.py:748-760 and 880-888: generator 1 last delivered cell 6 and continues with 6 -> its pending rows go right in
    front of its chunk; the other pending cell (2) goes in front of everything; runs of it at a chunk head are moved up."""
    partials = {2: _chunk([2, 2]), 6: _chunk([6])}
    reads = [(0, [_chunk([2, 3, 3])], 1), (1, [_chunk([6, 6, 7])], 6)]
    pieces = PrefetchBatchProcessor._mos_build(reads, partials)
    assert np.concatenate([p.cell for p in pieces]).tolist() == [2, 2, 2, 3, 3, 6, 6, 6, 7]
    assert list(partials) == [2]                                                   # 6 was consumed by its generator


@pytest.mark.parametrize("mos", [False, True])
def test_arrow_fragments_are_fed_as_record_batch_views(stand_in, tmp_path, mos):
    """An Arrow fragment with several record batches reaches the engine as views of those batches (``to_pieces``), never gathered
    into one array on the host; the stream of loads is the one of the same table held in one numpy array per column."""
    from slaf_b200.distributed import RowRangeSource
    from slaf_b200.ml.sources import ArrowExpressionSource, write_arrow_dataset

    b = _data()
    write_arrow_dataset(str(tmp_path / "expr"), b.cell, b.gene, b.value, max_rows_per_file=2500, rows_per_batch=700)
    arrow = ArrowExpressionSource(str(tmp_path / "expr"))
    frag = arrow.get_fragments()[1]
    pieces = frag.to_pieces()
    assert [len(p) for p in pieces] == [700, 700, 700, 400] and [p.row_lo for p in pieces] == [2500, 3200, 3900, 4600]
    assert all(not p.cell.flags.owndata for p in pieces)                       # views of the memory map
    whole = frag.to_table()
    assert np.array_equal(np.concatenate([p.cell for p in pieces]), whole.cell) and np.array_equal(np.concatenate([p.value for p in pieces]), whole.value)
    sliced = RowRangeSource(arrow, 3000, 4700).get_fragments()[0].to_pieces()  # a rank's cell-aligned row range keeps the views
    assert [len(p) for p in sliced] == [200, 700, 700, 100] and sliced[0].row_lo == 3000
    runs = []
    for source in (arrow, ArrayExpressionSource(b.cell, b.gene, b.value, 2500)):
        arr = SyntheticSLAFArray(source, b.cell_start, N_GENES)
        proc = PrefetchBatchProcessor(arr, RandomShuffle(), GeneformerTokenizer(arr, max_genes=16), seed=42, verbose=False, n_epochs=2,
                                      use_mixture_of_scanners=mos, by_fragment=True, n_scanners=2, prefetch_batch_size=1000)
        np.random.seed(9)
        runs.append(_drain(proc))
    if not mos:      # under mixture-of-scanners the chunking differs (record batches vs array slices), so only coverage is fixed
        _same(*runs)
    for run in runs:
        for epoch in (0, 1):
            assert sorted(int(c) for x in run if x[0] == epoch for c in x[2]) == list(range(N_CELLS))


def test_reset_for_epoch_from_the_consumer_thread_while_the_prefetcher_runs(stand_in):
    """reference tests/test_pytorch_datasets.py:1064-1067 resets the processor of a running dataset from the main thread; with
    loads in flight that must not interleave with the prefetcher's load (the processor serialises the two)."""
    b = _data()
    arr = SyntheticSLAFArray(ArrayExpressionSource(b.cell, b.gene, b.value, 900), b.cell_start, N_GENES)
    ds = SLAFIterableDataset(arr, GeneformerTokenizer(arr, max_genes=16), batch_size=16, use_mixture_of_scanners=False, by_fragment=True,
                             verbose=False, n_epochs=3, max_queue_size=2)
    it = iter(ds)
    first = [next(it) for _ in range(5)]
    assert all(x["epoch"] == 0 for x in first)
    ds.batch_processor.reset_for_epoch(2)
    assert ds.batch_processor.current_epoch == 2
    rest = list(it)
    assert ds.prefetcher.error is None
    tail = [x for x in rest if x["epoch"] == 2]
    assert tail and rest[-1]["epoch"] == 2
    ids = [int(c) for x in tail for c in x["cell_ids"]]
    assert sorted(ids) == list(range(N_CELLS))                      # the epoch after the reset is complete and clean
    assert not any(eng.slots for eng in stand_in)                   # nothing left on the device


@pytest.mark.parametrize("mode", ["fragment", "sequential", "mos"])
def test_cell_start_index_route_gives_the_same_stream(stand_in, mode):
    """use_cell_start_index=True: the segment table handed to the engine is derived from _cell_start_index and the chunk
    positions alone; loads, order and pending cells are those of the default route (which reads the table off the cell column)."""
    b = _data()
    kw = dict(use_mixture_of_scanners=(mode == "mos"), by_fragment=(mode != "sequential"), n_scanners=3, prefetch_batch_size=1000,
              batches_per_chunk=2 if mode == "sequential" else 1, n_epochs=2)
    runs = []
    for csi_route in (False, True):
        proc = _proc(b, use_cell_start_index=csi_route, **kw)
        np.random.seed(6)
        runs.append(_drain(proc))
        assert not proc._engine.slots
    _same(*runs)


def test_cell_start_index_that_does_not_match_the_rows_is_refused(stand_in):
    b = _data()
    for shift in (3, -2):
        csi = b.cell_start.copy()
        csi[1:-1] += shift                                             # every cell starts somewhere else than the index says
        proc = _proc(b, csi=csi, use_mixture_of_scanners=False, by_fragment=True, use_cell_start_index=True)
        with pytest.raises(ValueError, match="_cell_start_index does not describe"):
            _drain(proc)
    relabel = _proc(b, cell=(b.cell + 1).astype(b.cell.dtype), csi=np.concatenate(([0], b.cell_start)), use_mixture_of_scanners=False,
                    by_fragment=True, use_cell_start_index=True)
    _drain(relabel)                                                    # ids shifted together with the index: consistent again


def test_cell_start_index_is_the_default_route_and_falls_back_when_stale(stand_in):
    """use_cell_start_index=None (the default): the index is used while it validates; a stale one is noticed by the probes, warned
    about once, and the processor continues on the cell column -- same stream either way."""
    b = _data()
    ref = _drain(_proc(b, use_mixture_of_scanners=False, by_fragment=True))
    auto = _proc(b, use_mixture_of_scanners=False, by_fragment=True, use_cell_start_index=None)
    assert auto._csi_arr is not None and not auto._lookahead_ok()
    got = _drain(auto)
    _same(ref, got)
    assert not any(kind == "segments" and False for kind, _ in auto._engine.log)
    csi = b.cell_start.copy()
    csi[1:-1] += 3
    stale = _proc(b, csi=csi, use_mixture_of_scanners=False, by_fragment=True, use_cell_start_index=None)
    with pytest.warns(RuntimeWarning, match="_cell_start_index does not describe"):
        got = _drain(stale)
    assert stale._csi_arr is None
    _same(ref, got)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("which", [150, 68, 129, 165, N_CELLS - 1, "one_inside_every_batch"])     # (68, seeds 1 and 3) hit the epoch edge
def test_misprediction_on_the_last_load_of_an_epoch_keeps_the_epoch(stand_in, monkeypatch, seed, which):
    """The index overstates a cell, so the look-ahead predicts "nothing pending" where the table later shows a leftover.  When that
    happens on the last load of an epoch the look-ahead has already opened the next epoch; the leftover must still be flushed as
    its own load of the OLD epoch (epoch tag, batch id, seed and order of the synchronous feeder), not spliced into the new one."""
    b = _data()
    counts = np.diff(b.cell_start)
    if which == "one_inside_every_batch":     # whatever the last load of the epoch is, it holds a cell the prediction cannot see
        inner = {int(b.cell[r]) for r in range(400, b.n_rows, 1000) if b.cell[r - 30] != b.cell[r + 30]}
        counts[sorted(inner)] += 5
    else:
        counts[which] += 5
    csi = np.concatenate(([0], np.cumsum(counts)))
    kw = dict(csi=csi, use_mixture_of_scanners=True, n_scanners=3, prefetch_batch_size=1000, n_epochs=3)
    runs = []
    for lookahead in (True, False):
        proc = _proc(b, **kw)
        if not lookahead:
            monkeypatch.setattr(proc, "_lookahead_ok", lambda: False)
        np.random.seed(seed)
        runs.append(_drain(proc))
    _same(*runs)
    for ep in range(3):
        ids = np.concatenate([x[2] for x in runs[0] if x[0] == ep]).tolist()
        assert sorted(ids) == list(range(N_CELLS)), f"epoch {ep} is complete and clean"
