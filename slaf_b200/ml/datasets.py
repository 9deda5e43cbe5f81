"""Prefetch batch processor, async prefetcher and iterable dataset: mirrors of the reference's
``slaf.ml.datasets`` for the hot path (slaf/ml/datasets.py:240-296 containers, :299-1103
``PrefetchBatchProcessor``, :1105-1459 ``AsyncPrefetcher``, :1462-1938 ``SLAFIterableDataset``).

What changes against the reference:

* the tokenized branch of ``load_prefetch_batch`` (shuffle -> window -> to_list -> tokenize,
  datasets.py:1011-1098) is ONE fused device pass: the combined rows go to the GPU once
  (pinned staging, async H2D on the engine's side stream), K1 builds the segment table, the host
  does the reference's completeness check on that table (C entries, not R rows) against
  ``_cell_start_index``, and K2 emits the token tensors of the complete cells in the CPython-exact
  shuffled order.  The rows of incomplete boundary cells are kept on the host exactly like the
  reference's ``partial_cell_data`` (datasets.py:891-953);
* the fused routes read ahead: the rows of load i+1 are submitted before the host waits for the segment table of load i; the
  incomplete cells load i+1 needs are predicted from the runs at the piece boundaries and checked against the table when it
  arrives (``_predict`` / ``_front``; a wrong prediction rebuilds the look-ahead, so the stream never differs from the
  synchronous feeder's);
* batches are device tensors (views of the prefetch batch); the queue hands over a CUDA event, the
  consumer stream waits on it -- no host synchronisation per batch;
* a full queue blocks the producer instead of silently dropping the batch (datasets.py:1277-1281).

Reading fragments (Lance / Arrow decode), the mixture-of-scanners sampling and the boundary
bookkeeping stay on the host, as in the reference.
"""

from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Any, Iterator

import numpy as np
import torch

from .. import _native as N
from ..core.tabular_schema import SLAF_LANCE_COO_SCHEMA
from ..engine import HotPathEngine, is_page_locked, pinned_empty, shuffle_perm
from . import frames
from .samplers import RandomShuffle, StratifiedShuffle
from .sources import CooChunk, ExpressionSource, concat_chunks, open_expression
from .tokenizers import GeneformerTokenizer, ScGPTTokenizer, SLAFTokenizer

try:  # IterableDataset keeps torch.utils.data.DataLoader(batch_size=None) compatibility
    from torch.utils.data import IterableDataset
except Exception:  # pragma: no cover
    IterableDataset = object  # type: ignore


# ----------------------------------------------------------------------------------------------
# containers (reference: datasets.py:240-296)
# ----------------------------------------------------------------------------------------------
class TokenizedPrefetchBatch:
    """Same fields as the reference's dataclass; ``cell_integer_ids`` (a Python list there) is
    materialised lazily from the device tensor ``cell_ids`` so the producer never synchronises."""

    def __init__(self, batch_id: int, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                 cell_integer_ids: list[int] | None = None, values: torch.Tensor | None = None,
                 partial_cell_data: dict | None = None, tokenize_time: float = 0.0, cell_ids: torch.Tensor | None = None,
                 epoch: int = 0, ready_event: Any = None):
        self.batch_id = batch_id
        self.input_ids = input_ids
        self.attention_mask = attention_mask
        self.values = values
        self.partial_cell_data = partial_cell_data
        self.tokenize_time = tokenize_time
        self.epoch = epoch
        self.ready_event = ready_event
        self._cell_list = list(cell_integer_ids) if cell_integer_ids is not None else None
        if cell_ids is None:
            cell_ids = torch.tensor(self._cell_list or [], dtype=torch.long, device=input_ids.device)
        self.cell_ids = cell_ids

    @property
    def cell_integer_ids(self) -> list[int]:
        if self._cell_list is None:
            if self.ready_event is not None:
                self.ready_event.synchronize()
            self._cell_list = self.cell_ids.cpu().tolist()
        return self._cell_list

    @property
    def n_cells(self) -> int:
        return int(self.input_ids.shape[0])


@dataclass
class RawPrefetchBatch:
    """reference: datasets.py:288-296 (frames are dict-of-columns / Arrow here, polars there)."""

    batch_id: int
    batch_dfs: list[Any]
    cell_integer_ids: list[int]
    process_time: float
    memory_mb: float


PrefetchBatch = TokenizedPrefetchBatch | RawPrefetchBatch


@dataclass
class _Staged:
    """One load of the fused route between reading its rows and tokenizing them (see ``load_prefetch_batch``)."""

    body: list                       # rows read from the source, without the partial cells put in front
    pieces: list                     # partial cells + body: what is (or will be) submitted
    flush: bool                      # end-of-data flush of the pending partial cells: every cell counts as complete
    batch_id: int
    epoch: int
    seed: int
    start_time: float
    eng: Any = None
    slot: int | None = None          # None: not submitted (yet)
    pred: list | None = None         # [(cell id, rows), ...] of the cells predicted to stay incomplete, in row order; None = no prediction
    stash: dict | None = None        # the predicted partial_cell_data after this load
    reads: list | None = None        # mixture-of-scanners: the draw (generator, chunks, last cell before) the pieces were built from
    mispredicted: bool = False       # the segment table contradicted ``pred`` (the look-ahead was repaired from the table)


def _runs(cell: np.ndarray):
    """(starts, ends, ids) of the maximal runs of equal ids, in row order (host, O(R))."""
    n = len(cell)
    if n == 0:
        z = np.zeros(0, np.int64)
        return z, z, cell[:0]
    heads = np.flatnonzero(cell[1:] != cell[:-1]) + 1
    starts = np.concatenate(([0], heads)).astype(np.int64)
    ends = np.concatenate((heads, [n])).astype(np.int64)
    return starts, ends, cell[starts]


def _head_run(cell: np.ndarray, x) -> int:
    """Number of leading rows equal to x (galloping: touches O(run length) rows)."""
    n, w = len(cell), 256
    while True:
        neq = np.flatnonzero(cell[:w] != x)
        if neq.size:
            return int(neq[0])
        if w >= n:
            return n
        w *= 4


def _tail_run(cell: np.ndarray, x) -> int:
    """Number of trailing rows equal to x."""
    return _head_run(cell[::-1], x)


def _splice_partials(partials: dict, body: list[CooChunk]) -> list[CooChunk]:
    """[partial rows of cell X, the runs of X found at the head / tail of later chunks, ...other partials..., rest].
    All pieces are views; the order of cells is the order of first appearance in [partials..., body] and the
    rows of a cell keep their arrival order -- what partition_by(maintain_order) + concat produce."""
    if not partials:
        return body
    body = list(body)
    out: list[CooChunk] = []
    for x, part in partials.items():
        out.append(part)
        for i, ch in enumerate(body):
            if len(ch) == 0:
                continue
            if ch.cell[0] == x:
                n = _head_run(ch.cell, x)
                out.append(ch.slice(0, n))
                body[i] = ch = ch.slice(n, len(ch))
            if len(ch) and ch.cell[-1] == x:
                n = _tail_run(ch.cell, x)
                out.append(ch.slice(len(ch) - n, len(ch)))
                body[i] = ch.slice(0, len(ch) - n)
    return out + [b for b in body if len(b)]


def _regroup_segments(pieces: list[CooChunk], seg_start: np.ndarray, seg_cell: np.ndarray) -> list[CooChunk]:
    """Safety net for cells that still arrive in separated runs: reorder whole SEGMENTS (not rows) so that
    every cell is one run, cells by first appearance, runs of a cell in arrival order.  Returns views."""
    _, first, inv = np.unique(seg_cell, return_index=True, return_inverse=True)
    order = np.argsort(first[inv], kind="stable")
    offs = np.concatenate(([0], np.cumsum([len(p) for p in pieces])))
    out: list[CooChunk] = []
    run_lo = run_hi = None
    ranges = []
    for j in order:
        lo, hi = int(seg_start[j]), int(seg_start[j + 1])
        if run_hi == lo:
            run_hi = hi
        else:
            if run_lo is not None:
                ranges.append((run_lo, run_hi))
            run_lo, run_hi = lo, hi
    if run_lo is not None:
        ranges.append((run_lo, run_hi))
    for lo, hi in ranges:
        k = int(np.searchsorted(offs, lo, side="right")) - 1
        while lo < hi:
            a, b = lo - int(offs[k]), min(hi, int(offs[k + 1])) - int(offs[k])
            if b > a:
                out.append(pieces[k].slice(a, b))
            lo = int(offs[k + 1])
            k += 1
    return out


class _Staging:
    """Ring of pinned host buffers the combined rows of a batch are assembled in (so the H2D copy
    is a true async DMA).  Copies run on a small thread pool: numpy releases the GIL in memcpy."""

    def __init__(self, depth: int = 2, threads: int = 4):
        self.depth = depth
        self.capacity = 0
        self.bufs: list[tuple[np.ndarray, np.ndarray, np.ndarray]] = []
        self.turn = 0
        self.pool = ThreadPoolExecutor(max_workers=max(1, threads)) if threads > 1 else None
        self.users: list[tuple[Any, int] | None] = [None] * depth     # (engine, slot) whose H2D copy last read each buffer
        self.last = 0

    def used_by(self, eng: Any, slot: int) -> None:
        """The buffer handed out last is the DMA source of `slot`: it is not refilled before that copy has finished."""
        self.users[self.last] = (eng, slot)

    def ensure(self, rows: int) -> None:
        if rows <= self.capacity:
            return
        cap = max(rows + rows // 4, 1 << 16)
        self._wait_all()
        self.bufs = [tuple(pinned_empty(cap * 4, np.uint8) for _ in range(3)) for _ in range(self.depth)]
        self.capacity = cap

    def _wait(self, i: int) -> None:
        user, self.users[i] = self.users[i], None
        if user is not None:
            eng, slot = user
            if getattr(eng, "_ctx", True) and hasattr(eng, "wait_loaded"):
                eng.wait_loaded(slot)            # (a later batch in that slot sits behind ours on the same copy stream: waiting for it covers ours)

    def _wait_all(self) -> None:
        for i in range(len(self.users)):
            self._wait(i)

    def assemble(self, pieces: list[CooChunk], with_cell: bool = True) -> CooChunk:
        """The pieces back to back in the next pinned buffer of the ring.  ``with_cell=False`` leaves the cell column out
        (cell-start-index route: it does not travel); the returned chunk's ``cell`` is then an empty view."""
        n = sum(len(p) for p in pieces)
        self.ensure(n)
        self._wait(self.turn)                    # the copy engine may still be reading what this buffer held three loads ago
        raw = self.bufs[self.turn]
        self.last = self.turn
        self.turn = (self.turn + 1) % self.depth
        first = pieces[0]
        dst = CooChunk(*(raw[i][: (n if (i or with_cell) else 0) * src.dtype.itemsize].view(src.dtype)
                         for i, src in enumerate((first.cell, first.gene, first.value))))
        tasks = []
        off = 0
        step = 1 << 21   # rows per copy task
        for p in pieces:
            if (p.cell.dtype, p.gene.dtype, p.value.dtype) != (dst.cell.dtype, dst.gene.dtype, dst.value.dtype):
                raise TypeError("fragments of one expression table must share column dtypes")
            for lo in range(0, len(p), step):
                hi = min(lo + step, len(p))
                for d, s_ in (((dst.cell, p.cell),) if with_cell else ()) + ((dst.gene, p.gene), (dst.value, p.value)):
                    tasks.append((d, off + lo, s_, lo, hi))
            off += len(p)
        if self.pool is not None and len(tasks) > 3:
            list(self.pool.map(lambda t: np.copyto(t[0][t[1]: t[1] + (t[4] - t[3])], t[2][t[3]: t[4]]), tasks))
        else:
            for d, o, s_, lo, hi in tasks:
                np.copyto(d[o: o + (hi - lo)], s_[lo:hi])
        return dst

    def close(self) -> None:
        if self.pool is not None:
            self.pool.shutdown(wait=False)
        try:
            self._wait_all()
        except Exception:
            pass
        self.bufs = []
        self.capacity = 0


# ----------------------------------------------------------------------------------------------
# PrefetchBatchProcessor
# ----------------------------------------------------------------------------------------------
class PrefetchBatchProcessor:
    """reference: slaf/ml/datasets.py:299-1103.  Same constructor arguments, attributes, error
    messages, epoch handling and ``StopIteration("No more epochs available")`` protocol."""

    def __init__(self, slaf_array: Any, shuffle: Any, tokenizer: SLAFTokenizer | None, seed: int = 42,
                 batches_per_chunk: int = 1, n_expression_bins: int = 10, use_binned_expressions: bool = True,
                 n_epochs: int = 1, raw_mode: bool = False, verbose: bool = True, log_metrics: bool = False,
                 batch_size: int = 32, by_fragment: bool = True, use_mixture_of_scanners: bool = True, n_scanners: int = 16,
                 prefetch_batch_size: int = 4194304, parallelize_fragment_reads: bool = False, device=None,
                 copy_threads: int = 4, device_raw: bool = False, use_cell_start_index: bool | None = None, csi_check: int = 1):
        self.slaf_array = slaf_array
        self.device_raw = device_raw
        # EXTENSION: take the cell boundaries from slaf_array._cell_start_index (slaf/core/slaf.py:645-702) instead of the cell
        # column, so that column neither crosses PCIe nor is scanned (needs a source that knows the table position of its
        # chunks).  None (default) = use the index whenever the array has one and it keeps validating against the rows
        # (probes, csi_check: 1 sampled / 2 every cell); the first mismatch switches this processor to the cell-column route for
        # good.  True = insist (a mismatch is a ValueError).  False = always ship and scan the cell column (the reference's contract).
        self.use_cell_start_index = use_cell_start_index
        self.csi_check = int(csi_check)
        self._csi_arr: np.ndarray | None = None
        if use_cell_start_index is not False:
            idx = getattr(slaf_array, "_cell_start_index", None)
            if idx is not None:
                self._csi_arr = np.ascontiguousarray(np.asarray(idx), np.int64)
            elif use_cell_start_index:
                raise ValueError("use_cell_start_index=True needs slaf_array._cell_start_index")
        self.shuffle = shuffle
        self.tokenizer = tokenizer
        self.seed = seed
        self.batches_per_chunk = batches_per_chunk
        self.n_expression_bins = n_expression_bins
        self.use_binned_expressions = use_binned_expressions
        self.n_epochs = n_epochs
        self.raw_mode = raw_mode
        self.verbose = verbose
        self.log_metrics = log_metrics
        self.batch_size = batch_size
        self.by_fragment = by_fragment
        self.use_mixture_of_scanners = use_mixture_of_scanners
        self.n_scanners = n_scanners
        self.prefetch_batch_size = prefetch_batch_size
        self.parallelize_fragment_reads = parallelize_fragment_reads
        self.device = device
        if self.use_mixture_of_scanners:
            if self.n_scanners < 1:
                raise ValueError("n_scanners must be at least 1")
            if self.n_scanners > 100:
                raise ValueError("n_scanners cannot exceed 100")
            if self.prefetch_batch_size < 1000:
                raise ValueError("prefetch_batch_size must be at least 1,000")
            if self.prefetch_batch_size > 10000000:
                raise ValueError("prefetch_batch_size cannot exceed 10,000,000")
        self.batch_id = 0
        self.current_epoch = 0
        self.partial_cell_data: dict[Any, CooChunk] = {}
        self._non_mos_pending_tail_flush = False
        self._mos_final_flush = False
        self.window_kwargs: dict[str, Any] = {}
        self.total_prefetch_cells = 0
        self.expression_dataset: ExpressionSource = open_expression(slaf_array)
        if self.use_mixture_of_scanners:
            self.by_fragment = True
        self._open_readers()
        self._timing_metrics: dict[str, list[float]] | None = (
            {"lance_loading": [], "window": [], "shuffle": [], "tokenize": [], "total": [], "cells_processed": []}
            if log_metrics else None)
        self._engine: HotPathEngine | None = None
        self._staging = _Staging(depth=3, threads=copy_threads)     # three: with the look-ahead two loads can be in flight while a third is assembled
        self._ahead: _Staged | BaseException | None = None
        self._inject: list[_Staged] = []       # loads that have to run before ``_ahead`` (an epoch's late flush, see _restage_ahead)
        # the reference's tests (and users) call reset_for_epoch from the consumer thread while the prefetcher thread is inside
        # load_prefetch_batch (tests/test_pytorch_datasets.py:1064-1067); with batches in flight on the device those two must
        # not interleave
        self._lock = threading.RLock()
        self._read_pool: ThreadPoolExecutor | None = None

    # ------------------------------------------------------------------ readers / epochs
    def _open_readers(self) -> None:
        if self.use_mixture_of_scanners:
            self.fragment_generators: list[Iterator[CooChunk]] = []
            self.generator_last_cells: list[int | None] = []
            self.generator_active: list[bool] = []
            for fragment in self.expression_dataset.get_fragments():
                self.fragment_generators.append(iter(fragment.to_batches(batch_size=self.prefetch_batch_size)))
                self.generator_last_cells.append(None)
                self.generator_active.append(True)
        elif self.by_fragment:
            self.fragment_iterator = iter(self.expression_dataset.get_fragments())
        else:
            self.batch_generator = iter(self.expression_dataset.to_batches())

    def reset_for_epoch(self, epoch: int) -> None:
        if epoch < 0 or epoch >= self.n_epochs:
            raise ValueError(f"Invalid epoch {epoch}. Must be 0 <= epoch < {self.n_epochs}")
        with self._lock:
            self._drop_ahead()
            self.current_epoch = epoch
            self.batch_id = 0
            self.partial_cell_data = {}
            self._non_mos_pending_tail_flush = False
            self._mos_final_flush = False
            self._open_readers()

    def get_timing_metrics(self) -> dict | None:
        if not self.log_metrics or self._timing_metrics is None:
            return None
        return {k: (sum(v) / len(v) if v else 0.0) for k, v in self._timing_metrics.items()}

    def _record_timing(self, step: str, duration: float, cells_processed: int = 0) -> None:
        if not self.log_metrics or self._timing_metrics is None:
            return
        if step in self._timing_metrics and step != "cells_processed":
            self._timing_metrics[step].append(duration)
        if cells_processed > 0:
            self._timing_metrics["cells_processed"].append(cells_processed)

    def _read_from_generator(self, idx: int) -> tuple[list[CooChunk] | None, int, bool]:
        """``batches_per_chunk`` batches of one fragment generator (datasets.py:561-606).  Unlike the
        reference, batches already read are kept when the generator ends inside the chunk (the
        reference drops them, losing rows whenever batches_per_chunk > 1)."""
        got: list[CooChunk] = []
        exhausted = False
        for _ in range(self.batches_per_chunk):
            try:
                got.append(next(self.fragment_generators[idx]))
            except StopIteration:
                exhausted = True
                break
        if not got:
            return None, idx, True
        if exhausted:
            self.generator_active[idx] = False
        return [g for g in got if len(g)], idx, False      # consecutive storage rows: kept as separate views, never copied

    def _mos_read(self) -> tuple[str, list | None, int]:
        """One mixture-of-scanners draw (datasets.py:672-760): ("rows", [(generator, chunks, its last cell before), ...], n) |
        ("none", ...) nothing came back | ("flush", ...) every generator is exhausted and partial cells are pending |
        ("epoch", ...) every generator is exhausted.  Advances the generators, does not touch partial_cell_data."""
        active = [i for i, a in enumerate(self.generator_active) if a]
        if not active:
            return ("flush" if self.partial_cell_data else "epoch"), None, 0
        n_to_sample = min(self.n_scanners, len(active))
        selected = np.random.choice(active, size=n_to_sample, replace=False)      # datasets.py:689 (unseeded there too)
        if self.parallelize_fragment_reads and len(selected) > 1:
            if self._read_pool is None:
                self._read_pool = ThreadPoolExecutor(max_workers=min(32, self.n_scanners))
            results = list(self._read_pool.map(self._read_from_generator, [int(i) for i in selected]))
        else:
            results = [self._read_from_generator(int(i)) for i in selected]
        reads = []
        for got, idx, exhausted in results:
            if exhausted:
                self.generator_active[idx] = False
                continue
            if not got:
                continue
            reads.append((idx, got, self.generator_last_cells[idx]))
            self.generator_last_cells[idx] = int(got[-1].cell[-1])
        if not reads:
            return "none", None, 0
        return "rows", reads, len(reads)

    @staticmethod
    def _mos_build(reads: list, partials: dict) -> list[CooChunk]:
        """Rows of one load in the order the reference concatenates them: a generator's own pending cell goes right in front of
        its chunk (datasets.py:748-760), the other pending cells in front of everything (datasets.py:880-888), and runs of a
        pending cell at the head / tail of a chunk are moved up behind it -- where partition_by would gather them.
        ``partials`` is consumed (entries put in front of a chunk are popped)."""
        body: list[CooChunk] = []
        for _idx, got, last in reads:
            if last is not None:
                first_cell = int(got[0].cell[0])
                if (first_cell == last or first_cell == last + 1) and last in partials:
                    got = [partials.pop(last)] + got
            body.extend(got)
        return _splice_partials(partials, body)

    def _next_pieces(self) -> tuple[list[CooChunk], int] | None:
        """Rows of the next load: (pieces in row order, batch_count); None = epoch transition handled
        (caller loops).  Raises StopIteration at the end of the last epoch."""
        pieces: list[CooChunk] = []
        if self.use_mixture_of_scanners:
            kind, reads, n_chunks = self._mos_read()
            if kind == "flush":
                pieces = list(self.partial_cell_data.values())
                self.partial_cell_data = {}
                self._mos_final_flush = True
                return pieces, 0
            if kind == "epoch":
                return self._advance_epoch()
            if kind == "none":
                return None
            return self._mos_build(reads, self.partial_cell_data), n_chunks
        elif self.by_fragment:
            try:
                fragment = next(self.fragment_iterator)
                pieces, count = fragment.to_pieces(), 1          # storage chunks as they lie (views), never concatenated on the host
            except StopIteration:
                if self.partial_cell_data:
                    pieces, count = list(self.partial_cell_data.values()), 0
                    self.partial_cell_data = {}
                    self._non_mos_pending_tail_flush = True
                    return pieces, count
                return self._advance_epoch()
        else:
            got = []
            for _ in range(self.batches_per_chunk):
                try:
                    got.append(next(self.batch_generator))
                except StopIteration:
                    break
            if not got:
                if self.partial_cell_data:
                    pieces, count = list(self.partial_cell_data.values()), 0
                    self.partial_cell_data = {}
                    self._non_mos_pending_tail_flush = True
                    return pieces, count
                return self._advance_epoch()
            pieces, count = got, len(got)
        if self.partial_cell_data:                                                        # datasets.py:880-888
            pieces = list(self.partial_cell_data.values()) + pieces           # the last cell's rows continue at the top of the new rows
        return pieces, count

    def _advance_epoch(self) -> None:
        if self.current_epoch + 1 < self.n_epochs:
            self.reset_for_epoch(self.current_epoch + 1)
            return None
        raise StopIteration("No more epochs available") from None

    # ------------------------------------------------------------------ completeness (host, C entries)
    def _complete_mask(self, seg_cell: np.ndarray, seg_len: np.ndarray, flush: bool | None = None) -> np.ndarray:
        """Which cells of the combined rows are complete (datasets.py:891-953).  ``flush``: the load is the end-of-data flush
        of the pending partial cells (taken from the processor's flag when not given)."""
        if flush is None:
            flush = self._take_flush_flag()
        if self.use_mixture_of_scanners:
            csi = np.asarray(self.slaf_array._cell_start_index)
            ids = seg_cell.astype(np.int64)
            expected = csi[ids + 1] - csi[ids]
            done = seg_len >= expected
            if flush and not done.any():
                # every generator is exhausted: nothing can complete these cells any more.  The reference
                # would spin forever on such rows (datasets.py:660-668 + 1099-1102); emit them instead.
                return np.ones(len(seg_cell), bool)
            return done
        if flush:
            return np.ones(len(seg_cell), bool)
        return seg_cell < seg_cell.max()

    def _stash_partials(self, pieces: list[CooChunk], seg_start: np.ndarray, seg_cell: np.ndarray, complete: np.ndarray) -> None:
        """Keep the rows of the incomplete cells on the host (copies: the pieces may be recycled buffers)."""
        self.partial_cell_data = {}
        bad = np.flatnonzero(~complete)
        if bad.size == 0:
            return
        offs = np.concatenate(([0], np.cumsum([len(p) for p in pieces])))
        for j in bad:
            lo, hi = int(seg_start[j]), int(seg_start[j + 1])
            parts = []
            k = int(np.searchsorted(offs, lo, side="right")) - 1
            while lo < hi:
                a, b = lo - int(offs[k]), min(hi, int(offs[k + 1])) - int(offs[k])
                p = pieces[k]
                parts.append(p.slice(a, b).copy())
                lo = int(offs[k + 1])
                k += 1
            part = concat_chunks(parts)
            key = int(seg_cell[j])
            self.partial_cell_data[key] = concat_chunks([self.partial_cell_data[key], part]) if key in self.partial_cell_data else part

    # ------------------------------------------------------------------ engine
    def _engine_for(self, rows: int, all_cells: bool = False) -> HotPathEngine:
        eng = self._engine
        if eng is None or eng.max_rows < rows or (all_cells and eng.max_cells < rows):
            if eng is not None:
                torch.cuda.synchronize(eng.device)
                eng.close()
            from ..runtime import resolve_device

            cap = max(rows + rows // 4, 1 << 20)
            cells = cap if all_cells else cap // 8 + (1 << 16)     # cells of >= 8 rows on average; grown on demand
            self._engine = eng = HotPathEngine(resolve_device(self.device if self.device is not None
                                                              else getattr(self.tokenizer, "device", None)), max_rows=cap, max_cells=cells)
        return eng

    def _fused_ok(self) -> bool:
        return (type(self.shuffle) in (RandomShuffle, StratifiedShuffle) and type(self.tokenizer) in (GeneformerTokenizer, ScGPTTokenizer)
                and type(self.tokenizer.window).__module__ == "slaf_b200.ml.aggregators")

    # ------------------------------------------------------------------ the hot path
    def _lookahead_ok(self) -> bool:
        """The fused routes read ahead: the rows of load i+1 are submitted (H2D + CSR build) before the host waits for the
        segment table of load i.  What load i+1 needs from load i are its incomplete cells (datasets.py:891-953).  In a table
        sorted by cell only the cells at the head or tail of a piece can be cut, so the host predicts them from those runs
        alone (``_predict``), checks the prediction against the segment table when it arrives (``_front``) and, if it was wrong
        (rows not sorted by cell, an index that does not match), rebuilds the next load from the table."""
        if self._csi_arr is not None:
            return False                      # the table comes from the host: no load ever waits for the GPU, nothing to look ahead for
        if self.raw_mode:
            return self.device_raw and type(self.shuffle) in (RandomShuffle, StratifiedShuffle)
        return self._fused_ok()

    def load_prefetch_batch(self) -> PrefetchBatch:
        with self._lock:
            return self._load_prefetch_batch()

    def _load_prefetch_batch(self) -> PrefetchBatch:
        if self._lookahead_ok():
            return self._load_with_lookahead()
        while True:
            start_time = time.time()
            nxt = self._next_pieces()
            if nxt is None:
                continue
            pieces, _count = nxt
            pieces = [p for p in pieces if len(p)]
            load_time = time.time() - start_time
            self._record_timing("lance_loading", load_time)
            n_rows = sum(len(p) for p in pieces)
            if n_rows == 0:
                self.batch_id += 1
                continue
            seed = self.seed + self.batch_id + self.current_epoch * 10000               # datasets.py:1016
            if self.raw_mode and self.device_raw and type(self.shuffle) in (RandomShuffle, StratifiedShuffle):
                out = self._load_raw_device(self._sync_record(pieces, seed, start_time))
            elif self.raw_mode or not self._fused_ok():
                out = self._load_on_host_frames(pieces, seed, start_time)
            else:
                if self.tokenizer is None:
                    raise RuntimeError("Tokenizer is required for tokenized mode")
                out = self._load_fused(self._sync_record(pieces, seed, start_time))
            if out is None:                    # no complete cell in this chunk (datasets.py:1099-1102)
                self.batch_id += 1
                continue
            self.batch_id += 1
            return out

    def _sync_record(self, pieces: list[CooChunk], seed: int, start_time: float) -> _Staged:
        return _Staged(body=pieces, pieces=pieces, flush=self._take_flush_flag(), batch_id=self.batch_id, epoch=self.current_epoch,
                       seed=seed, start_time=start_time)

    def _take_flush_flag(self) -> bool:
        """Was the load just read the end-of-data flush of the pending partial cells?  (read and cleared)"""
        if self.use_mixture_of_scanners:
            flush, self._mos_final_flush = self._mos_final_flush, False
        else:
            flush, self._non_mos_pending_tail_flush = self._non_mos_pending_tail_flush, False
        return flush

    # -- look-ahead ------------------------------------------------------------------------------
    def _load_with_lookahead(self) -> PrefetchBatch:
        if self.tokenizer is None and not self.raw_mode:
            raise RuntimeError("Tokenizer is required for tokenized mode")
        while True:
            injected = bool(self._inject)
            if injected:
                cur = self._inject.pop(0)                      # the look-ahead (already staged, next epoch) keeps waiting behind it
            else:
                cur, self._ahead = self._ahead, None
            if cur is None:
                cur = self._stage()
            if isinstance(cur, BaseException):                 # the end of the data, or an error met while reading ahead:
                raise cur                                       # delivered after the load before it
            if not cur.pieces:                                  # a rebuilt flush load that turned out to have nothing left to flush
                continue
            if cur.slot is None:
                self._submit_staged(cur)                       # deferred: the engine had to grow, nothing else is in flight now
            if cur.pred is not None and not injected:
                try:
                    self._ahead = self._stage()                # the next load's rows go out before this one's table is awaited
                except Exception as end:                       # StopIteration (last epoch done) or a reader error
                    self._ahead = end
            out = self._load_raw_device(cur) if self.raw_mode else self._load_fused(cur)
            if injected and cur.stash and not cur.mispredicted:
                self._inject.insert(0, self._flush_record(cur.stash, cur))     # cells this flush could not complete: the epoch's next flush
            if out is not None:
                return out

    def _stage(self) -> _Staged:
        """Read the next load, predict its incomplete cells, submit its rows.  Raises StopIteration at the end of the last epoch."""
        while True:
            start_time = time.time()
            reads = None
            if self.use_mixture_of_scanners:
                kind, reads, _n = self._mos_read()
                if kind == "epoch":
                    self._advance_epoch()
                    continue
                if kind == "none":
                    continue
                if kind == "flush":
                    pieces, self.partial_cell_data, self._mos_final_flush = list(self.partial_cell_data.values()), {}, True
                else:
                    pieces = self._mos_build(reads, self.partial_cell_data)
                body: list[CooChunk] = []
            else:
                n_partials = len(self.partial_cell_data)
                nxt = self._next_pieces()
                if nxt is None:
                    continue
                pieces, _count = nxt
                body = [] if self._non_mos_pending_tail_flush else [p for p in pieces[n_partials:] if len(p)]
            flush = self._take_flush_flag()
            pieces = [p for p in pieces if len(p)]
            self._record_timing("lance_loading", time.time() - start_time)
            if sum(len(p) for p in pieces) == 0:
                self.batch_id += 1
                continue
            rec = _Staged(body=body, pieces=pieces, flush=flush, batch_id=self.batch_id, epoch=self.current_epoch,
                          seed=self.seed + self.batch_id + self.current_epoch * 10000, start_time=start_time, reads=reads)     # datasets.py:1016
            self.batch_id += 1
            self._predict(rec)
            eng = self._engine
            n_rows = sum(len(p) for p in pieces)
            if eng is not None and eng.max_rows >= n_rows:
                self._submit_staged(rec)                       # otherwise deferred until nothing is in flight (the engine is re-created)
            return rec

    @staticmethod
    def _boundary_runs(pieces: list[CooChunk]) -> list[tuple[int, int, list[CooChunk]]]:
        """(cell id, rows, [views]) of the runs that touch a piece boundary, in row order; a run that continues across adjacent
        pieces is one entry.  Touches O(run length) rows per piece."""
        out: list[tuple[int, int, list[CooChunk]]] = []
        open_run = None                                            # the run reaching the end of the previous piece
        for p in pieces:
            n = len(p)
            if n == 0:
                continue
            first = p.cell[0]
            h = _head_run(p.cell, first)
            if open_run is not None and open_run[0] == int(first):
                open_run = (open_run[0], open_run[1] + h, open_run[2] + [p.slice(0, h)])
            else:
                if open_run is not None:
                    out.append(open_run)
                open_run = (int(first), h, [p.slice(0, h)])
            if h < n:
                out.append(open_run)
                last = p.cell[-1]
                t = _tail_run(p.cell, last)
                open_run = (int(last), t, [p.slice(n - t, n)])
        if open_run is not None:
            out.append(open_run)
        return out

    def _predict(self, rec: _Staged) -> None:
        """partial_cell_data as it will be after ``rec``, from the runs at the piece boundaries (the rows are copied: the
        pieces may be recycled buffers).  Mixture of scanners: the runs shorter than the cell's row count in
        ``_cell_start_index``; otherwise the reference's rule for one stream, the cell with the largest id waits; the
        end-of-data flush leaves nothing.  No prediction (``pred`` None: this load is finished before the next is read) when a
        cell shows up in two separated runs -- that load goes through the splice safety net."""
        runs = self._boundary_runs(rec.pieces)
        ids = [r[0] for r in runs]
        if len(set(ids)) != len(ids):
            rec.pred = rec.stash = None
            return
        if self.use_mixture_of_scanners:
            csi = np.asarray(self.slaf_array._cell_start_index)
            a = np.asarray(ids, np.int64)
            done = np.asarray([r[1] for r in runs], np.int64) >= csi[a + 1] - csi[a]
            # the final flush holds nothing but pending cells (every run is a whole piece); if none of them can be completed
            # any more they are all emitted (see _complete_mask)
            bad = [] if (rec.flush and not done.any()) else [r for r, d in zip(runs, done) if not d]
        elif rec.flush:
            bad = []
        else:
            x = max(ids)
            bad = [r for r in runs if r[0] == x]
        rec.pred = [(r[0], r[1]) for r in bad]
        rec.stash = {r[0]: concat_chunks([v.copy() for v in r[2]]) for r in bad}
        self.partial_cell_data = dict(rec.stash)

    def _submit_staged(self, rec: _Staged) -> None:
        rec.eng = self._engine_for(sum(len(p) for p in rec.pieces))
        rec.slot, rec.pieces = self._submit(rec.eng, rec.pieces)

    def _unsubmit_ahead(self) -> None:
        """Take the look-ahead's rows back off the device (its slot is released; it is submitted again when its turn comes)."""
        a = self._ahead
        if isinstance(a, _Staged) and a.slot is not None:
            a.eng.release(a.slot)
            a.slot = a.eng = None

    def _drop_ahead(self) -> None:
        self._unsubmit_ahead()
        self._ahead = None
        self._inject = []

    def _flush_record(self, stash: dict, after: _Staged) -> _Staged:
        """The end-of-data flush of ``stash`` as the load that follows ``after`` in ITS epoch (batch id, seed and epoch tag of the
        synchronous feeder, datasets.py:660-668 + 1016).  The processor's reader state is left as it is."""
        rec = _Staged(body=[], pieces=[p for p in stash.values() if len(p)], flush=True, batch_id=after.batch_id + 1, epoch=after.epoch,
                      seed=self.seed + after.batch_id + 1 + after.epoch * 10000, start_time=time.time())
        keep = self.partial_cell_data
        self._predict(rec)                                      # (mixture of scanners: cells that still cannot complete wait for the next flush)
        self.partial_cell_data = keep
        return rec

    def _restage_ahead(self, stash: dict, rec: _Staged) -> None:
        """The prediction for the current load ``rec`` was wrong: rebuild the look-ahead from the true partial cells."""
        a = self._ahead
        if not isinstance(a, _Staged):
            self._ahead = None                                  # (an end-of-data marker is re-derived: pending cells get their flush load)
            self.partial_cell_data = stash
            return
        if a.epoch != rec.epoch:
            # ``rec`` was predicted to close its epoch with nothing pending, so the look-ahead already opened the next epoch.  The
            # cells that are pending after all get their own flush load in the OLD epoch -- what the synchronous feeder does --
            # ahead of the look-ahead, which was built on "nothing pending" and stays as it is.
            if stash:
                self._inject.insert(0, self._flush_record(stash, rec))
            return
        if a.flush:
            pieces = list(stash.values())
        elif a.reads is not None:
            pieces = self._mos_build(a.reads, dict(stash))
        else:
            pieces = list(stash.values()) + a.body
        a.pieces = [p for p in pieces if len(p)]
        self.partial_cell_data = {}
        self._predict(a)                                        # sets partial_cell_data to what follows the look-ahead
        if a.pred is None:
            self.partial_cell_data = {}                         # decided by the table when its turn comes (nothing is read ahead of it)

    #: pieces smaller than this are handed to the driver as they are (its pageable path copies them synchronously)
    SMALL_PIECE_ROWS = 1 << 15

    def _submit(self, eng: HotPathEngine, pieces: list[CooChunk]) -> tuple[int, list[CooChunk]]:
        """H2D + CSR build of the combined rows.  Page-locked pieces (pinned / registered source memory) go
        straight from where they live; otherwise the rows are assembled in the pinned staging ring first."""
        direct = all(len(p) <= self.SMALL_PIECE_ROWS or (is_page_locked(p.cell) and is_page_locked(p.gene) and is_page_locked(p.value))
                     for p in pieces)
        if direct and len(pieces) <= 256:
            return eng.submit_pieces([(p.cell, p.gene, p.value) for p in pieces]), pieces
        rows = self._staging.assemble(pieces)
        slot = eng.submit(rows.cell, rows.gene, rows.value)
        self._staging.used_by(eng, slot)
        return slot, [rows]

    def _submit_with_csi(self, eng: HotPathEngine, pieces: list[CooChunk], regrouped: bool = False):
        """Gene / value columns + the segment table the index implies (built and probed against the cell column inside the library,
        ``slafb_submit_csi``).  Returns (engine, slot, pieces, seg_start, seg_cell)."""
        direct = all(len(p) <= self.SMALL_PIECE_ROWS or (is_page_locked(p.gene) and is_page_locked(p.value)) for p in pieces)
        src = pieces
        if not (direct and len(pieces) <= 256):
            rows = self._staging.assemble(pieces, with_cell=False)      # one pinned block; the piece structure (table positions) is kept
            src, off = [], 0
            for p in pieces:
                src.append(CooChunk(p.cell, rows.gene[off:off + len(p)], rows.value[off:off + len(p)], p.row_lo))
                off += len(p)
        try:
            slot, _n = eng.submit_csi([(p.cell, p.gene, p.value) for p in src], [p.row_lo for p in src], self._csi_arr, check=self.csi_check)
        except N.SlafB200Error as err:
            if err.code == N.ERR_CAPACITY:                              # more cells than the segment tables hold: size for the worst case
                eng = self._engine_for(sum(len(p) for p in pieces), all_cells=True)
                return self._submit_with_csi(eng, pieces, regrouped)
            if err.code == N.ERR_INVALID and "cell_start_index" in err.message:
                raise ValueError(f"_cell_start_index does not describe the expression rows ({err.message})") from err
            raise
        if src is not pieces:
            self._staging.used_by(eng, slot)
        seg_start, seg_cell = eng.segments(slot)                       # (host-built table: no wait for the GPU)
        ascending = seg_cell.size < 2 or bool(np.all(seg_cell[1:] > seg_cell[:-1]))        # the usual case, and it implies distinct ids
        if not ascending and not regrouped and len(np.unique(seg_cell)) != len(seg_cell):  # a cell in two separated places: splice (views) first
            eng.release(slot)
            return self._submit_with_csi(eng, _regroup_segments(pieces, seg_start, seg_cell), True)
        return eng, slot, pieces, seg_start, seg_cell

    def _front(self, rec: _Staged):
        """submit -> segment table -> (splice safety net) -> completeness -> partial stash.
        Returns (engine, slot, indices of the complete cells in first-appearance order, their row counts)."""
        pieces = rec.pieces
        n_rows = sum(len(p) for p in pieces)
        done = False
        if rec.slot is None and self._csi_arr is not None and all(p.row_lo is not None for p in pieces):
            try:
                eng, slot, pieces, seg_start, seg_cell = self._submit_with_csi(self._engine_for(n_rows), pieces)
                done = True
            except ValueError:
                if self.use_cell_start_index:
                    raise
                import warnings

                warnings.warn("slaf_array._cell_start_index does not describe the expression rows; falling back to the cell column "
                              "(8 B/row over PCIe instead of 4)", RuntimeWarning, stacklevel=2)
                self._csi_arr = None
        if not done:
            if rec.slot is None:
                self._submit_staged(rec)
            eng, slot, pieces = rec.eng, rec.slot, rec.pieces
            try:
                seg_start, seg_cell = eng.segments(slot)
            except N.SlafB200Error as err:
                if err.code != N.ERR_CAPACITY:
                    raise
                eng.release(slot)                                   # more cells than the segment tables hold: size for the worst case
                self._unsubmit_ahead()
                eng = self._engine_for(n_rows, all_cells=True)
                slot, pieces = self._submit(eng, pieces)
                seg_start, seg_cell = eng.segments(slot)
            if not (seg_cell.size < 2 or bool(np.all(seg_cell[1:] > seg_cell[:-1]))) and len(np.unique(seg_cell)) != len(seg_cell):
                # a cell re-joined from two separated runs (partial + continuation with other chunks in
                # between): splice on the host like partition_by does, then resubmit (rare)
                eng.release(slot)
                self._unsubmit_ahead()
                slot, pieces = self._submit(eng, _regroup_segments(pieces, seg_start, seg_cell))
                seg_start, seg_cell = eng.segments(slot)
        seg_len = np.diff(seg_start.astype(np.int64))
        complete = self._complete_mask(seg_cell, seg_len, rec.flush)
        if rec.pred is not None:
            bad = np.flatnonzero(~complete)
            as_predicted = bad.size == len(rec.pred) and [(int(seg_cell[j]), int(seg_len[j])) for j in bad] == rec.pred
            if not as_predicted:
                rec.mispredicted = True
                self._unsubmit_ahead()
                keep = self.partial_cell_data
                self._stash_partials(pieces, seg_start, seg_cell, complete)
                rec.stash, true_stash = dict(self.partial_cell_data), self.partial_cell_data
                self.partial_cell_data = keep
                self._restage_ahead(true_stash, rec)
        else:
            self._stash_partials(pieces, seg_start, seg_cell, complete)
            rec.stash = dict(self.partial_cell_data)
        sel = np.flatnonzero(complete).astype(np.int32)
        return eng, slot, sel, seg_len

    def _load_raw_device(self, rec: _Staged) -> RawPrefetchBatch | None:
        """Raw mode on the device: the complete cells of the batch as CSR tensors in shuffled
        order, cut into ``batch_size``-cell chunks (views; the row pointers are rebased per chunk)."""
        t0 = time.time()
        eng, slot, sel, seg_len = self._front(rec)
        if sel.size == 0:
            eng.release(slot)
            return None
        order = sel[shuffle_perm(rec.seed, sel.size)]
        csr = eng.collect_csr(slot, perm=order)
        ends = np.concatenate(([0], np.cumsum(seg_len[order])))            # row offsets per output cell, known on the host
        chunks = []
        for lo in range(0, csr.n_cells, self.batch_size):
            hi = min(lo + self.batch_size, csr.n_cells)
            a, b = int(ends[lo]), int(ends[hi])
            chunks.append({"crow_indices": csr.crow_indices[lo:hi + 1] - a, "col_indices": csr.col_indices[a:b], "values": csr.values[a:b],
                           "cell_ids": csr.cell_ids[lo:hi]})
        process_time = time.time() - t0
        self._record_timing("shuffle", process_time)
        self._record_timing("total", time.time() - rec.start_time)
        self._record_timing("cells_processed", 0, csr.n_cells)
        self.total_prefetch_cells += csr.n_cells
        return RawPrefetchBatch(batch_id=rec.batch_id, batch_dfs=chunks, cell_integer_ids=csr.cell_ids.cpu().tolist(),
                                process_time=process_time, memory_mb=0.0)

    def _load_fused(self, rec: _Staged) -> TokenizedPrefetchBatch | None:
        t0 = time.time()
        eng, slot, sel, _seg_len = self._front(rec)
        shuffle_time = time.time() - t0
        if sel.size == 0:
            eng.release(slot)
            return None
        t1 = time.time()
        order = sel[shuffle_perm(rec.seed, sel.size)]              # random.seed(seed); random.shuffle(chunks), samplers.py:89-96
        # (capacity rounded up to a multiple of 512 cells: prefetch batches of nearly equal size then ask the caching allocator for
        # the SAME block sizes again and again instead of a new size per load -- new sizes are device-synchronising cudaMallocs)
        out = eng.collect(slot, perm=order, capacity_cells=(int(order.size) + 511) & ~511, **self._params())
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(eng.device))
        tokenize_time = time.time() - t1
        total = time.time() - rec.start_time
        self._record_timing("shuffle", shuffle_time)
        self._record_timing("window", 0.0)
        self._record_timing("tokenize", tokenize_time)
        self._record_timing("total", total)
        self._record_timing("cells_processed", 0, out.n_cells)
        self.total_prefetch_cells += out.n_cells
        return TokenizedPrefetchBatch(batch_id=rec.batch_id, input_ids=out.input_ids, attention_mask=out.attention_mask,
                                      values=out.values, cell_ids=out.cell_ids, partial_cell_data=dict(rec.stash or {}),
                                      tokenize_time=tokenize_time, epoch=rec.epoch, ready_event=ev)

    def _params(self) -> dict:
        return self.tokenizer.hot_path_params(use_binned_expressions=self.use_binned_expressions,
                                              n_expression_bins=self.n_expression_bins, window_kwargs=self.window_kwargs)

    # generic route: user-supplied Shuffle / Window / tokenizer objects, and raw mode -- the three
    # reference calls on host frames (still GPU-computed when the objects are this package's facades)
    def _host_complete(self, pieces: list[CooChunk]) -> CooChunk:
        rows = concat_chunks(pieces)
        starts, ends, ids = _runs(rows.cell)
        uniq, inv = np.unique(ids, return_inverse=True)
        observed = np.bincount(inv, weights=(ends - starts).astype(np.float64)).astype(np.int64)
        complete_u = self._complete_mask(uniq, observed)
        keep_run = complete_u[inv]
        self.partial_cell_data = {}
        for j in np.flatnonzero(~keep_run):
            lo, hi = int(starts[j]), int(ends[j])
            part = rows.slice(lo, hi)
            key = int(ids[j])
            self.partial_cell_data[key] = concat_chunks([self.partial_cell_data[key], part]) if key in self.partial_cell_data else \
                CooChunk(part.cell.copy(), part.gene.copy(), part.value.copy())
        if keep_run.all():
            return rows
        row_mask = np.repeat(keep_run, ends - starts)
        return CooChunk(rows.cell[row_mask], rows.gene[row_mask], rows.value[row_mask])

    def _load_on_host_frames(self, pieces: list[CooChunk], seed: int, start_time: float) -> PrefetchBatch | None:
        complete = self._host_complete(pieces)
        if len(complete) == 0:
            return None
        frame = complete.as_dict()
        if self.raw_mode:
            t0 = time.time()
            chunks = self.shuffle.apply(frame, seed, batch_size=self.batch_size)
            shuffle_time = time.time() - t0
            ids = np.unique(complete.cell)
            self._record_timing("shuffle", shuffle_time)
            self._record_timing("total", time.time() - start_time)
            self._record_timing("cells_processed", 0, len(ids))
            self.total_prefetch_cells += len(ids)
            return RawPrefetchBatch(batch_id=self.batch_id, batch_dfs=chunks, cell_integer_ids=ids.tolist(),
                                    process_time=shuffle_time, memory_mb=0.0)
        if self.tokenizer is None:
            raise RuntimeError("Tokenizer is required for tokenized mode")
        t0 = time.time()
        shuffled = self.shuffle.apply(frame, seed)
        if isinstance(shuffled, list):
            shuffled = {k: np.concatenate([np.asarray(c[k]) for c in shuffled]) for k in frame}
        shuffle_time = time.time() - t0
        t1 = time.time()
        params = {"n_expression_bins": self.n_expression_bins, "use_binned_expressions": self.use_binned_expressions}
        params.update(self.window_kwargs)
        grouped = self.tokenizer.window.apply(shuffled, SLAF_LANCE_COO_SCHEMA, self.tokenizer.max_genes, **params)
        window_time = time.time() - t1
        t2 = time.time()
        names = grouped.column_names if frames.frame_kind(grouped) == "arrow" else list(grouped.columns)
        ids, mask, vals = self.tokenizer.tokenize(
            gene_sequences=frames.list_column(grouped, "gene_sequence"),
            expr_sequences=frames.list_column(grouped, "expr_sequence") if "expr_sequence" in names else None)
        tokenize_time = time.time() - t2
        cells = [int(x) for x in frames.column(grouped, "cell_integer_id")]
        for k, v in (("shuffle", shuffle_time), ("window", window_time), ("tokenize", tokenize_time), ("total", time.time() - start_time)):
            self._record_timing(k, v)
        self._record_timing("cells_processed", 0, len(cells))
        self.total_prefetch_cells += len(cells)
        return TokenizedPrefetchBatch(batch_id=self.batch_id, input_ids=ids, attention_mask=mask, values=vals,
                                      cell_integer_ids=cells, partial_cell_data=dict(self.partial_cell_data),
                                      tokenize_time=tokenize_time, epoch=self.current_epoch)

    def close(self) -> None:
        with self._lock:
            self._close()

    def _close(self) -> None:
        try:
            self._drop_ahead()
        except Exception:
            pass
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        self._staging.close()
        if self._read_pool is not None:
            self._read_pool.shutdown(wait=False)
            self._read_pool = None


# ----------------------------------------------------------------------------------------------
# AsyncPrefetcher (reference: datasets.py:1105-1459)
# ----------------------------------------------------------------------------------------------
class AsyncPrefetcher:
    """Background thread that runs ``load_prefetch_batch`` ahead of the consumer.  The queue holds
    device-resident prefetch batches, so ``max_queue_size`` bounds HBM use (a cfg-2 prefetch batch of
    32 k cells is ~0.7 GB of token tensors); the producer blocks on a full queue."""

    def __init__(self, batch_processor: PrefetchBatchProcessor, max_queue_size: int = 500):
        if batch_processor is None:
            raise ValueError("batch_processor cannot be None")
        self.batch_processor = batch_processor
        self.max_queue_size = max_queue_size
        self.queue: queue.Queue = queue.Queue(maxsize=max_queue_size)
        self.worker_thread: threading.Thread | None = None
        self.should_stop = False
        self.finished = False
        self.error: BaseException | None = None
        self.total_cells_added = 0
        self.start_time: float | None = None
        self.total_tokenize_time = 0.0
        self.tokenize_count = 0
        self.current_epoch = 0

    def start(self) -> None:
        if self.worker_thread is None or not self.worker_thread.is_alive():
            self.should_stop = False
            self.finished = False
            self.start_time = time.time()
            self.worker_thread = threading.Thread(target=self._prefetch_worker, daemon=True)
            self.worker_thread.start()

    def stop(self) -> None:
        self.should_stop = True
        if self.worker_thread and self.worker_thread.is_alive():
            self.worker_thread.join(timeout=1.0)

    def _prefetch_worker(self) -> None:
        stream = None
        try:
            if torch.cuda.is_available():
                from ..runtime import resolve_device

                dev = resolve_device(self.batch_processor.device if self.batch_processor.device is not None
                                     else getattr(self.batch_processor.tokenizer, "device", None))
                torch.cuda.set_device(dev)
                stream = torch.cuda.Stream(dev)
            while not self.should_stop:
                try:
                    if stream is not None:
                        with torch.cuda.stream(stream):
                            batch = self.batch_processor.load_prefetch_batch()
                    else:
                        batch = self.batch_processor.load_prefetch_batch()
                except StopIteration:
                    break
                n = batch.n_cells if isinstance(batch, TokenizedPrefetchBatch) else len(batch.cell_integer_ids)
                self.total_cells_added += n
                self.total_tokenize_time += batch.tokenize_time if isinstance(batch, TokenizedPrefetchBatch) else batch.process_time
                self.tokenize_count += 1
                self.current_epoch = self.batch_processor.current_epoch
                while not self.should_stop:
                    try:
                        self.queue.put(batch, timeout=0.1)     # blocks: never drops a batch
                        break
                    except queue.Full:
                        continue
        except BaseException as err:  # surfaced to the consumer instead of being logged and lost
            self.error = err
        finally:
            self.finished = True

    def get_batch(self, timeout: float = 10.0) -> PrefetchBatch | None:
        deadline = time.time() + timeout
        while True:
            try:
                return self.queue.get(timeout=0.05)
            except queue.Empty:
                if self.error is not None:
                    raise RuntimeError(f"prefetch worker failed: {self.error!r}") from self.error
                if self.finished and self.queue.empty():
                    return None
                if time.time() > deadline:
                    return None

    def has_batch(self) -> bool:
        return not self.queue.empty()

    def is_done(self) -> bool:
        return self.finished and self.queue.empty()

    def get_stats(self) -> dict:
        elapsed = time.time() - self.start_time if self.start_time else 0.0
        return {"total_cells": self.total_cells_added, "elapsed_time": elapsed,
                "cells_per_sec": self.total_cells_added / elapsed if elapsed > 0 else 0.0,
                "queue_size": self.queue.qsize(), "queue_full": self.queue.full(),
                "current_epoch": self.current_epoch if not self.is_done() else self.batch_processor.n_epochs,
                "n_epochs": self.batch_processor.n_epochs}


# ----------------------------------------------------------------------------------------------
# SLAFIterableDataset (reference: datasets.py:1462-1938)
# ----------------------------------------------------------------------------------------------
class SLAFIterableDataset(IterableDataset):
    def __init__(self, slaf_array: Any, tokenizer: SLAFTokenizer | None, batch_size: int = 32, seed: int = 42,
                 max_queue_size: int = 8, pin_memory: bool = False, sampler_strategy: str = "sequential",
                 use_binned_expressions: bool = False, n_epochs: int = 1, raw_mode: bool = False, verbose: bool = True,
                 batches_per_chunk: int = 1, by_fragment: bool = True, use_mixture_of_scanners: bool = True, n_scanners: int = 16,
                 prefetch_batch_size: int = 4194304, parallelize_fragment_reads: bool = False,
                 prefetcher_ready_timeout: float = 10.0, device=None, output_device=None, device_raw: bool = False,
                 use_cell_start_index: bool | None = None):
        super().__init__()
        self.slaf_array = slaf_array
        self.tokenizer = tokenizer
        self.batch_size = batch_size
        self.seed = seed
        self.max_queue_size = max_queue_size
        self.pin_memory = pin_memory
        self.n_epochs = n_epochs
        self.raw_mode = raw_mode
        self.verbose = verbose
        self.by_fragment = by_fragment
        self.parallelize_fragment_reads = parallelize_fragment_reads
        self.output_device = torch.device(output_device) if output_device is not None else None
        from .samplers import ShuffleType, create_shuffle

        self.batch_processor = PrefetchBatchProcessor(
            slaf_array=slaf_array, shuffle=create_shuffle(ShuffleType.RANDOM), tokenizer=tokenizer, seed=seed,
            batches_per_chunk=batches_per_chunk, n_expression_bins=getattr(tokenizer, "n_expression_bins", 10),
            use_binned_expressions=use_binned_expressions, n_epochs=n_epochs, raw_mode=raw_mode, verbose=verbose,
            log_metrics=False, batch_size=batch_size, by_fragment=by_fragment, use_mixture_of_scanners=use_mixture_of_scanners,
            n_scanners=n_scanners, prefetch_batch_size=prefetch_batch_size, parallelize_fragment_reads=parallelize_fragment_reads,
            device=device, device_raw=device_raw, use_cell_start_index=use_cell_start_index)
        self.prefetcher = AsyncPrefetcher(self.batch_processor, max_queue_size=max_queue_size)
        self.prefetcher.start()
        self.device = self.output_device if self.output_device is not None else torch.device("cuda")

    def __iter__(self) -> Iterator[dict]:
        while True:
            data = self.prefetcher.get_batch(timeout=60.0)
            if data is None:
                break
            if isinstance(data, RawPrefetchBatch):
                for df in data.batch_dfs:
                    if isinstance(df, dict) and "crow_indices" in df:          # device raw mode: CSR tensors of the batch's cells
                        out = {"x": df, "cell_ids": df["cell_ids"]}
                        if self.n_epochs > 1:
                            out["epoch"] = self.prefetcher.current_epoch
                        yield out
                        continue
                    ids = frames.column(df, "cell_integer_id")
                    out = {"x": df, "cell_ids": [int(x) for x in ids[np.sort(np.unique(ids, return_index=True)[1])]]}
                    if self.n_epochs > 1:
                        out["epoch"] = self.prefetcher.current_epoch
                    yield out
                continue
            input_ids, attention_mask, values, cell_ids = data.input_ids, data.attention_mask, data.values, data.cell_ids
            if values is not None:                                   # datasets.py:1837-1847
                if values.shape != input_ids.shape:
                    raise ValueError("scGPT dual-stream contract violated: values and input_ids must "
                                     f"have identical shapes, got {tuple(values.shape)} vs {tuple(input_ids.shape)}")
                if values.shape != attention_mask.shape:
                    raise ValueError("scGPT dual-stream contract violated: values and attention_mask must "
                                     f"have identical shapes, got {tuple(values.shape)} vs {tuple(attention_mask.shape)}")
            if input_ids.is_cuda:
                cur = torch.cuda.current_stream(input_ids.device)
                if data.ready_event is not None:
                    cur.wait_event(data.ready_event)               # consumer stream waits; the host does not
                for t in (input_ids, attention_mask, values, cell_ids):
                    if t is not None:
                        t.record_stream(cur)
            if self.output_device is not None and self.output_device.type == "cpu":
                input_ids, attention_mask, cell_ids = input_ids.cpu(), attention_mask.cpu(), cell_ids.cpu()
                values = values.cpu() if values is not None else None
            # batch_size-cell views of the prefetch batch (datasets.py:1885-1912); split() makes all views of a tensor in one call
            if input_ids.shape[0] == 0:
                continue
            bs = self.batch_size
            epoch = data.epoch if self.n_epochs > 1 else None
            parts = [input_ids.split(bs), attention_mask.split(bs), cell_ids.split(bs)]
            if values is not None:
                parts.append(values.split(bs))
            for views in zip(*parts):
                batch = {"input_ids": views[0], "attention_mask": views[1], "cell_ids": views[2]}
                if values is not None:
                    batch["values"] = views[3]
                if epoch is not None:
                    batch["epoch"] = epoch
                yield batch

    def __del__(self):
        try:
            self.prefetcher.stop()
        except Exception:
            pass
