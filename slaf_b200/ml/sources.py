"""Expression-table sources for the host feeder.

The reference reads ``lance.dataset(f"{slaf_path}/expression.lance")`` and uses exactly three calls of
it on this path: ``get_fragments()``, ``fragment.to_batches(batch_size=...)`` / ``fragment.to_table()``
and ``dataset.to_batches()`` (slaf/ml/datasets.py:378-398,797-798,829-830).  ``ExpressionSource``
is that interface reduced to column arrays; chunks are ``CooChunk`` = three host numpy arrays in
the storage dtypes (zero-copy views of Arrow buffers where the source is Arrow).

    ArrayExpressionSource   in-memory columns cut into fragments of ``max_rows_per_file`` rows
                            (not aligned to cells, like the converter's files, converter.py:121)
    ArrowExpressionSource   a directory of Arrow IPC files (one per fragment), memory mapped
    LanceExpressionSource   a real ``expression.lance`` (only when pylance is importable)

Lance decode, object-store I/O and SQL stay on the host and are not part of the accelerated path.
"""

from __future__ import annotations

import glob
import os
from dataclasses import dataclass
from typing import Any, Iterator

import numpy as np

from . import frames

COLUMNS = ("cell_integer_id", "gene_integer_id", "value")


@dataclass
class CooChunk:
    cell: np.ndarray
    gene: np.ndarray
    value: np.ndarray
    row_lo: int | None = None     # absolute row of the first row in the expression table (when the source knows it)

    def __len__(self) -> int:
        return int(self.cell.shape[0])

    def slice(self, lo: int, hi: int) -> "CooChunk":
        return CooChunk(self.cell[lo:hi], self.gene[lo:hi], self.value[lo:hi], None if self.row_lo is None else self.row_lo + lo)

    def copy(self) -> "CooChunk":
        return CooChunk(self.cell.copy(), self.gene.copy(), self.value.copy(), self.row_lo)

    def as_dict(self) -> dict[str, np.ndarray]:
        return {"cell_integer_id": self.cell, "gene_integer_id": self.gene, "value": self.value}


def concat_chunks(chunks: list[CooChunk]) -> CooChunk:
    if len(chunks) == 1:
        return chunks[0]
    row_lo, pos = chunks[0].row_lo, chunks[0].row_lo
    for c in chunks:                                     # the result keeps its table position only if the parts are consecutive rows
        if pos is None or c.row_lo != pos:
            row_lo = None
            break
        pos += len(c)
    return CooChunk(np.concatenate([c.cell for c in chunks]), np.concatenate([c.gene for c in chunks]),
                    np.concatenate([c.value for c in chunks]), row_lo)


def chunk_from_frame(frame: Any) -> CooChunk:
    """Arrow RecordBatch / Table, polars or pandas frame, or dict -> CooChunk in kernel dtypes."""
    cols = frames.extract(frame, _SCHEMA)
    return CooChunk(cols.cell, cols.gene, cols.value)


class _Schema:
    group_key, item_key, value_key = COLUMNS


_SCHEMA = _Schema()


class Fragment:
    """One storage fragment: a contiguous row range of the expression table."""

    def count_rows(self) -> int:
        raise NotImplementedError

    def to_table(self) -> CooChunk:
        raise NotImplementedError

    def to_pieces(self) -> list[CooChunk]:
        """The fragment's rows as consecutive blocks in storage order, WITHOUT gathering them into one array: what the feeder
        hands to the library (which copies the blocks back to back into its staging columns).  Sources whose storage is
        chunked (Arrow record batches, Lance pages) return views of the chunks here; ``to_table`` has to concatenate them."""
        return [self.to_table()]

    def to_batches(self, batch_size: int = 8192) -> Iterator[CooChunk]:
        table = self.to_table()
        for lo in range(0, len(table), batch_size):
            yield table.slice(lo, min(lo + batch_size, len(table)))


class ExpressionSource:
    def get_fragments(self) -> list[Fragment]:
        raise NotImplementedError

    def count_rows(self) -> int:
        return sum(f.count_rows() for f in self.get_fragments())

    def to_batches(self, batch_size: int = 8192) -> Iterator[CooChunk]:
        """Sequential scan of the whole table (``dataset.to_batches()``, datasets.py:399)."""
        for frag in self.get_fragments():
            yield from frag.to_batches(batch_size)


# ------------------------------------------------------------------------------ in-memory
class _ArrayFragment(Fragment):
    def __init__(self, chunk: CooChunk):
        self._chunk = chunk

    def count_rows(self) -> int:
        return len(self._chunk)

    def to_table(self) -> CooChunk:
        return self._chunk


class ArrayExpressionSource(ExpressionSource):
    def __init__(self, cell: np.ndarray, gene: np.ndarray, value: np.ndarray, max_rows_per_file: int = 10_000_000):
        if not (len(cell) == len(gene) == len(value)):
            raise ValueError("column lengths differ")
        if max_rows_per_file < 1:
            raise ValueError("max_rows_per_file must be >= 1")
        whole = CooChunk(np.ascontiguousarray(cell), np.ascontiguousarray(gene), np.ascontiguousarray(value), 0)
        self._fragments = [_ArrayFragment(whole.slice(lo, min(lo + max_rows_per_file, len(whole))))
                           for lo in range(0, max(len(whole), 1), max_rows_per_file)]
        self._whole = whole
        self.pinned = False

    def get_fragments(self) -> list[Fragment]:
        return list(self._fragments)

    def pin(self) -> bool:
        """Page-lock the column arrays (cudaHostRegister): every chunk then reaches the GPU by DMA
        straight from these arrays, the feeder copies nothing on the host."""
        from ..engine import register_host_memory

        self.pinned = all(register_host_memory(a) for a in (self._whole.cell, self._whole.gene, self._whole.value))
        return self.pinned

    def unpin(self) -> None:
        from ..engine import unregister_host_memory

        for a in (self._whole.cell, self._whole.gene, self._whole.value):
            unregister_host_memory(a)
        self.pinned = False


class PinnedCacheSource(ExpressionSource):
    """The rows of another source copied ONCE into page-locked host memory (``cudaHostAlloc``), fragment structure and table
    positions kept: every later epoch then reaches the GPU by DMA straight from these arrays, with no host copy per load.  This is
    the resident-dataset mode for sources whose own memory cannot be page-locked (the driver refuses to register read-only file
    mappings on many hosts) and that fit in RAM -- typically one rank's shard.  ``pin_cell=False`` keeps the cell column in
    ordinary memory (the cell-start-index route never ships it; it is only probed on the host)."""

    def __init__(self, source: ExpressionSource, pin_cell: bool = False, threads: int = 4):
        from concurrent.futures import ThreadPoolExecutor

        from ..engine import pinned_empty

        frags = source.get_fragments()
        pieces_per_frag = [f.to_pieces() for f in frags]
        rows = [sum(len(p) for p in ps) for ps in pieces_per_frag]
        total = sum(rows)
        first = next((p for ps in pieces_per_frag for p in ps if len(p)), None)
        if first is None:
            self._fragments, self.rows, self.pinned = [], 0, True
            return
        self._cell = pinned_empty(total, first.cell.dtype) if pin_cell else np.empty(total, first.cell.dtype)
        self._gene = pinned_empty(total, first.gene.dtype)
        self._value = pinned_empty(total, first.value.dtype)
        tasks, off = [], 0
        self._fragments = []
        for ps, n in zip(pieces_per_frag, rows):
            row_lo = next((p.row_lo for p in ps if len(p)), None)
            self._fragments.append(_ArrayFragment(CooChunk(self._cell[off:off + n], self._gene[off:off + n], self._value[off:off + n], row_lo)))
            pos = off
            for p in ps:
                for lo in range(0, len(p), 1 << 21):
                    hi = min(lo + (1 << 21), len(p))
                    tasks.append((pos + lo, p, lo, hi))
                pos += len(p)
            off += n

        def copy(t):
            o, p, lo, hi = t
            np.copyto(self._cell[o:o + hi - lo], p.cell[lo:hi])
            np.copyto(self._gene[o:o + hi - lo], p.gene[lo:hi])
            np.copyto(self._value[o:o + hi - lo], p.value[lo:hi])

        if threads > 1 and len(tasks) > 1:
            with ThreadPoolExecutor(max_workers=threads) as pool:
                list(pool.map(copy, tasks))
        else:
            for t in tasks:
                copy(t)
        self.rows, self.pinned = total, True

    def get_fragments(self) -> list[Fragment]:
        return list(self._fragments)

    def nbytes_pinned(self) -> int:
        return 0 if not self._fragments else int(self._gene.nbytes + self._value.nbytes)


# ------------------------------------------------------------------------------ Arrow IPC directory
class _ArrowFileFragment(Fragment):
    def __init__(self, path: str):
        self.path = path
        self._rows = None
        self._map = None
        self._pinned: tuple[int, int] | None = None
        self.row_lo: int | None = None      # set by the source: rows of the preceding fragments

    def _open(self):
        import pyarrow as pa

        if self._map is None:                   # one mapping per fragment for its lifetime: the views handed out keep pointing
            self._map = pa.memory_map(self.path, "r")      # into it, and pin() page-locks exactly this address range
        return pa.ipc.open_file(self._map)

    def pin(self) -> bool:
        """Page-lock the file's mapping (cudaHostRegister, read-only), so the record-batch views are DMA sources and the feeder
        copies nothing on the host.  Needs the file to fit in RAM; False when the driver refuses, the staged path then remains."""
        from ..engine import register_host_range

        self._open()
        buf = self._map.read_buffer()           # zero-copy view of the whole mapping
        self._map.seek(0)
        self._pinned = (buf.address, buf.size) if register_host_range(buf.address, buf.size) else None
        return self._pinned is not None

    def unpin(self) -> None:
        from ..engine import unregister_host_range

        if self._pinned is not None:
            unregister_host_range(self._pinned[0])
            self._pinned = None

    def count_rows(self) -> int:
        if self._rows is None:
            rd = self._open()
            self._rows = sum(rd.get_batch(i).num_rows for i in range(rd.num_record_batches))
        return self._rows

    def to_table(self) -> CooChunk:
        out = chunk_from_frame(self._open().read_all().combine_chunks())
        out.row_lo = self.row_lo
        return out

    def to_pieces(self) -> list[CooChunk]:
        rd = self._open()
        out, pos = [], 0
        for i in range(rd.num_record_batches):       # zero-copy views of the memory map, one per record batch
            ch = chunk_from_frame(rd.get_batch(i))
            ch.row_lo = None if self.row_lo is None else self.row_lo + pos
            pos += len(ch)
            if len(ch):
                out.append(ch)
        return out

    def to_batches(self, batch_size: int = 8192) -> Iterator[CooChunk]:
        rd = self._open()
        pos = 0
        for i in range(rd.num_record_batches):
            rb = rd.get_batch(i)
            for lo in range(0, rb.num_rows, batch_size):
                out = chunk_from_frame(rb.slice(lo, min(batch_size, rb.num_rows - lo)))
                out.row_lo = None if self.row_lo is None else self.row_lo + pos + lo
                yield out
            pos += rb.num_rows


class ArrowExpressionSource(ExpressionSource):
    """``<path>/fragment-*.arrow`` (Arrow IPC file format, uncompressed -> zero-copy memory maps)."""

    def __init__(self, path: str):
        self.path = path
        files = sorted(glob.glob(os.path.join(path, "fragment-*.arrow")))
        if not files:
            raise FileNotFoundError(f"no fragment-*.arrow files under {path}")
        self._fragments = [_ArrowFileFragment(f) for f in files]
        pos = 0
        for f in self._fragments:            # fragments are consecutive row ranges of the table, in file order
            f.row_lo = pos
            pos += f.count_rows()

    def get_fragments(self) -> list[Fragment]:
        return list(self._fragments)

    def pin(self) -> bool:
        """Page-lock every fragment's mapping (see ``_ArrowFileFragment.pin``); True when all of them were accepted."""
        ok = [f.pin() for f in self._fragments]
        self.pinned = all(ok)
        return self.pinned

    def unpin(self) -> None:
        for f in self._fragments:
            f.unpin()
        self.pinned = False


def write_arrow_dataset(path: str, cell: np.ndarray, gene: np.ndarray, value: np.ndarray, max_rows_per_file: int,
                        rows_per_batch: int = 1 << 20) -> list[str]:
    """Write COO columns as an ``ArrowExpressionSource`` directory (test / benchmark fixture writer)."""
    import pyarrow as pa

    os.makedirs(path, exist_ok=True)
    out = []
    for i, lo in enumerate(range(0, max(len(cell), 1), max_rows_per_file)):
        hi = min(lo + max_rows_per_file, len(cell))
        table = pa.table({"cell_integer_id": cell[lo:hi], "gene_integer_id": gene[lo:hi], "value": value[lo:hi]})
        name = os.path.join(path, f"fragment-{i:06d}.arrow")
        with pa.OSFile(name, "wb") as sink, pa.ipc.new_file(sink, table.schema) as writer:
            for b in table.to_batches(max_chunksize=rows_per_batch):
                writer.write_batch(b)
        out.append(name)
    return out


# ------------------------------------------------------------------------------ Lance (optional)
class _LanceFragment(Fragment):
    def __init__(self, frag: Any, row_lo: int | None = None):
        self._frag = frag
        self.row_lo = row_lo

    def count_rows(self) -> int:
        return int(self._frag.count_rows())

    def to_table(self) -> CooChunk:
        out = chunk_from_frame(self._frag.to_table(columns=list(COLUMNS)).combine_chunks())
        out.row_lo = self.row_lo
        return out

    def to_pieces(self) -> list[CooChunk]:
        out, pos = [], 0
        for rb in self._frag.to_table(columns=list(COLUMNS)).to_batches():      # the decoded chunks as they are, no concatenation
            ch = chunk_from_frame(rb)
            ch.row_lo = None if self.row_lo is None else self.row_lo + pos
            pos += len(ch)
            if len(ch):
                out.append(ch)
        return out

    def to_batches(self, batch_size: int = 8192) -> Iterator[CooChunk]:
        pos = 0
        for rb in self._frag.to_batches(batch_size=batch_size, columns=list(COLUMNS)):
            out = chunk_from_frame(rb)
            out.row_lo = None if self.row_lo is None else self.row_lo + pos
            pos += len(out)
            yield out


class LanceExpressionSource(ExpressionSource):
    def __init__(self, path_or_dataset: Any):
        if isinstance(path_or_dataset, (str, os.PathLike)):
            try:
                import lance
            except ImportError as err:  # pragma: no cover - pylance is not in the build image
                raise ImportError("reading expression.lance needs the `pylance` package") from err
            self._ds = lance.dataset(str(path_or_dataset))
        else:
            self._ds = path_or_dataset

    def get_fragments(self) -> list[Fragment]:
        out, pos = [], 0
        for f in self._ds.get_fragments():       # dataset order = row order (the converter appends cell-major)
            out.append(_LanceFragment(f, pos))
            pos += int(f.count_rows())
        return out

    def count_rows(self) -> int:
        return int(self._ds.count_rows())


def open_expression(slaf_array: Any) -> ExpressionSource:
    """The expression table of a SLAF array-like object:
    an ``ExpressionSource`` in ``.expression``; a Lance dataset object in ``.expression``;
    else ``<slaf_path>/expression.lance`` (Lance) or ``<slaf_path>/expression.arrow/`` (Arrow IPC)."""
    if isinstance(slaf_array, ExpressionSource):
        return slaf_array
    expr = getattr(slaf_array, "expression", None)
    if isinstance(expr, ExpressionSource):
        return expr
    if expr is not None and hasattr(expr, "get_fragments") and type(expr).__module__.split(".")[0] == "lance":
        return LanceExpressionSource(expr)
    path = getattr(slaf_array, "slaf_path", None) if not isinstance(slaf_array, (str, os.PathLike)) else slaf_array
    if path is None:
        raise TypeError("slaf_array needs an `.expression` ExpressionSource or a `.slaf_path`")
    path = str(path)
    arrow_dir = os.path.join(path, "expression.arrow")
    if os.path.isdir(arrow_dir):
        return ArrowExpressionSource(arrow_dir)
    return LanceExpressionSource(os.path.join(path, "expression.lance"))


# ------------------------------------------------------------------------------ array-like stand-in
class SyntheticSLAFArray:
    """The slice of ``SLAFArray`` the hot path touches (slaf/core/slaf.py:295,481-492,645-702):
    ``slaf_path``, ``shape``, ``expression``, ``_cell_start_index`` (row offsets per cell,
    int64[n_cells + 1]) and gene metadata for the tokenizer vocabulary."""

    def __init__(self, source: ExpressionSource, cell_start_index: np.ndarray, n_genes: int, slaf_path: str | None = None,
                 gene_ids: list[str] | None = None):
        self.expression = source
        self._cell_start_index = np.asarray(cell_start_index, np.int64)
        self.shape = (len(self._cell_start_index) - 1, int(n_genes))
        self.slaf_path = slaf_path
        ids = gene_ids if gene_ids is not None else [f"gene_{i}" for i in range(int(n_genes))]
        self.var = {"gene_id": np.asarray(ids, dtype=object), "gene_integer_id": np.arange(int(n_genes), dtype=np.int64)}

    def is_metadata_ready(self) -> bool:
        return True

    @classmethod
    def from_batch(cls, batch: Any, n_genes: int, max_rows_per_file: int = 10_000_000) -> "SyntheticSLAFArray":
        src = ArrayExpressionSource(batch.cell, batch.gene, batch.value, max_rows_per_file)
        first = int(batch.cell[0]) if batch.n_rows else 0
        csi = np.zeros(first + batch.n_cells + 1, np.int64)
        csi[first:] = batch.cell_start
        return cls(src, csi, n_genes)
