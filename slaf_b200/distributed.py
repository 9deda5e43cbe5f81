"""Multi-GPU plumbing of the hot path: one process per GPU, cells sharded by fragment range, no
collective on the per-batch path.

* ``assign_fragments`` -- which fragments a rank reads.  ``shuffle=True`` reproduces the reference's
  ``Coordinator.assign_partitions`` (slaf/distributed/coordinator.py:40-80: seeded CPython shuffle
  of the fragment indices, contiguous split, remainder to the first workers); the default keeps the
  ranges contiguous so that at most ``world_size - 1`` cells straddle a shard boundary.
* ``shard_rows`` / ``RowRangeSource`` -- cell-aligned row range of a rank: a cell belongs to the shard
  that holds its FIRST row, so every cell is tokenized by exactly one rank and no rows are exchanged.
* ``allreduce_gene_stats`` -- the one cross-GPU step (BASELINE.json config 4, an extension: the
  reference has no per-gene normalisation): per-gene nonzero ``sum`` / ``count`` vectors accumulated by
  the ``gene_stats`` kernel directly into one ``[2, G]`` int64 buffer and summed over ranks
  (NCCL over NVLink/NVSwitch on GPU tensors; any ``torch.distributed`` backend works, the CPU tests use
  gloo).  Integer sums: the result is identical for any rank count and reduction order.
"""

from __future__ import annotations

from typing import Any, Iterator

import numpy as np
import torch

from .engine import shuffle_perm
from .ml.sources import CooChunk, ExpressionSource, Fragment


def assign_fragments(n_fragments: int, world_size: int, seed: int = 42, shuffle: bool = False) -> list[list[int]]:
    """Fragment indices per rank (rank r gets ``result[r]``)."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if n_fragments < 0:
        raise ValueError("n_fragments must be >= 0")
    order = shuffle_perm(seed, n_fragments).tolist() if shuffle else list(range(n_fragments))
    base, rem = divmod(n_fragments, world_size)
    out, start = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append(order[start:start + n])
        start += n
    return out


def shard_rows(fragment_rows: list[int], cell_start_index: np.ndarray, world_size: int, rank: int) -> tuple[int, int]:
    """Cell-aligned row range ``[lo, hi)`` of ``rank`` for contiguous fragment ranges: the fragment
    range's row range, with both ends moved up to the next cell start."""
    parts = assign_fragments(len(fragment_rows), world_size)
    offs = np.concatenate(([0], np.cumsum(np.asarray(fragment_rows, np.int64))))
    mine = parts[rank]
    if not mine:
        return 0, 0
    csi = np.asarray(cell_start_index, np.int64)

    def up(x: int) -> int:
        return int(csi[np.searchsorted(csi, x, side="left")]) if x < csi[-1] else int(csi[-1])

    return up(int(offs[mine[0]])), up(int(offs[mine[-1] + 1]))


class _SlicedFragment(Fragment):
    def __init__(self, frag: Fragment, lo: int, hi: int):
        self._frag, self._lo, self._hi = frag, lo, hi

    def count_rows(self) -> int:
        return self._hi - self._lo

    def to_table(self) -> CooChunk:
        return self._frag.to_table().slice(self._lo, self._hi)

    def to_pieces(self) -> list[CooChunk]:
        out, pos = [], 0
        for piece in self._frag.to_pieces():
            lo, hi = max(self._lo - pos, 0), min(self._hi - pos, len(piece))
            if hi > lo:
                out.append(piece.slice(lo, hi))
            pos += len(piece)
        return out

    def to_batches(self, batch_size: int = 8192) -> Iterator[CooChunk]:
        pos = 0
        for chunk in self._frag.to_batches(batch_size):
            lo, hi = max(self._lo - pos, 0), min(self._hi - pos, len(chunk))
            if hi > lo:
                yield chunk.slice(lo, hi)
            pos += len(chunk)
            if pos >= self._hi:
                break


class RowRangeSource(ExpressionSource):
    """The rows ``[lo, hi)`` of another source, fragment structure kept (fragments are trimmed)."""

    def __init__(self, source: ExpressionSource, lo: int, hi: int):
        self._fragments: list[Fragment] = []
        pos = 0
        for frag in source.get_fragments():
            n = frag.count_rows()
            a, b = max(lo - pos, 0), min(hi - pos, n)
            if b > a:
                self._fragments.append(frag if (a == 0 and b == n) else _SlicedFragment(frag, a, b))
            pos += n

    def get_fragments(self) -> list[Fragment]:
        return list(self._fragments)


def shard_source(source: ExpressionSource, cell_start_index: np.ndarray, world_size: int, rank: int) -> RowRangeSource:
    rows = [f.count_rows() for f in source.get_fragments()]
    lo, hi = shard_rows(rows, cell_start_index, world_size, rank)
    return RowRangeSource(source, lo, hi)


# ------------------------------------------------------------------------------ per-gene statistics
def allreduce_gene_stats(sum_cnt: torch.Tensor, group: Any = None) -> torch.Tensor:
    """Sum the ``[2, G]`` int64 (sum, count) buffer over all ranks in place."""
    import torch.distributed as dist

    if sum_cnt.dtype != torch.int64 or sum_cnt.dim() != 2 or sum_cnt.shape[0] != 2:
        raise ValueError("expected an int64 tensor of shape [2, n_genes]")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sum_cnt, op=dist.ReduceOp.SUM, group=group)
    return sum_cnt


def gene_stats(engine: Any, batches, n_genes: int, group: Any = None, mark: Any = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-gene nonzero sum and count over this rank's batches (device column triples or pairs
    ``(gene, value)``), all-reduced over ``group``.  Returns int64 tensors ``(sum, count)`` of length G.
    ``mark``: a CUDA event recorded between the kernels and the all-reduce (bench.py times the two halves)."""
    buf = torch.zeros((2, n_genes), dtype=torch.int64, device=engine.device)
    for cols in batches:
        gene, value = cols[-2], cols[-1]
        engine.gene_stats(gene, value, n_genes, sum_out=buf[0], cnt_out=buf[1])     # the kernel accumulates into the reduce buffer
    if mark is not None:
        mark.record()
    allreduce_gene_stats(buf, group)
    return buf[0], buf[1]


def gene_medians(engine: Any, batches, n_genes: int, n_bins: int = 1024, group: Any = None):
    """Exact per-gene nonzero medians over ALL ranks' rows: per-rank histograms ([G, n_bins] int32, accumulated by
    the kernel straight into the reduce buffer), one all-reduce, then the median kernel.  Returns
    (median float32 [G], exact bool [G]); exact is False where a middle value is >= n_bins - 1."""
    import torch.distributed as dist

    hist = torch.zeros((n_genes, n_bins), dtype=torch.int32, device=engine.device)
    for cols in batches:
        engine.gene_hist(cols[-2], cols[-1], n_genes, n_bins, hist_out=hist)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return engine.gene_median(hist)


def normalisation_vector(total: torch.Tensor, count: torch.Tensor) -> torch.Tensor:
    """Per-gene mean of the nonzero values as float32 (1.0 for genes never seen); f64 division after
    the integer reduce, so every rank computes the bit-identical vector."""
    mean = total.to(torch.float64) / count.clamp(min=1).to(torch.float64)
    return torch.where(count > 0, mean, torch.ones_like(mean)).to(torch.float32)


# ------------------------------------------------------------------------------ names of the reference's slaf.distributed
class PartitionAssignment:
    """reference: slaf/distributed/coordinator.py:14-19."""

    def __init__(self, worker_id: str, partition_indices: list[int]):
        self.worker_id = worker_id
        self.partition_indices = partition_indices


class Coordinator:
    """reference: slaf/distributed/coordinator.py:22-80 -- seeded shuffle of the partition (fragment) indices, contiguous
    split, remainder to the first workers.  ``data_source`` needs ``get_partition_count()`` (or is an ExpressionSource)."""

    def __init__(self, data_source: Any, n_workers: int):
        self.data_source = data_source
        self.n_workers = n_workers
        self.total_partitions = (data_source.get_partition_count() if hasattr(data_source, "get_partition_count")
                                 else len(data_source.get_fragments()))

    def assign_partitions(self, seed: int = 42) -> dict[str, PartitionAssignment]:
        parts = assign_fragments(self.total_partitions, self.n_workers, seed=seed, shuffle=True)
        return {f"worker_{i}": PartitionAssignment(f"worker_{i}", p) for i, p in enumerate(parts)}
