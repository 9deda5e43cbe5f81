"""Shuffle strategies: mirrors ``slaf.ml.samplers`` (slaf/ml/samplers.py:22-221) and the generic
``slaf.distributed.shuffle.Shuffle`` (slaf/distributed/shuffle.py:18-52).

The block shuffle is control-plane work: the order of the cell blocks depends only on
(seed, number of cells) and is produced by the library's CPython-exact host code
(``slafb_shuffle_perm``).  In the fused dataloader path it is never materialised as a re-ordered
frame -- the kernels apply it as an indirection; these classes exist so code written against the
reference's Shuffle interface keeps working.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from typing import Any, Literal

import numpy as np

from ..engine import shuffle_perm
from . import frames


class ShuffleType(str, Enum):
    RANDOM = "random"
    STRATIFIED = "stratified"


class Shuffle(ABC):
    @abstractmethod
    def apply(self, df: Any, seed: int, batch_size: int | None = None, **kwargs: Any) -> Any:
        raise NotImplementedError


def _blocks(cell: np.ndarray):
    """(row order that groups cells by first appearance, block starts [C+1] in that order)."""
    _, first, inv = np.unique(cell, return_index=True, return_inverse=True)
    rank = np.empty(len(first), np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(first))
    g = rank[inv]
    order = np.argsort(g, kind="stable")
    counts = np.bincount(g, minlength=len(first))
    starts = np.zeros(len(first) + 1, np.int64)
    np.cumsum(counts, out=starts[1:])
    return order, starts


class RandomShuffle(Shuffle):
    """reference: slaf/ml/samplers.py:54-106."""

    group_column = "cell_integer_id"

    def apply(self, df: Any, seed: int, batch_size: int | None = None, **kwargs: Any) -> Any:
        n = frames.frame_len(df)
        if n == 0:
            return [] if batch_size is not None else df
        cell = frames.column(df, kwargs.get("group_key", self.group_column))
        order, starts = _blocks(cell)
        n_cells = len(starts) - 1
        perm = shuffle_perm(seed, n_cells)                       # random.seed(seed); random.shuffle(chunks)
        lens = (starts[1:] - starts[:-1])[perm]
        out_starts = np.zeros(n_cells + 1, np.int64)
        np.cumsum(lens, out=out_starts[1:])
        src = np.repeat(starts[:-1][perm] - out_starts[:-1], lens) + np.arange(n, dtype=np.int64)
        rows = order[src]
        shuffled = frames.take_rows(df, rows)
        if batch_size is None:
            return shuffled
        return [frames.slice_rows(shuffled, int(out_starts[i]), int(out_starts[min(i + batch_size, n_cells)]))
                for i in range(0, n_cells, batch_size)]


class StratifiedShuffle(Shuffle):
    """reference: slaf/ml/samplers.py:109-191.  Without a cell-type column the reference falls back
    to RandomShuffle (:140-142); that fallback is what the hot path ever sees (expression rows
    carry no cell type).  Stratification proper is host-side metadata work and out of scope."""

    def apply(self, df: Any, seed: int, batch_size: int | None = None, **kwargs: Any) -> Any:
        name = kwargs.get("cell_type_column", "cell_type")
        kind = frames.frame_kind(df)
        names = df.column_names if kind == "arrow" else (list(df.keys()) if kind == "dict" else list(df.columns))
        if frames.frame_len(df) == 0 or name not in names:
            return RandomShuffle().apply(df, seed, batch_size, **kwargs)
        raise NotImplementedError("stratified shuffling needs per-cell metadata; outside the B200 hot path (DESIGN.md, out of scope)")


class GenericShuffle:
    """reference: slaf/distributed/shuffle.py:18-52 -- apply(df, schema, seed)."""

    def apply(self, df: Any, schema: Any, seed: int, **kwargs: Any) -> Any:
        if frames.frame_len(df) == 0:
            return df
        return RandomShuffle().apply(df, seed, None, group_key=schema.group_key)


def create_shuffle(shuffle_type: ShuffleType | Literal["random", "stratified"]) -> Shuffle:
    """reference: slaf/ml/samplers.py:195-221."""
    if isinstance(shuffle_type, str):
        try:
            shuffle_type = ShuffleType(shuffle_type.lower())
        except ValueError as err:
            raise ValueError(f"Unsupported shuffle type: {shuffle_type}") from err
    if shuffle_type == ShuffleType.RANDOM:
        return RandomShuffle()
    if shuffle_type == ShuffleType.STRATIFIED:
        return StratifiedShuffle()
    raise ValueError(f"Unsupported shuffle type: {shuffle_type}")
