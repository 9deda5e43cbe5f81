"""Window functions on the B200: same classes and call signature as the reference's
``slaf.ml.aggregators`` (slaf/ml/aggregators.py:17-277) and the generic
``slaf.distributed.window.Window`` (slaf/distributed/window.py:14-59).

    window.apply(fragment_df, schema, max_items, **kwargs) -> grouped frame

Per group (cell): dense rank of the value (descending), keep rows with rank <= max_items, list
the items (and values / log1p bins) of the kept rows IN ROW ORDER; groups in order of first
appearance.  The arithmetic runs in the CUDA kernels behind ``HotPathEngine.window`` (K1 CSR build
+ general per-cell kernel); this module only adapts frames.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from typing import Any

import numpy as np

from .. import _native as N
from .. import runtime
from ..engine import regroup_contiguous
from . import frames


class Window(ABC):
    """Base class (reference: slaf/ml/aggregators.py:17-50; protocol slaf/distributed/processor.py:32-56)."""

    device = None  # None = current CUDA device

    @abstractmethod
    def apply(self, fragment_df: Any, schema: Any, max_items: int, **kwargs: Any) -> Any:
        raise NotImplementedError

    # shared implementation ------------------------------------------------------------------
    def _run(self, fragment_df: Any, schema: Any, max_items: int, mode: str, with_values: bool, n_bins: int = 10,
             min_percentile: float | None = None, maintain_order: bool = True) -> Any:
        if max_items < 1:
            raise ValueError(f"max_items must be >= 1, got {max_items}")
        cols = frames.extract(fragment_df, schema)
        if cols.n_rows == 0:
            return frames.build_grouped(cols, schema, np.zeros(0, np.int64), np.zeros(1, np.int64), np.zeros(0, np.int64),
                                        np.zeros(0, np.float64) if with_values else None, mode == "scgpt_binned")
        cell, gene, value = cols.cell, cols.gene, cols.value
        kw = dict(mode=mode, max_items=int(max_items), n_bins=int(n_bins), min_percentile=min_percentile, check_contiguous=True)
        with runtime.session(cols.n_rows, device=self.device) as eng:
            try:
                res = eng.window(cell, gene, value, **kw)
            except N.SlafB200Error as err:
                if err.code != N.ERR_NONCONTIGUOUS:
                    raise
                cell, gene, value = regroup_contiguous(cell, gene, value)   # partition_by semantics for split cells
                res = eng.window(cell, gene, value, **kw)
        exprs = res.exprs.cpu().numpy() if (with_values and res.exprs is not None) else None
        return frames.build_grouped(cols, schema, res.cell_ids.cpu().numpy(), res.offsets.cpu().numpy(),
                                    res.genes.cpu().numpy(), exprs, mode == "scgpt_binned")


class ScGPTWindow(Window):
    """reference: slaf/ml/aggregators.py:53-136.  kwargs: n_expression_bins (10),
    use_binned_expressions (True)."""

    def apply(self, fragment_df: Any, schema: Any, max_items: int, **kwargs: Any) -> Any:
        n_bins = kwargs.get("n_expression_bins", 10)
        binned = kwargs.get("use_binned_expressions", True)
        return self._run(fragment_df, schema, max_items, "scgpt_binned" if binned else "scgpt_raw", True, n_bins=n_bins)


class GeneformerWindow(Window):
    """reference: slaf/ml/aggregators.py:139-216.  kwargs: min_percentile (None | 0..100)."""

    def apply(self, fragment_df: Any, schema: Any, max_items: int, **kwargs: Any) -> Any:
        pct = kwargs.get("min_percentile", None)
        if pct is not None and not (0.0 <= float(pct) <= 100.0):
            raise ValueError("min_percentile must be within [0, 100]")
        return self._run(fragment_df, schema, max_items, "geneformer_pct" if pct is not None else "geneformer", False,
                         min_percentile=pct)


class SimpleWindow(Window):
    """reference: slaf/ml/aggregators.py:219-256 (rank filter, raw value lists)."""

    def apply(self, fragment_df: Any, schema: Any, max_items: int, **kwargs: Any) -> Any:
        return self._run(fragment_df, schema, max_items, "scgpt_raw", True)


class GenericWindow(Window):
    """reference: slaf/distributed/window.py:14-59 -- the schema decides whether a value list is
    produced; kwargs are ignored.  (The reference does not pass maintain_order here, so its group
    order is unspecified; first-appearance order is one valid outcome.)"""

    def apply(self, df: Any, schema: Any, max_items: int, **kwargs: Any) -> Any:
        with_values = bool(getattr(schema, "value_list_key", None))
        return self._run(df, schema, max_items, "scgpt_raw" if with_values else "geneformer", with_values)


class WindowType(str, Enum):
    GENEFORMER = "geneformer"
    SCPGPT = "scgpt"
    SIMPLE = "simple"


def create_window(window_type: WindowType | str) -> Window:
    """reference: slaf/ml/aggregators.py:267-277."""
    if isinstance(window_type, str):
        window_type = WindowType(window_type.lower())
    if window_type is WindowType.GENEFORMER:
        return GeneformerWindow()
    if window_type is WindowType.SCPGPT:
        return ScGPTWindow()
    if window_type is WindowType.SIMPLE:
        return SimpleWindow()
    raise ValueError(f"Unknown window type: {window_type!r}")
