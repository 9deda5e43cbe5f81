"""Frame adapters of the drop-in facades.

The reference's Window / Shuffle interfaces take and return polars DataFrames
(slaf/ml/aggregators.py:26-50).  polars is optional here: the facades accept a polars DataFrame
(when polars is importable), a pyarrow Table / RecordBatch, a pandas DataFrame or a plain dict of
column arrays, and answer in the same family.  Only column extraction and the packaging of the
kernels' flat list output into list columns happens here -- no hot-path arithmetic.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np

try:  # optional
    import polars as _pl
except Exception:  # pragma: no cover - polars is not installed in the build image
    _pl = None
try:
    import pyarrow as _pa
except Exception:  # pragma: no cover
    _pa = None
try:
    import pandas as _pd
except Exception:  # pragma: no cover
    _pd = None


@dataclass
class CooColumns:
    """COO columns in kernel dtypes plus what is needed to answer in the caller's dtypes."""

    cell: np.ndarray          # uint32 | int32
    gene: np.ndarray          # uint16 | int32 | uint32
    value: np.ndarray         # uint16 | float32
    kind: str                 # "polars" | "arrow" | "pandas" | "dict"
    cell_dtype: np.dtype      # caller's dtypes (for the answer)
    gene_dtype: np.dtype
    value_dtype: np.dtype
    value_is_float: bool      # caller's value column was floating point (bins are floats of that width)

    @property
    def n_rows(self) -> int:
        return int(self.cell.shape[0])


def frame_kind(frame: Any) -> str:
    if _pl is not None and isinstance(frame, _pl.DataFrame):
        return "polars"
    if _pa is not None and isinstance(frame, (_pa.Table, _pa.RecordBatch)):
        return "arrow"
    if _pd is not None and isinstance(frame, _pd.DataFrame):
        return "pandas"
    if isinstance(frame, dict):
        return "dict"
    raise TypeError(f"unsupported frame type {type(frame)!r}: expected polars/pyarrow/pandas frame or dict of columns")


def frame_len(frame: Any) -> int:
    kind = frame_kind(frame)
    if kind == "arrow":
        return frame.num_rows
    if kind == "dict":
        return len(next(iter(frame.values()))) if frame else 0
    return len(frame)


def column(frame: Any, name: str) -> np.ndarray:
    kind = frame_kind(frame)
    if kind == "polars":
        return frame[name].to_numpy()
    if kind == "arrow":
        col = frame.column(name) if isinstance(frame, _pa.Table) else frame.column(frame.schema.get_field_index(name))
        if isinstance(col, _pa.ChunkedArray):
            col = col.combine_chunks()
        if _pa.types.is_null(col.type):
            return np.zeros(len(col), np.int64)
        return col.to_numpy(zero_copy_only=False)
    if kind == "pandas":
        return frame[name].to_numpy()
    return np.asarray(frame[name])


def _as_cell(a: np.ndarray) -> np.ndarray:
    if a.dtype in (np.uint32, np.int32):
        return np.ascontiguousarray(a)
    if a.size == 0:
        return a.astype(np.int32)
    if not np.issubdtype(a.dtype, np.integer):
        raise TypeError(f"group key column must be integer, got {a.dtype}")
    lo, hi = int(a.min()), int(a.max())
    if lo >= 0 and hi <= 0xFFFFFFFF:
        return a.astype(np.uint32)
    if lo >= -(2**31) and hi < 2**31:
        return a.astype(np.int32)
    raise ValueError("group key values do not fit 32 bits (SLAF cell_integer_id is uint32)")


def _as_gene(a: np.ndarray) -> np.ndarray:
    if a.dtype in (np.uint16, np.int32, np.uint32):
        return np.ascontiguousarray(a)
    if a.size == 0:
        return a.astype(np.int32)
    if not np.issubdtype(a.dtype, np.integer):
        raise TypeError(f"item key column must be integer, got {a.dtype}")
    lo, hi = int(a.min()), int(a.max())
    if lo < 0 or hi >= 2**31:
        raise ValueError("item key values must be in [0, 2^31)")
    return a.astype(np.int32)


def _as_value(a: np.ndarray) -> tuple[np.ndarray, bool]:
    """uint16 counts or float32.  Wider integer / float64 columns are narrowed only when that is
    exact; integral float64 values in uint16 range take the uint16 path, whose log1p binning is
    computed in float64 exactly like polars does for a Float64 column."""
    if a.dtype == np.uint16:
        return np.ascontiguousarray(a), False
    if a.dtype == np.float32:
        return np.ascontiguousarray(a), True
    if a.size == 0:
        return a.astype(np.float32 if np.issubdtype(a.dtype, np.floating) else np.uint16), bool(np.issubdtype(a.dtype, np.floating))
    if np.issubdtype(a.dtype, np.integer) or a.dtype == np.bool_:
        if int(a.min()) < 0 or int(a.max()) > 65535:
            raise ValueError("integer values outside the uint16 count range are not supported (SLAF stores uint16 counts or float32)")
        return a.astype(np.uint16), False
    if np.issubdtype(a.dtype, np.floating):
        if np.isnan(a).any():
            raise ValueError("NaN values are outside the SLAF value contract")
        if np.all(a == np.floor(a)) and a.min() >= 0 and a.max() <= 65535:
            return a.astype(np.uint16), True
        f = a.astype(np.float32)
        if not np.array_equal(f.astype(a.dtype), a):
            raise TypeError("float64 values that are not exactly representable in float32 are not supported "
                            "(SLAF stores float32); cast the column to float32 first")
        return f, True
    raise TypeError(f"unsupported value dtype {a.dtype}")


def extract(frame: Any, schema: Any) -> CooColumns:
    kind = frame_kind(frame)
    cell = column(frame, schema.group_key)
    gene = column(frame, schema.item_key)
    value = column(frame, schema.value_key)
    v, is_float = _as_value(value)
    return CooColumns(_as_cell(cell), _as_gene(gene), v, kind, cell.dtype, gene.dtype, value.dtype, is_float)


def take_rows(frame: Any, order: np.ndarray) -> Any:
    """frame[order] in the frame's own family (used by the shuffle facade)."""
    kind = frame_kind(frame)
    if kind == "polars":
        return frame[order]
    if kind == "arrow":
        return frame.take(_pa.array(order))
    if kind == "pandas":
        return frame.iloc[order].reset_index(drop=True)
    return {k: np.asarray(v)[order] for k, v in frame.items()}


def slice_rows(frame: Any, lo: int, hi: int) -> Any:
    kind = frame_kind(frame)
    if kind in ("polars", "arrow"):
        return frame.slice(lo, hi - lo)
    if kind == "pandas":
        return frame.iloc[lo:hi].reset_index(drop=True)
    return {k: np.asarray(v)[lo:hi] for k, v in frame.items()}


class GroupedFrame(dict):
    """dict-of-columns answer for dict input: list columns are Python lists of numpy arrays.
    ``len(frame)`` is the number of groups, ``frame.columns`` the column names."""

    @property
    def columns(self) -> list[str]:
        return list(self.keys())

    def __len__(self) -> int:  # number of groups, like a DataFrame
        for v in self.values():
            return len(v)
        return 0


def build_grouped(cols: CooColumns, schema: Any, cell_ids: np.ndarray, offsets: np.ndarray, genes: np.ndarray,
                  exprs: np.ndarray | None, expr_is_bins: bool) -> Any:
    """Package flat list output as the grouped frame Window.apply returns."""
    gk = schema.group_key
    item_out = schema.item_list_key
    value_out = (schema.value_list_key or "expr_sequence") if exprs is not None else None
    cell_np = cell_ids.astype(cols.cell_dtype, copy=False)
    gene_np = genes.astype(cols.gene_dtype, copy=False)
    expr_np = None
    if exprs is not None:
        if expr_is_bins:
            # polars: integer column -> Float64 bins; Float32 stays Float32; Float64 stays Float64
            bins_dtype = cols.value_dtype if cols.value_is_float else np.dtype(np.float64)
            expr_np = exprs.astype(bins_dtype, copy=False)
        else:
            expr_np = exprs.astype(cols.value_dtype, copy=False)
    if cols.kind in ("polars", "arrow", "pandas") and _pa is not None:
        offs = _pa.array(offsets.astype(np.int64))
        arrays = [_pa.array(cell_np), _pa.LargeListArray.from_arrays(offs, _pa.array(gene_np))]
        names = [gk, item_out]
        if expr_np is not None:
            arrays.append(_pa.LargeListArray.from_arrays(offs, _pa.array(expr_np)))
            names.append(value_out)
        table = _pa.table(arrays, names=names)
        if cols.kind == "polars":
            return _pl.from_arrow(table)
        if cols.kind == "pandas":
            return table.to_pandas()
        return table
    out = GroupedFrame()
    out[gk] = cell_np
    out[item_out] = [gene_np[offsets[i]:offsets[i + 1]] for i in range(len(cell_np))]
    if expr_np is not None:
        out[value_out] = [expr_np[offsets[i]:offsets[i + 1]] for i in range(len(cell_np))]
    return out


def list_column(frame: Any, name: str) -> list[list]:
    """grouped[name].to_list() for any supported frame family."""
    kind = frame_kind(frame)
    if kind == "polars":
        return frame[name].to_list()
    if kind == "arrow":
        return frame.column(name).to_pylist()
    if kind == "pandas":
        return [list(x) for x in frame[name]]
    return [np.asarray(x).tolist() for x in frame[name]]
