"""The second caller of the hot path: mirrors of the reference's *generic* batch pipeline
(slaf/distributed/processor.py:105-301 ``BatchProcessor``, slaf/distributed/boundary.py:24-357
``GroupBoundaryHandler`` -- local part only, slaf/distributed/worker.py:120-168 ``tokenize_grouped``,
slaf/distributed/dataloader.py:357-475 ``_format_batch``).

    Boundary handling -> Shuffle -> Window -> Tokenize -> per-group samples -> collated batch

The Modal transport, queues and cross-worker KV store of that subsystem are control plane and out of
scope; what is mirrored is the compute those workers run on their CPUs, here one fused device pass
(same kernels as ``PrefetchBatchProcessor``) whenever the strategy objects are this package's, and
the three separate calls on host frames otherwise.  Samples are views of the batch tensors.
"""

from __future__ import annotations

from typing import Any, Iterator

import numpy as np
import torch

from .. import _native as N
from .. import runtime
from ..engine import regroup_contiguous
from . import frames
from .aggregators import GenericWindow, GeneformerWindow, ScGPTWindow, SimpleWindow
from .samplers import GenericShuffle, RandomShuffle
from .tokenizers import GeneformerTokenizer, ScGPTTokenizer


class GroupBoundaryHandler:
    """Local boundary tracking (slaf/distributed/boundary.py:24-153, 285-357 without the KV store):
    the last group of a batch may continue in the next batch of the same partition."""

    def __init__(self, schema: Any, continuity_check: str = "sequential", partial_groups_kv: Any = None):
        if partial_groups_kv is not None:
            raise NotImplementedError("cross-worker partial-group sharing is the Modal control plane (out of scope)")
        self.schema = schema
        self.continuity_check = continuity_check
        self.last_seen_groups: dict[int, Any] = {}

    def check_continuity(self, last_group_id: Any, first_group_id: Any, partition_id: int | None = None) -> bool:
        if last_group_id == first_group_id:
            return True
        if self.continuity_check == "sequential":
            try:
                return int(first_group_id) == int(last_group_id) + 1
            except (ValueError, TypeError):
                return False
        if self.continuity_check == "ordered":
            try:
                return first_group_id > last_group_id
            except TypeError:
                return False
        return False

    def merge_partial_data(self, partial_data: dict, new_batch: dict, partition_id: int | None = None,
                           is_partition_exhausted: bool = False) -> tuple[dict, dict]:
        """(complete rows, remaining partial groups); frames are dict-of-columns."""
        gk = self.schema.group_key
        if len(new_batch[gk]) == 0:
            return new_batch, partial_data
        first = new_batch[gk][0].item()
        merged, remaining = new_batch, dict(partial_data)
        last = self.last_seen_groups.get(partition_id) if partition_id is not None else None
        if last is not None and self.check_continuity(last, first, partition_id) and last in partial_data:
            part = remaining.pop(last)
            merged = {k: np.concatenate([np.asarray(part[k]), np.asarray(new_batch[k])]) for k in new_batch}
        if is_partition_exhausted:
            return merged, remaining
        ids = np.asarray(merged[gk])
        last_id = ids.max()                                  # unique().sort()[-1], boundary.py:139
        is_last = ids == last_id
        complete = {k: np.asarray(v)[~is_last] for k, v in merged.items()}
        remaining[last_id.item()] = {k: np.asarray(v)[is_last] for k, v in merged.items()}
        if partition_id is not None:
            self.last_seen_groups[partition_id] = last_id.item()
        return complete, remaining


def tokenize_grouped(tokenizer: Any):
    """``TokenizerProtocol`` callable over a grouped frame (slaf/distributed/worker.py:120-168)."""

    def call(grouped: Any, schema: Any) -> dict[str, Any]:
        genes = frames.list_column(grouped, schema.item_list_key)
        kind = frames.frame_kind(grouped)
        names = grouped.column_names if kind == "arrow" else list(grouped.columns)
        if schema.value_list_key and schema.value_list_key in names:
            ids, mask, vals = tokenizer.tokenize(genes, frames.list_column(grouped, schema.value_list_key))
        else:
            ids, mask, vals = tokenizer.tokenize(genes)
        if hasattr(tokenizer, "n_expression_bins") and vals is None:
            raise ValueError("scGPT tokenizer must return an aligned values stream")
        out = {"input_ids": ids, "attention_mask": mask}
        if vals is not None:
            out["values"] = vals
        return out

    call.tokenizer = tokenizer
    return call


class BatchProcessor:
    """reference: slaf/distributed/processor.py:105-301."""

    def __init__(self, schema: Any, window: Any = None, shuffle: Any = None, tokenizer: Any = None, boundary_handler: Any = None,
                 max_items: int = 1024, seed: int = 42, continuity_check: str = "sequential", **window_kwargs: Any):
        self.schema = schema
        self.window = window
        self.shuffle = shuffle
        self.tokenizer = tokenizer
        self.max_items = max_items
        self.seed = seed
        self.window_kwargs = window_kwargs
        self.batch_id = 0
        self.current_epoch = 0
        self.boundary_handler = boundary_handler or GroupBoundaryHandler(schema=schema, continuity_check=continuity_check)
        self.partial_groups: dict[Any, Any] = {}

    def _fused_params(self) -> dict | None:
        tok = getattr(self.tokenizer, "tokenizer", None)
        if type(tok) not in (GeneformerTokenizer, ScGPTTokenizer):
            return None
        if self.shuffle is not None and type(self.shuffle) not in (GenericShuffle, RandomShuffle):
            return None
        if type(self.window) not in (GenericWindow, GeneformerWindow, ScGPTWindow, SimpleWindow):
            return None
        s = self.schema
        if (s.group_key, s.item_key, s.value_key) != ("cell_integer_id", "gene_integer_id", "value"):
            return None
        kw = dict(self.window_kwargs)
        if type(tok) is ScGPTTokenizer:
            if not s.value_list_key:
                return None
            if type(self.window) in (GenericWindow, SimpleWindow):
                kw["use_binned_expressions"] = False
            elif type(self.window) is not ScGPTWindow:
                return None
            p = tok.hot_path_params(use_binned_expressions=kw.get("use_binned_expressions", True),
                                    n_expression_bins=kw.get("n_expression_bins", 10), window_kwargs=kw)
        else:
            # every window applies the same dense-rank filter; the Geneformer tokenizer ignores value lists
            p = tok.hot_path_params(window_kwargs=kw if type(self.window) is GeneformerWindow else {})
        p["max_items"] = int(self.max_items)
        return p

    def process_batch(self, raw_batches: list[Any], epoch: int = 0, partition_id: int | None = None,
                      is_partition_exhausted: bool = False) -> list[dict[str, Any]]:
        s = self.schema
        names = (s.group_key, s.item_key, s.value_key)
        combined = {k: np.concatenate([frames.column(b, k) for b in raw_batches]) for k in names}
        complete, self.partial_groups = self.boundary_handler.merge_partial_data(
            self.partial_groups, combined, partition_id=partition_id, is_partition_exhausted=is_partition_exhausted)
        if len(complete[s.group_key]) == 0:
            return []
        seed = self.seed + self.batch_id + epoch * 10000
        fused = self._fused_params() if self.tokenizer is not None else None
        if fused is not None:
            cols = frames.extract(complete, s)
            cell, gene, value = cols.cell, cols.gene, cols.value
            kw = dict(shuffle_seed=seed if self.shuffle is not None else None, check_contiguous=True, **fused)
            with runtime.session(len(cell), device=self.tokenizer.tokenizer.device) as eng:
                try:
                    out = eng.tokenize(cell, gene, value, **kw)
                except N.SlafB200Error as err:
                    if err.code != N.ERR_NONCONTIGUOUS:
                        raise
                    out = eng.tokenize(*regroup_contiguous(cell, gene, value), **kw)      # partition_by semantics for split groups
            keys = out.cell_ids.cpu().tolist()
            samples = []
            for i, key in enumerate(keys):
                smp = {"input_ids": out.input_ids[i], "attention_mask": out.attention_mask[i], "group_key": key}
                if out.values is not None:
                    smp["values"] = out.values[i]
                samples.append(smp)
            self.batch_id += 1
            return samples
        # generic route: the reference's three calls on host frames
        shuffled = complete
        if self.shuffle is not None:
            shuffled = self.shuffle.apply(complete, schema=s, seed=seed)
            if isinstance(shuffled, list):
                shuffled = {k: np.concatenate([np.asarray(c[k]) for c in shuffled]) for k in names}
        window = self.window if self.window is not None else GenericWindow()
        max_items = self.max_items if self.window is not None else max(1, len(shuffled[s.group_key]))
        grouped = window.apply(shuffled, schema=s, max_items=max_items, **self.window_kwargs)
        key_col = s.group_key_out or s.group_key
        if key_col not in (grouped.column_names if frames.frame_kind(grouped) == "arrow" else list(grouped.columns)):
            key_col = s.group_key
        keys = [k.item() if hasattr(k, "item") else k for k in frames.column(grouped, key_col)]
        if self.tokenizer is not None:
            tok = self.tokenizer(grouped, s)
            ids, mask = tok["input_ids"], tok["attention_mask"]
            others = {k: v for k, v in tok.items() if k not in ("input_ids", "attention_mask")}
            samples = [{"input_ids": ids[i], "attention_mask": mask[i], "group_key": key,
                        **{k: (v[i] if hasattr(v, "__getitem__") and len(v) > i else v) for k, v in others.items()}}
                       for i, key in enumerate(keys)]
        else:
            items = frames.list_column(grouped, s.item_list_key)
            cols_out = grouped.column_names if frames.frame_kind(grouped) == "arrow" else list(grouped.columns)
            vals = frames.list_column(grouped, s.value_list_key) if (s.value_list_key and s.value_list_key in cols_out) else None
            samples = []
            for i, key in enumerate(keys):
                g = {key_col: [key], s.item_list_key: [items[i]]}
                if vals is not None:
                    g[s.value_list_key] = [vals[i]]
                samples.append({"grouped": g, "group_key": key})
        self.batch_id += 1
        return samples


def format_batch(samples: list[dict[str, Any]], return_tensors: bool = True) -> Iterator[dict[str, Any]]:
    """reference: DistributedDataLoader._format_batch (slaf/distributed/dataloader.py:357-475)."""
    if not samples:
        return
    first = samples[0]
    if "input_ids" in first and "attention_mask" in first:
        if not return_tensors:
            yield {"input_ids": [x["input_ids"].tolist() for x in samples], "attention_mask": [x["attention_mask"].tolist() for x in samples],
                   "cell_ids": [x.get("group_key", 0) or 0 for x in samples]}
            return
        input_ids = torch.stack([x["input_ids"] for x in samples])
        attention_mask = torch.stack([x["attention_mask"] for x in samples])
        cell_ids = torch.tensor([x.get("group_key", 0) for x in samples], dtype=torch.long, device=input_ids.device)
        batch = {"input_ids": input_ids, "attention_mask": attention_mask, "cell_ids": cell_ids}
        for key in first:
            if key not in ("input_ids", "attention_mask", "group_key", "values"):
                batch[key] = [x.get(key) for x in samples]
        if "values" in first:
            values = torch.stack([x["values"] for x in samples])
            if values.shape != input_ids.shape:
                raise ValueError("Distributed scGPT dual-stream contract violated: values and "
                                 f"input_ids shapes differ ({tuple(values.shape)} vs {tuple(input_ids.shape)})")
            if values.shape != attention_mask.shape:
                raise ValueError("Distributed scGPT dual-stream contract violated: values and "
                                 f"attention_mask shapes differ ({tuple(values.shape)} vs {tuple(attention_mask.shape)})")
            batch["values"] = values
        yield batch
    elif "grouped" in first:
        yield {"grouped": [x["grouped"] for x in samples], "group_keys": [x.get("group_key", 0) for x in samples]}
    else:
        yield {"samples": samples}
