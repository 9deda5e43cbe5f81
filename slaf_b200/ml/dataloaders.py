"""SLAFDataLoader on the B200: same constructor arguments, validation errors, attributes and batch
dictionaries as the reference's ``slaf.ml.dataloaders.SLAFDataLoader`` (slaf/ml/dataloaders.py:159-671).

    for batch in SLAFDataLoader(slaf_array, tokenizer_type="geneformer", batch_size=32, max_genes=2048):
        batch["input_ids"]       # int64 [<=batch_size, L]     on cuda:N (views of the prefetch batch)
        batch["attention_mask"]  # bool  [<=batch_size, L]
        batch["cell_ids"]        # int64 [<=batch_size]
        batch["values"]          # int64 [<=batch_size, L]     scGPT only
        batch["epoch"]           # only when n_epochs > 1

``device_raw=True`` (with ``raw_mode=True``) keeps raw mode on the GPU: ``batch["x"]`` is a dict of CSR device tensors
(``crow_indices``, ``col_indices``, ``values``, ``cell_ids``) over the batch's cells in shuffled order instead of a host frame.

``use_cell_start_index``: the cell boundaries of every chunk come from ``slaf_array._cell_start_index`` instead of the
``cell_integer_id`` column (which then neither crosses PCIe -- half of the bytes -- nor is scanned); the token tensors are identical.
Default ``None``: used whenever the array carries the index and it keeps validating against the rows (probed per load; the first
mismatch falls back to the cell column with a warning); ``True`` insists (a mismatch raises); ``False`` never uses it.

Two additions: ``device`` (CUDA device the path runs on; default current) and ``output_device``
("cpu" adds an explicit D2H copy of the finished batch for code that asserts the reference's CPU
placement, tests/test_dataloaders.py:186-196).  ``max_queue_size`` counts *prefetch batches*
resident in HBM, so its default is small (8) where the reference's host queue default is 5000.
"""

from __future__ import annotations

from typing import Any

import torch

from .datasets import SLAFIterableDataset
from .tokenizers import GeneformerTokenizer, ScGPTTokenizer


def get_optimal_device():
    """reference: slaf/ml/dataloaders.py:31-78.  Here the answer can only be a CUDA device: the path has no CPU form."""
    return torch.device("cuda") if torch.cuda.is_available() else None


def get_device_info() -> dict:
    """reference: slaf/ml/dataloaders.py:81-154 (same keys)."""
    info = {"torch_available": True, "cuda_available": torch.cuda.is_available(), "mps_available": False,
            "optimal_device": str(get_optimal_device())}
    if torch.cuda.is_available():
        info["cuda_device_count"] = torch.cuda.device_count()
        info["cuda_device_name"] = torch.cuda.get_device_name(0)
        info["cuda_device_capability"] = torch.cuda.get_device_capability(0)
    return info


class SLAFDataLoader:
    def __init__(self, slaf_array: Any, tokenizer_type: str = "geneformer", batch_size: int = 32, max_genes: int = 2048,
                 vocab_size: int = 50000, n_expression_bins: int = 10, n_epochs: int = 1, raw_mode: bool = False,
                 verbose: bool = True, batches_per_chunk: int = 1, by_fragment: bool = True,
                 use_mixture_of_scanners: bool = True, n_scanners: int = 16, prefetch_batch_size: int = 4194304,
                 max_queue_size: int = 8, parallelize_fragment_reads: bool = False, prefetcher_ready_timeout: float = 10.0,
                 device=None, output_device=None, device_raw: bool = False, use_cell_start_index: bool | None = None):
        self.slaf_array = slaf_array
        self.batch_size = batch_size
        self.max_genes = max_genes
        self.n_epochs = n_epochs
        self.raw_mode = raw_mode
        self.verbose = verbose
        self.batches_per_chunk = batches_per_chunk
        self.by_fragment = by_fragment
        self.use_mixture_of_scanners = use_mixture_of_scanners
        self.n_scanners = n_scanners
        self.prefetch_batch_size = prefetch_batch_size
        self.max_queue_size = max_queue_size
        self.parallelize_fragment_reads = parallelize_fragment_reads
        if self.use_mixture_of_scanners:                         # dataloaders.py:437-447
            if self.n_scanners < 1:
                raise ValueError("n_scanners must be at least 1")
            if self.n_scanners > 100:
                raise ValueError("n_scanners cannot exceed 100")
            if self.prefetch_batch_size < 1000:
                raise ValueError("prefetch_batch_size must be at least 1,000")
            if self.prefetch_batch_size > 10000000:
                raise ValueError("prefetch_batch_size cannot exceed 10,000,000")
        self.device = device
        if not self.raw_mode:
            if tokenizer_type == "geneformer":
                self.tokenizer = GeneformerTokenizer(slaf_array=slaf_array, vocab_size=vocab_size, max_genes=max_genes, device=device)
            elif tokenizer_type == "scgpt":
                self.tokenizer = ScGPTTokenizer(slaf_array=slaf_array, vocab_size=vocab_size, n_expression_bins=n_expression_bins,
                                                max_genes=max_genes, device=device)
            else:
                raise ValueError("tokenizer_type must be one of ['geneformer', 'scgpt']; "
                                 f"{tokenizer_type=} is not supported.")
            self.special_tokens = self.tokenizer.special_tokens
        else:
            self.tokenizer = None
            self.special_tokens = None
        self._dataset = SLAFIterableDataset(
            slaf_array=slaf_array, tokenizer=self.tokenizer, batch_size=batch_size, seed=42, max_queue_size=max_queue_size,
            n_epochs=n_epochs, raw_mode=raw_mode, verbose=verbose, batches_per_chunk=batches_per_chunk, by_fragment=by_fragment,
            use_mixture_of_scanners=use_mixture_of_scanners, n_scanners=n_scanners, prefetch_batch_size=prefetch_batch_size,
            parallelize_fragment_reads=parallelize_fragment_reads, prefetcher_ready_timeout=prefetcher_ready_timeout,
            device=device, output_device=output_device, device_raw=device_raw, use_cell_start_index=use_cell_start_index)

    def __iter__(self):
        yield from self._dataset

    def __len__(self) -> int:
        return 0   # streaming: unknown length, like the reference (dataloaders.py:628-651)

    def __del__(self):
        ds = getattr(self, "_dataset", None)
        if ds is not None:
            try:
                ds.prefetcher.stop()
            except Exception:
                pass
