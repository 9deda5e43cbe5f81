"""Tokenizers on the B200: same classes, constructor arguments, attributes and ``tokenize``
signature as the reference's ``slaf.ml.tokenizers`` (slaf/ml/tokenizers.py:18-774).

    tokenizer.tokenize(gene_sequences, expr_sequences=None) -> (input_ids, attention_mask, values)

    Geneformer : [CLS] gene+4 ... [SEP] truncated / right-padded to max_genes        (:635-722)
    scGPT      : ids    [CLS] gene+4 ... [SEP] PAD...     width max_genes + 2         (:375-479)
                 values [PAD] tok(expr) ... [PAD] PAD...   aligned position by position

The token arithmetic runs in the CUDA kernels (``slafb_tokenize_lists`` for pre-windowed lists,
``slafb_tokenize`` for the fused COO path); this module flattens the Python lists the reference
interface hands over, owns the vocabulary bookkeeping (host metadata, never on the per-batch
path) and re-raises the reference's exceptions.

Output placement: tensors live on the tokenizer's CUDA device by default (``device=None`` = the
current device).  ``output_device="cpu"`` adds an explicit D2H copy of the finished tensors for
code that asserts the reference's CPU placement -- the computation is on the GPU either way.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from typing import Any

import numpy as np
import torch

from .. import runtime
from .aggregators import GeneformerWindow, ScGPTWindow, Window

SPECIAL_TOKENS = {"PAD": 0, "CLS": 1, "SEP": 2, "MASK": 3}
GENE_TOKEN_OFFSET = 4          # gene_integer_id -> token id (tokenizers.py:439,686)


class TokenizerType(str, Enum):
    GENEFORMER = "geneformer"
    SCGPT = "scgpt"


def _gene_table(slaf_array: Any):
    """(gene_id list, gene_integer_id list) from ``slaf_array.var`` / ``slaf_array.genes`` when they
    look like real tables; None for mocks and anything else (reference falls back the same way,
    tokenizers.py:78-186)."""
    from . import frames

    candidates = []
    try:
        ready = slaf_array.is_metadata_ready()
        if ready is True:
            var = slaf_array.var
            if hasattr(var, "reset_index"):
                var = var.reset_index()
            candidates.append(var)
    except Exception:
        pass
    try:
        genes = slaf_array.genes
        table = genes.to_table() if hasattr(genes, "to_table") else genes
        candidates.append(table)
    except Exception:
        pass
    for cand in candidates:
        try:
            kind = frames.frame_kind(cand)
            names = cand.column_names if kind == "arrow" else (list(cand.keys()) if kind == "dict" else list(cand.columns))
            if "gene_integer_id" in names and "gene_id" in names:
                ids = frames.column(cand, "gene_id").tolist()
                ints = [int(x) for x in frames.column(cand, "gene_integer_id").tolist()]
                return ids, ints
        except Exception:
            continue
    return None


def _flatten(seqs, dtype) -> tuple[np.ndarray, np.ndarray]:
    """list of sequences -> (int64 offsets [B+1], flat array).  Control-plane packing only."""
    n = len(seqs)
    lens = np.fromiter((len(s) if s is not None else 0 for s in seqs), dtype=np.int64, count=n)
    offs = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    total = int(offs[-1])
    flat = np.empty(total, dtype)
    for i, s in enumerate(seqs):
        if lens[i]:
            flat[offs[i]:offs[i + 1]] = s
    return offs, flat


class SLAFTokenizer(ABC):
    """reference: slaf/ml/tokenizers.py:25-294."""

    def __init__(self, slaf_array: Any, vocab_size: int = 50000, max_genes: int = 2048, device=None,
                 output_device: str | torch.device | None = None):
        self.slaf_array = slaf_array
        self.vocab_size = vocab_size
        if max_genes < 1:
            raise ValueError(f"max_genes must be >= 1, got {max_genes}")
        self.max_genes = max_genes
        self.device = device
        self.output_device = torch.device(output_device) if output_device is not None else None
        self.window = self.create_window()
        self.window.device = device
        self._build_gene_vocabulary()
        self._setup_special_tokens()

    # ------------------------------------------------------------------ vocabulary (host metadata)
    def _build_gene_vocabulary(self) -> None:
        table = _gene_table(self.slaf_array)
        if table is None:
            self.gene_vocab = {f"gene_{i}": i for i in range(1000)}
            self.token_to_gene = {i + GENE_TOKEN_OFFSET: g for g, i in self.gene_vocab.items()}
            return
        ids, ints = table
        self.gene_vocab = {g: i for g, i in zip(ids, ints) if i < self.vocab_size}
        self.token_to_gene = {i + GENE_TOKEN_OFFSET: g for g, i in self.gene_vocab.items()}
        numeric = []
        for g in self.gene_vocab:
            try:
                numeric.append(int(g))
            except (ValueError, TypeError):
                pass
        size = (max(numeric) if numeric else 0) + 1
        self.vocab_mapping = np.full(size, -1, dtype=int)
        for g, tok in self.gene_vocab.items():
            try:
                self.vocab_mapping[int(g)] = tok
            except (ValueError, TypeError):
                continue

    def _setup_special_tokens(self) -> None:
        self.special_tokens = dict(SPECIAL_TOKENS)

    def _map_gene_ids_to_tokens_vectorized(self, gene_ids) -> np.ndarray:
        """reference: tokenizers.py:198-232 (not used by ``tokenize``)."""
        if gene_ids is None or len(gene_ids) == 0:
            return np.array([], dtype=int)
        arr = np.array(gene_ids, dtype=int)
        mapping = getattr(self, "vocab_mapping", None)
        if mapping is None:
            return arr + GENE_TOKEN_OFFSET
        tokens = mapping[arr]
        return tokens[tokens != -1]

    @abstractmethod
    def create_window(self) -> Window:
        raise NotImplementedError

    def get_vocab_info(self) -> dict[str, Any]:
        return {"vocab_size": self.vocab_size, "special_tokens": self.special_tokens,
                "gene_vocab_size": len(self.gene_vocab), "max_genes": self.max_genes}

    @abstractmethod
    def tokenize(self, gene_sequences, expr_sequences=None):
        raise NotImplementedError

    @abstractmethod
    def decode_tokens(self, tokens: list[int]) -> dict[str, Any]:
        raise NotImplementedError

    # ------------------------------------------------------------------ shared plumbing
    @property
    def sequence_length(self) -> int:
        return self.max_genes

    def _place(self, *tensors):
        if self.output_device is None:
            return tensors
        return tuple(t.to(self.output_device) if t is not None else None for t in tensors)

    def _decode_common(self, tokens, with_expressions: bool) -> dict[str, Any]:
        out = {"genes": [], "expressions": [], "special_tokens": []}
        names = {v: k for k, v in self.special_tokens.items()}
        for tok in tokens or []:
            if tok in names:
                out["special_tokens"].append(names[tok])
            elif with_expressions and tok >= self.expr_bin_start:
                out["expressions"].append((tok - self.expr_bin_start) * self.expr_bin_size)
            else:
                out["genes"].append(self.token_to_gene.get(tok, f"unknown_gene_{tok}"))
        return out

    # fused path used by the prefetch batch processor ------------------------------------------
    def hot_path_params(self, use_binned_expressions: bool = True, n_expression_bins: int | None = None,
                        window_kwargs: dict | None = None) -> dict:
        """Keyword arguments of ``HotPathEngine.tokenize`` equivalent to
        ``self.window.apply(...)`` followed by ``self.tokenize(...)``."""
        raise NotImplementedError


class ScGPTTokenizer(SLAFTokenizer):
    """reference: slaf/ml/tokenizers.py:297-600."""

    def __init__(self, slaf_array: Any, vocab_size: int = 50000, n_expression_bins: int = 10, max_genes: int = 1024,
                 device=None, output_device=None):
        self.n_expression_bins = n_expression_bins
        super().__init__(slaf_array=slaf_array, vocab_size=vocab_size, max_genes=max_genes, device=device,
                         output_device=output_device)

    def create_window(self) -> Window:
        return ScGPTWindow()

    def _setup_special_tokens(self) -> None:
        super()._setup_special_tokens()
        self.expr_bin_start = self.vocab_size
        self.expr_bin_size = 1.0 / self.n_expression_bins

    @property
    def sequence_length(self) -> int:
        return self.max_genes + 2

    def get_vocab_info(self) -> dict[str, Any]:
        return {**super().get_vocab_info(), "n_expression_bins": self.n_expression_bins}

    def _expression_to_bin(self, expression_value: float) -> int:
        if expression_value <= 0:
            return self.special_tokens["PAD"]
        return self.expr_bin_start + min(int(expression_value / self.expr_bin_size), self.n_expression_bins - 1)

    def _expression_to_bin_vectorized(self, expression_values) -> np.ndarray:
        """reference: tokenizers.py:521-543, computed by the list-tokenizer kernel (float32 divide by
        float32(1/n_bins), truncate, clip, ``> 0`` select)."""
        x = np.asarray(expression_values)
        if x.size == 0:
            return np.array([], dtype=int)
        with runtime.session(x.size, 1, device=self.device) as eng:
            dev = eng.device
            offs = torch.tensor([0, x.size], dtype=torch.int64, device=dev)
            genes = torch.zeros(x.size, dtype=torch.int64, device=dev)
            exprs = torch.from_numpy(x.astype(np.float32).astype(np.float64)).to(dev)
            _, _, vals = eng.tokenize_lists(offs, genes, exprs, scgpt=True, exprs_are_int=False, max_genes=int(x.size),
                                            n_bins=self.n_expression_bins, vocab_size=self.expr_bin_start)
            return vals[0, 1:1 + x.size].cpu().numpy().astype(int)

    def tokenize(self, gene_sequences, expr_sequences=None):
        if gene_sequences is None or len(gene_sequences) == 0:
            raise ValueError("Gene sequences cannot be empty")
        B = len(gene_sequences)
        if expr_sequences is None:
            expr_sequences = [[] for _ in range(B)]
        # pairs = min(len(genes), len(exprs), max_genes) per cell (tokenizers.py:436): trim to the common length
        lens = [min(len(g) if g is not None else 0, len(e) if e is not None else 0) for g, e in zip(gene_sequences, expr_sequences)]
        genes_t = [g[:n] if n else [] for g, n in zip(gene_sequences, lens)]
        exprs_t = [e[:n] if n else [] for e, n in zip(expr_sequences, lens)]
        # the branch is chosen per cell from the type of its first expression (tokenizers.py:441)
        is_int = np.fromiter((bool(n) and isinstance(e[0], (int, np.integer)) and not isinstance(e[0], bool) or
                              (bool(n) and isinstance(e[0], (bool, np.bool_)))
                              for e, n in zip(exprs_t, lens)), dtype=bool, count=B)
        nonempty = np.asarray(lens) > 0
        offs, gflat = _flatten(genes_t, np.int64)
        _, xflat = _flatten(exprs_t, np.float64)
        if not is_int[nonempty].any():
            xflat = xflat.astype(np.float32).astype(np.float64)   # np.array(exprs, dtype=np.float32), tokenizers.py:446
        kw = dict(scgpt=True, max_genes=self.max_genes, n_bins=self.n_expression_bins, vocab_size=self.expr_bin_start)
        all_int = bool(is_int[nonempty].all()) if nonempty.any() else False
        all_float = not is_int[nonempty].any()
        with runtime.session(max(int(offs[-1]), 1), B, device=self.device) as eng:
            dev = eng.device
            d_offs, d_genes, d_exprs = (torch.from_numpy(a).to(dev) for a in (offs, gflat, xflat))
            if all_int or all_float:
                ids, mask, vals = eng.tokenize_lists(d_offs, d_genes, d_exprs, exprs_are_int=all_int, **kw)
            else:
                # mixed Python int / float cells in one call: run both branches, select per cell
                x32 = torch.from_numpy(xflat.astype(np.float32).astype(np.float64)).to(dev)
                ids, mask, vals_i = eng.tokenize_lists(d_offs, d_genes, d_exprs, exprs_are_int=True, **kw)
                _, _, vals_f = eng.tokenize_lists(d_offs, d_genes, x32, exprs_are_int=False, **kw)
                sel = torch.from_numpy(is_int).to(dev)[:, None]
                vals = torch.where(sel, vals_i, vals_f)
        return self._place(ids, mask, vals)

    def decode_tokens(self, tokens: list[int]) -> dict[str, Any]:
        return self._decode_common(tokens, with_expressions=True)

    def hot_path_params(self, use_binned_expressions: bool = True, n_expression_bins: int | None = None,
                        window_kwargs: dict | None = None) -> dict:
        kw = dict(window_kwargs or {})
        binned = kw.get("use_binned_expressions", use_binned_expressions)
        # the WINDOW bins with the processor's n_expression_bins, the TOKENIZER with its own
        # (datasets.py:1025-1046 vs tokenizers.py:507-508); the kernels take both
        n_window = kw.get("n_expression_bins", n_expression_bins if n_expression_bins is not None else self.n_expression_bins)
        return dict(mode="scgpt_binned" if binned else "scgpt_raw", max_genes=self.max_genes, max_items=self.max_genes,
                    n_bins=self.n_expression_bins, window_n_bins=int(n_window), vocab_size=self.expr_bin_start)


class GeneformerTokenizer(SLAFTokenizer):
    """reference: slaf/ml/tokenizers.py:603-774."""

    def __init__(self, slaf_array: Any, vocab_size: int = 50000, max_genes: int = 2048, device=None, output_device=None):
        super().__init__(slaf_array=slaf_array, vocab_size=vocab_size, max_genes=max_genes, device=device,
                         output_device=output_device)

    def create_window(self) -> Window:
        return GeneformerWindow()

    def tokenize(self, gene_sequences, expr_sequences=None):
        if gene_sequences is None or len(gene_sequences) == 0:
            raise ValueError("Gene sequences cannot be empty")
        offs, gflat = _flatten(gene_sequences, np.int64)
        with runtime.session(max(int(offs[-1]), 1), len(gene_sequences), device=self.device) as eng:
            dev = eng.device
            ids, mask, _ = eng.tokenize_lists(torch.from_numpy(offs).to(dev), torch.from_numpy(gflat).to(dev), None,
                                              scgpt=False, exprs_are_int=True, max_genes=self.max_genes)
        ids, mask = self._place(ids, mask)
        return ids, mask, None

    def decode_tokens(self, tokens: list[int]) -> dict[str, Any]:
        return self._decode_common(tokens, with_expressions=False)

    def hot_path_params(self, use_binned_expressions: bool = True, n_expression_bins: int | None = None,
                        window_kwargs: dict | None = None) -> dict:
        kw = dict(window_kwargs or {})
        pct = kw.get("min_percentile", None)
        out = dict(mode="geneformer_pct" if pct is not None else "geneformer", max_genes=self.max_genes, max_items=self.max_genes)
        if pct is not None:
            out["min_percentile"] = float(pct)
        return out
