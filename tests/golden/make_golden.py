"""Generates tests/golden/tokenizer_golden.npz by EXECUTING the reference's own
slaf/ml/tokenizers.py (unmodified, imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

polars and lance are not installed here, so two stub modules stand in for them; the
tokenizer code path exercised (tokenize / _expression_to_bin*) never touches either.
The slaf_array argument is a unittest.mock.Mock(spec=...) exactly like the reference's
own tests/test_tokenizers.py:23-26.
"""

import os
import sys
import types
from unittest.mock import Mock

import numpy as np

REF = os.environ.get("SLAF_REFERENCE", "/root/reference")


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None})


def import_reference_tokenizers():
    for name in ("polars", "lance", "loguru", "rich", "rich.console", "rich.panel", "psutil"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Anything(name)
    sys.path.insert(0, REF)
    import slaf.ml.tokenizers as tk  # noqa: E402

    return tk


def ragged(rng, n_cells, max_len, gene_hi):
    seqs = []
    for i in range(n_cells):
        n = int(rng.integers(0, max_len + 1))
        if i % 7 == 3:
            n = 0  # empty cell (tests/test_tokenizers.py:174-182)
        seqs.append(rng.integers(0, gene_hi, n).tolist())
    return seqs


def flatten(seqs, dtype):
    offs = np.zeros(len(seqs) + 1, np.int64)
    for i, s in enumerate(seqs):
        offs[i + 1] = offs[i] + len(s)
    flat = np.array([x for s in seqs for x in s], dtype)
    return offs, flat


def main():
    tk = import_reference_tokenizers()
    rng = np.random.default_rng(20261019)
    out = {}
    arr = Mock(spec=tk.SLAFArray) if isinstance(tk.SLAFArray, type) else Mock()

    # ---- Geneformer: several widths incl. sequences longer than max_genes (truncation, SEP dropped)
    for tag, max_genes, max_len in (("gf8", 8, 12), ("gf64", 64, 80), ("gf2048", 2048, 2300)):
        t = tk.GeneformerTokenizer(arr, vocab_size=50000, max_genes=max_genes)
        seqs = ragged(rng, 24, max_len, 60000)
        seqs[0] = list(range(max_genes - 2))   # SEP is the last token
        seqs[1] = list(range(max_genes - 1))   # SEP truncated
        seqs[2] = list(range(max_genes))
        ids, mask, vals = t.tokenize(seqs)
        assert vals is None
        offs, flat = flatten(seqs, np.int64)
        out[f"{tag}_offsets"], out[f"{tag}_genes"] = offs, flat
        out[f"{tag}_ids"], out[f"{tag}_mask"] = ids.numpy(), mask.numpy()
        out[f"{tag}_max_genes"] = np.int64(max_genes)

    # ---- scGPT, integer expression lists (raw uint16 counts from polars .to_list())
    for tag, max_genes, max_len, vocab, nb in (("sgi4", 4, 7, 1000, 51), ("sgi1200", 1200, 1500, 65000, 51)):
        t = tk.ScGPTTokenizer(arr, vocab_size=vocab, n_expression_bins=nb, max_genes=max_genes)
        seqs = ragged(rng, 20, max_len, 60000)
        exprs = [rng.integers(1, 4000, len(s)).tolist() for s in seqs]
        ids, mask, vals = t.tokenize(seqs, exprs)
        offs, flat = flatten(seqs, np.int64)
        _, xflat = flatten(exprs, np.float64)
        out[f"{tag}_offsets"], out[f"{tag}_genes"], out[f"{tag}_exprs"] = offs, flat, xflat
        out[f"{tag}_ids"], out[f"{tag}_mask"], out[f"{tag}_vals"] = ids.numpy(), mask.numpy(), vals.numpy()
        out[f"{tag}_cfg"] = np.array([max_genes, vocab, nb], np.int64)

    # ---- scGPT, float expression lists: (a) window bins 0..n-1, (b) raw float32 values
    for tag, max_genes, max_len, vocab, nb, kind in (
        ("sgf_bins", 16, 24, 1000, 51, "bins"),
        ("sgf_bins10", 16, 24, 50000, 10, "bins"),
        ("sgf_raw", 32, 40, 65000, 51, "raw"),
        ("sgf_raw5", 32, 40, 1000, 5, "raw"),
    ):
        t = tk.ScGPTTokenizer(arr, vocab_size=vocab, n_expression_bins=nb, max_genes=max_genes)
        seqs = ragged(rng, 20, max_len, 60000)
        if kind == "bins":
            exprs = [[float(v) for v in rng.integers(0, nb, len(s))] for s in seqs]
        else:
            exprs = [[float(np.float32(v)) for v in np.r_[rng.uniform(0, 1.2, len(s) // 2), rng.uniform(0, 9, len(s) - len(s) // 2)]] for s in seqs]
        ids, mask, vals = t.tokenize(seqs, exprs)
        offs, flat = flatten(seqs, np.int64)
        _, xflat = flatten(exprs, np.float64)
        out[f"{tag}_offsets"], out[f"{tag}_genes"], out[f"{tag}_exprs"] = offs, flat, xflat
        out[f"{tag}_ids"], out[f"{tag}_mask"], out[f"{tag}_vals"] = ids.numpy(), mask.numpy(), vals.numpy()
        out[f"{tag}_cfg"] = np.array([max_genes, vocab, nb], np.int64)

    # ---- binning KAT of tests/test_tokenizers.py:199-213 re-executed
    t = tk.ScGPTTokenizer(arr, vocab_size=1000, n_expression_bins=10, max_genes=8)
    xs = np.array([0.0, 0.1, 0.5, 0.9, -1.0, 1.0, 0.3, 0.7, 123.0, 1e-9], np.float32)
    out["binkat_x"] = xs
    out["binkat_tok"] = t._expression_to_bin_vectorized(xs).astype(np.int64)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tokenizer_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
