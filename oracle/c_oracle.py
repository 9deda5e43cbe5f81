"""ctypes loader for oracle/libslaf_oracle.so (TEST INFRASTRUCTURE ONLY -- see slaf_oracle.c).

Builds the library on demand with the system gcc.  Nothing under slaf_b200/ imports this.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_int32, c_int64, c_uint8, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libslaf_oracle.so")
_SRC = os.path.join(_HERE, "slaf_oracle.c")

DT = {np.dtype(np.uint16): 0, np.dtype(np.int32): 1, np.dtype(np.uint32): 2, np.dtype(np.float32): 3}
MODES = {"geneformer": 0, "geneformer_pct": 1, "scgpt_raw": 2, "scgpt_binned": 3}


class OracleParams(ctypes.Structure):
    _fields_ = [
        ("mode", c_int32),
        ("max_items", c_int32),
        ("max_genes", c_int32),
        ("n_bins", c_int32),
        ("vocab_size", c_int64),
        ("min_percentile", c_double),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle (gcc, -fopenmp when libgomp is usable)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    base = ["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-std=c11", "-shared", "-o", _SO, _SRC, "-lm"]
    r = subprocess.run(base[:1] + ["-fopenmp"] + base[1:], capture_output=True, text=True)
    if r.returncode != 0:
        r = subprocess.run(base, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stderr)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.slaf_oracle_tokenize.restype = c_int64
        _lib.slaf_oracle_window.restype = c_int64
        _lib.slaf_oracle_tokenize_lists.restype = c_int64
        _lib.slaf_oracle_num_threads.restype = c_int
    return _lib


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _params(mode, max_genes, max_items, n_bins, vocab_size, min_percentile) -> OracleParams:
    return OracleParams(MODES[mode], int(max_items if max_items is not None else max_genes), int(max_genes),
                        int(n_bins), int(vocab_size), float(min_percentile or 0.0))


def num_threads() -> int:
    return int(lib().slaf_oracle_num_threads())


def tokenize(cell, gene, value, mode: str, max_genes: int, perm=None, max_items=None, n_bins: int = 10,
             vocab_size: int = 50000, min_percentile: float | None = None):
    """COO rows -> (ids int64 [C,L], mask bool [C,L], values int64 [C,L] | None, cell_ids int64 [C])."""
    cell, gene, value = (np.ascontiguousarray(a) for a in (cell, gene, value))
    R = len(cell)
    scgpt = mode.startswith("scgpt")
    L = max_genes + 2 if scgpt else max_genes
    cap = max(R, 1)
    p = _params(mode, max_genes, max_items, n_bins, vocab_size, min_percentile)
    perm_a = None if perm is None else np.ascontiguousarray(np.asarray(perm, np.int64))
    # two passes: the first call sizes C cheaply only when R is small; otherwise allocate by unique count
    C_guess = len(np.unique(cell)) if R else 0
    ids = np.zeros((max(C_guess, 1), L), np.int64)
    mask = np.zeros((max(C_guess, 1), L), np.uint8)
    vals = np.zeros((max(C_guess, 1), L), np.int64) if scgpt else None
    cids = np.zeros(max(C_guess, 1), np.int64)
    C = lib().slaf_oracle_tokenize(_p(cell), DT[cell.dtype], _p(gene), DT[gene.dtype], _p(value), DT[value.dtype],
                                   c_int64(R), _p(perm_a), ctypes.byref(p), c_int64(max(C_guess, 1)),
                                   _p(ids), _p(mask), _p(vals), _p(cids))
    if C < 0:
        raise RuntimeError(f"oracle error {C}")
    assert C == C_guess
    return ids[:C], mask[:C].astype(bool), (vals[:C] if scgpt else None), cids[:C]


def window(cell, gene, value, mode: str, max_items: int, perm=None, n_bins: int = 10,
           min_percentile: float | None = None):
    """COO rows -> (cell_ids [C], offsets [C+1], genes [K] int64, exprs [K] float64)."""
    cell, gene, value = (np.ascontiguousarray(a) for a in (cell, gene, value))
    R = len(cell)
    p = _params(mode, max_items, max_items, n_bins, 0, min_percentile)
    perm_a = None if perm is None else np.ascontiguousarray(np.asarray(perm, np.int64))
    C_guess = len(np.unique(cell)) if R else 0
    offs = np.zeros(C_guess + 1, np.int64)
    og = np.zeros(max(R, 1), np.int64)
    oe = np.zeros(max(R, 1), np.float64)
    cids = np.zeros(max(C_guess, 1), np.int64)
    C = lib().slaf_oracle_window(_p(cell), DT[cell.dtype], _p(gene), DT[gene.dtype], _p(value), DT[value.dtype],
                                 c_int64(R), _p(perm_a), ctypes.byref(p), c_int64(max(C_guess, 1)),
                                 _p(offs), _p(og), _p(oe), _p(cids))
    if C < 0:
        raise RuntimeError(f"oracle error {C}")
    K = int(offs[C])
    return cids[:C], offs[: C + 1], og[:K], oe[:K]


def tokenize_lists(scgpt: bool, offsets, genes, exprs, x_is_int: bool, max_genes: int, n_bins: int = 10,
                   vocab_size: int = 50000):
    offsets = np.ascontiguousarray(offsets, np.int64)
    genes = np.ascontiguousarray(genes, np.int64)
    exprs = np.ascontiguousarray(exprs if exprs is not None else np.zeros(len(genes)), np.float64)
    B = len(offsets) - 1
    L = max_genes + 2 if scgpt else max_genes
    ids = np.zeros((B, L), np.int64)
    mask = np.zeros((B, L), np.uint8)
    vals = np.zeros((B, L), np.int64)
    lib().slaf_oracle_tokenize_lists(c_int(int(scgpt)), _p(offsets), _p(genes), _p(exprs), c_int(int(x_is_int)),
                                     c_int64(B), c_int32(max_genes), c_int32(n_bins), c_int64(vocab_size),
                                     _p(ids), _p(mask), _p(vals))
    return ids, mask.astype(bool), (vals if scgpt else None)
