"""ctypes binding of libslaf_b200.so (include/slaf_b200.h).

The shared library is built in-tree by ``build()`` (nvcc, sm_100a only) and lives at
``slaf_b200/_lib/libslaf_b200.so``.  There is no fallback of any kind: if the library is
missing, or no sm_100 device is visible when a context is created, the caller gets a
RuntimeError -- the product never computes the hot path on the CPU.
"""

from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("SLAFB_LIB") or os.path.join(_HERE, "_lib", "libslaf_b200.so")   # SLAFB_LIB: A/B builds
SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("slaf_b200.cu", "kernels.cuh", "py_shuffle.h", "pack_rows.h")]
HEADER = os.path.join(_ROOT, "include", "slaf_b200.h")

# slafb_status
OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_BUSY, ERR_NO_DEVICE, ERR_NONCONTIGUOUS = 0, -1, -2, -3, -4, -5, -6
# slafb_dtype
U16, I32, U32, F32 = 0, 1, 2, 3
# slafb_mode
GENEFORMER, GENEFORMER_PCT, SCGPT_RAW, SCGPT_BINNED = 0, 1, 2, 3
LOC_DEVICE, LOC_HOST = 0, 1
SHUFFLE_NONE, SHUFFLE_SEED, SHUFFLE_PERM = 0, 1, 2
FLAG_CHECK_CONTIGUOUS = 1
OPT_TIMING_INTERVAL, OPT_FRONT_ON_CALLER = 1, 2
ORDER_ROW, ORDER_RANK = 0, 1
ABI_VERSION = 3

EXPORTS = [
    "slafb_version", "slafb_create", "slafb_destroy", "slafb_last_error", "slafb_tokenize", "slafb_submit", "slafb_submit_pieces", "slafb_submit_segments", "slafb_host_register", "slafb_host_unregister",
    "slafb_collect", "slafb_collect_csr", "slafb_prepare_shuffle", "slafb_segments", "slafb_window", "slafb_tokenize_lists", "slafb_shuffle_perm", "slafb_gene_stats", "slafb_gene_hist", "slafb_gene_median",
    "slafb_host_alloc", "slafb_host_free", "slafb_get_timing", "slafb_set_option",
    "slafb_submit_csi", "slafb_csi_segments", "slafb_wait_loaded", "slafb_pipe_create", "slafb_pipe_push", "slafb_pipe_pop", "slafb_pipe_destroy", "slafb_debug_check_failures",
    "slafb_pack_bound", "slafb_pack_rows", "slafb_unpack_rows", "slafb_submit_packed", "slafb_pipe_push_packed", "slafb_debug_log1pf",
]


class Coo(Structure):
    _fields_ = [
        ("cell", c_void_p), ("gene", c_void_p), ("value", c_void_p), ("n_rows", c_int64),
        ("cell_dtype", c_int32), ("gene_dtype", c_int32), ("value_dtype", c_int32), ("location", c_int32),
    ]


class Params(Structure):
    _fields_ = [
        ("mode", c_int32), ("max_items", c_int32), ("max_genes", c_int32), ("n_bins", c_int32),
        ("vocab_size", c_int64), ("min_percentile", c_double), ("shuffle", c_int32), ("flags", c_int32),
        ("seed", c_int64), ("perm", POINTER(c_int32)), ("n_perm", c_int64),
        ("window_n_bins", c_int32), ("order", c_int32), ("norm", c_void_p), ("n_norm", c_int64),
    ]


class Out(Structure):
    _fields_ = [
        ("input_ids", c_void_p), ("attention_mask", c_void_p), ("values", c_void_p), ("cell_ids", c_void_p),
        ("capacity_cells", c_int64),
    ]


class Packed(Structure):
    _fields_ = [("bytes", c_void_p), ("n_bytes", c_int64), ("blk_off", c_void_p), ("row_start", c_void_p), ("n_cells", c_int64),
                ("location", c_int32), ("gene_dtype", c_int32)]


class PipeResult(Structure):
    _fields_ = [("ring_index", c_int32), ("reserved", c_int32), ("n_cells", c_int64), ("seq", c_int64)]


class Timing(Structure):
    _fields_ = [
        ("csr_ms", c_double), ("tokenize_ms", c_double), ("slowpath_ms", c_double), ("h2d_ms", c_double),
        ("batches", c_int64), ("kernel_launches", c_int64), ("cells", c_int64), ("rows", c_int64),
        ("slow_cells", c_int64),
        ("host_submit_ms", c_double), ("host_collect_ms", c_double), ("host_wait_ms", c_double), ("host_shuffle_ms", c_double),
        ("timed_batches", c_int64),
    ]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libslaf_b200.so")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a in-tree (so the .so travels with the repo snapshot)."""
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(p) for p in SOURCES + [HEADER])
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH, SOURCES[0],
    ]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    for extra in os.environ.get("SLAFB_NVCC_FLAGS", "").split():
        cmd.insert(1, extra)
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a CC without libgomp specs; nvcc should pick the system g++
    env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """Load the library (never builds implicitly on a GPU box: a missing .so is a loud error)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). slaf_b200 has no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    L.slafb_version.restype = c_int
    L.slafb_last_error.restype = c_char_p
    L.slafb_last_error.argtypes = [c_void_p]
    L.slafb_create.argtypes = [POINTER(c_void_p), c_int, c_int64, c_int64]
    L.slafb_destroy.argtypes = [c_void_p]
    L.slafb_tokenize.argtypes = [c_void_p, POINTER(Coo), POINTER(Params), POINTER(Out), POINTER(c_int64), c_void_p]
    L.slafb_submit.argtypes = [c_void_p, POINTER(Coo), c_void_p, POINTER(c_int32)]
    L.slafb_submit_pieces.argtypes = [c_void_p, POINTER(Coo), c_int32, c_void_p, POINTER(c_int32)]
    L.slafb_submit_segments.argtypes = [c_void_p, POINTER(Coo), c_int32, c_void_p, c_void_p, c_int64, c_void_p, POINTER(c_int32)]
    L.slafb_host_register.argtypes = [c_void_p, c_size_t]
    L.slafb_host_unregister.argtypes = [c_void_p]
    L.slafb_collect_csr.argtypes = [c_void_p, c_int32, POINTER(Params), c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                    POINTER(c_int64), POINTER(c_int64), c_void_p]
    L.slafb_prepare_shuffle.argtypes = [c_void_p, c_int32, c_int64]
    L.slafb_collect.argtypes = [c_void_p, c_int32, POINTER(Params), POINTER(Out), POINTER(c_int64), c_void_p]
    L.slafb_segments.argtypes = [c_void_p, c_int32, POINTER(c_int64), c_void_p, c_void_p, c_int64]
    L.slafb_window.argtypes = [c_void_p, POINTER(Coo), POINTER(Params), c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int64, POINTER(c_int64), POINTER(c_int64), c_void_p]
    L.slafb_tokenize_lists.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32,
                                       c_int32, c_int64, POINTER(Out), c_void_p]
    L.slafb_shuffle_perm.argtypes = [c_int64, c_int64, c_void_p]
    L.slafb_gene_stats.argtypes = [c_void_p, POINTER(Coo), c_int64, c_void_p, c_void_p, c_void_p]
    L.slafb_gene_hist.argtypes = [c_void_p, POINTER(Coo), c_int64, c_int32, c_void_p, c_void_p]
    L.slafb_gene_median.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]
    L.slafb_host_alloc.argtypes = [POINTER(c_void_p), c_size_t]
    L.slafb_host_free.argtypes = [c_void_p]
    L.slafb_set_option.argtypes = [c_void_p, c_int32, c_int64]
    L.slafb_get_timing.argtypes = [c_void_p, POINTER(Timing)]
    L.slafb_submit_csi.argtypes = [c_void_p, POINTER(Coo), c_int32, c_void_p, c_void_p, c_int64, c_int32, c_void_p, POINTER(c_int32),
                                   POINTER(c_int64)]
    L.slafb_csi_segments.argtypes = [POINTER(Coo), c_int32, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int64, POINTER(c_int64)]
    L.slafb_wait_loaded.argtypes = [c_void_p, c_int32]
    L.slafb_pipe_create.argtypes = [c_void_p, POINTER(Params), POINTER(Out), c_int32, c_int32, POINTER(c_void_p)]
    L.slafb_pipe_push.argtypes = [c_void_p, POINTER(Coo), c_int32, c_void_p, c_void_p, c_int64, c_int64, c_void_p, POINTER(PipeResult)]
    L.slafb_pipe_pop.argtypes = [c_void_p, c_void_p, POINTER(PipeResult)]
    L.slafb_pipe_destroy.argtypes = [c_void_p]
    L.slafb_debug_check_failures.argtypes = [c_void_p]
    L.slafb_debug_log1pf.argtypes = [c_int32, c_void_p, c_void_p, c_int64]
    L.slafb_pack_bound.restype = c_int64
    L.slafb_pack_bound.argtypes = [c_int64, c_int64]
    L.slafb_pack_rows.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, POINTER(c_int64)]
    L.slafb_unpack_rows.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p]
    L.slafb_submit_packed.argtypes = [c_void_p, POINTER(Packed), c_int32, c_void_p, c_int32, c_void_p, POINTER(c_int32)]
    L.slafb_pipe_push_packed.argtypes = [c_void_p, POINTER(Packed), c_int32, c_void_p, c_int32, c_int64, c_void_p, POINTER(PipeResult)]
    for name in EXPORTS:
        getattr(L, name)  # AttributeError here = header and library out of sync
    _lib = L
    return L


class SlafB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libslaf_b200 error {code}: {message}")
        self.code = code
        self.message = message


def check(rc: int, ctx: c_void_p | None = None) -> None:
    if rc != OK:
        msg = lib().slafb_last_error(ctx)
        raise SlafB200Error(rc, msg.decode() if msg else "")


__all__ = [name for name in dir() if not name.startswith("_")] + ["byref"]
