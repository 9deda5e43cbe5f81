"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/slaf_b200.h declares; host-side control logic (CPython-exact shuffle) is right;
without a GPU the product fails loudly instead of falling back to a CPU path."""

import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

from slaf_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _native.build()
    return _native.lib()


def test_header_symbols_all_exported(lib):
    header = open(os.path.join(ROOT, "include", "slaf_b200.h")).read()
    declared = set(re.findall(r"\b(slafb_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.slafb_version() == _native.ABI_VERSION == 3


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_native.Coo) == 3 * 8 + 8 + 4 * 4
    assert ctypes.sizeof(_native.Params) == 4 * 4 + 8 + 8 + 4 + 4 + 8 + 8 + 8 + 4 + 4 + 8 + 8
    assert ctypes.sizeof(_native.Out) == 5 * 8
    assert ctypes.sizeof(_native.Timing) == 4 * 8 + 5 * 8 + 4 * 8 + 8
    assert ctypes.sizeof(_native.PipeResult) == 4 + 4 + 8 + 8


def test_csi_segments_host_routine_matches_the_index(lib):
    """slafb_csi_segments (host only): the table SLAFArray._cell_start_index implies for row blocks cut anywhere, including
    blocks that continue a cell, skip rows, or arrive out of order; a shifted index is refused by the probes."""
    from slaf_b200 import synthetic
    from slaf_b200.engine import csi_segments

    b = synthetic.make_batch(200, 3000, 40, seed=3, min_nnz=2)
    csi = np.ascontiguousarray(b.cell_start, np.int64)
    rng = np.random.default_rng(1)
    for trial in range(60):
        cuts = np.sort(rng.choice(b.n_rows, size=int(rng.integers(0, 6)), replace=False)).tolist()
        bounds = [0] + cuts + [b.n_rows]
        blocks = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)]
        if trial % 3 == 1 and len(blocks) > 2:
            blocks.pop(1)                                          # a gap: rows skipped
        if trial % 3 == 2 and len(blocks) > 2:
            blocks[0], blocks[1] = blocks[1], blocks[0]            # out of table order
        start, ids = csi_segments([(b.cell[lo:hi], b.gene[lo:hi], b.value[lo:hi]) for lo, hi in blocks], [lo for lo, _ in blocks], csi, check=2)
        cat = np.concatenate([b.cell[lo:hi] for lo, hi in blocks])
        heads = np.flatnonzero(np.r_[True, cat[1:] != cat[:-1]])
        assert start.tolist() == heads.tolist() + [len(cat)] and ids.tolist() == cat[heads].tolist()
    shifted = csi.copy()
    shifted[1:-1] += 2
    with pytest.raises(_native.SlafB200Error, match="cell_start_index does not describe"):
        csi_segments([(b.cell, b.gene, b.value)], [0], shifted, check=1)


@pytest.mark.parametrize("seed", [0, 1, 42, 43, 10042, 2**31 + 5, 2**40 + 7, -9])
def test_shuffle_matches_cpython_random(lib, seed):
    from slaf_b200.engine import shuffle_perm

    for n in (0, 1, 2, 3, 5, 31, 32, 33, 100, 1000, 4097, 20000):
        random.seed(seed)
        want = list(range(n))
        random.shuffle(want)
        assert shuffle_perm(seed, n).tolist() == want, (seed, n)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_gpu_is_a_loud_error_not_a_fallback(lib):
    ctx = ctypes.c_void_p()
    rc = lib.slafb_create(ctypes.byref(ctx), 0, 1 << 20, 1 << 10)
    assert rc == _native.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.slafb_last_error(None)
    from slaf_b200.engine import HotPathEngine

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        HotPathEngine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "slaf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle-free", ""), f"{f} mentions the oracle"


def test_packed_rows_round_trip_and_refusals(lib):
    """slafb_pack_rows / slafb_unpack_rows (host only): lossless for any ascending gene ids and any uint16 values -- big deltas and
    big values through the exception lists, empty cells, one-row cells, int32 gene ids -- at ~2 bytes per row for count data;
    rows whose genes do not ascend are refused, not mangled."""
    from slaf_b200 import synthetic
    from slaf_b200.engine import pack_rows

    rng = np.random.default_rng(0)
    for n_genes, gdt in ((60000, np.uint16), (90000, np.int32)):
        b = synthetic.make_batch(400, n_genes, 300, seed=7, min_nnz=1)
        assert b.gene.dtype == gdt
        value = b.value.copy()
        value[rng.integers(0, len(value), 300)] = rng.integers(255, 65536, 300).astype(np.uint16)      # exceptions, incl. 255 itself
        value[7] = 255
        value[rng.integers(0, len(value), 50)] = 0
        starts = np.concatenate((b.cell_start[:100], [b.cell_start[100]] * 3, b.cell_start[100:]))      # three empty cells
        ids = np.arange(len(starts) - 1, dtype=np.uint32)
        pk = pack_rows(b.gene, value, starts, ids)
        g, v = pk.unpack()
        assert np.array_equal(g, b.gene) and np.array_equal(v, value)
        assert pk.nbytes % 16 == 0 and np.all(np.diff(pk.blk_off.astype(np.int64)) >= 1)
        assert pk.nbytes < 10 * b.n_rows + 64 * len(ids)               # 300 of 60,000+ genes per cell: many gaps need the exception list
        sub = pack_rows(b.gene, value, b.cell_start[50:81], ids[50:80])                                  # a cell range from the middle
        g, v = sub.unpack()
        lo, hi = int(b.cell_start[50]), int(b.cell_start[80])
        assert np.array_equal(g, b.gene[lo:hi]) and np.array_equal(v, value[lo:hi])
        bad = b.gene.copy()
        r = int(b.cell_start[5]) + 1
        bad[r] = bad[r - 1]                                                                             # a duplicate gene inside a cell
        with pytest.raises(_native.SlafB200Error, match="do not ascend"):
            pack_rows(bad, value, b.cell_start, np.arange(b.n_cells, dtype=np.uint32))
    dense = synthetic.make_batch(200, 60000, 2000, seed=8)                 # the benchmark's shape: ~2 bytes per row
    pk = pack_rows(dense.gene, dense.value, dense.cell_start, np.arange(dense.n_cells, dtype=np.uint32))
    assert pk.nbytes < 2.06 * dense.n_rows + 48 * dense.n_cells
    with pytest.raises(TypeError):
        pack_rows(b.gene, b.value.astype(np.float32), b.cell_start, np.arange(b.n_cells, dtype=np.uint32))
    empty = pack_rows(np.zeros(0, np.uint16), np.zeros(0, np.uint16), np.zeros(1, np.int64), np.zeros(0, np.uint32))
    assert empty.n_cells == 0 and empty.nbytes == 0


def test_host_log1pf_restatement_matches_libm(lib):
    """float32 binned window (aggregators.py:96-105 on Float32 columns): polars calls the platform's log1pf.  The library carries a
    restatement of the fdlibm float algorithm (glibc <= 2.39) and the correctly rounded value (glibc >= 2.40) and selects the one
    this host's libm matches; the device code is the same source.  Here: the selected evaluation equals libm's log1pf on 4 M
    floats spread over [0, inf) and on every float of [0.5, 0.5001] -- scripts/micro/log1pf_check.c does all 2^31 in C."""
    libm = ctypes.CDLL("libm.so.6")
    libm.log1pf.restype = ctypes.c_float
    libm.log1pf.argtypes = [ctypes.c_float]
    kind = lib.slafb_debug_log1pf(-1, None, None, 0)
    assert kind in (0, 1)
    lo, hi = int(np.float32(0.5).view(np.uint32)), int(np.float32(0.5001).view(np.uint32))
    bits = np.concatenate([np.arange(0, 0x7F800000, 509, dtype=np.uint32), np.arange(lo, hi, dtype=np.uint32)])
    x = bits.view(np.float32)
    got = np.empty_like(x)
    assert lib.slafb_debug_log1pf(kind, x.ctypes.data, got.ctypes.data, x.size) == kind
    f = np.vectorize(lambda v: libm.log1pf(float(v)), otypes=[np.float32])
    step = max(1, x.size // 400_000)                      # (a ctypes call per value: a 400 k sample of the set, plus the dense range)
    idx = np.unique(np.concatenate([np.arange(0, x.size, step), np.arange(x.size - (hi - lo), x.size)]))
    want = f(x[idx])
    assert np.array_equal(got[idx].view(np.uint32), want.view(np.uint32))
    # the two evaluations do differ (by one ulp, on ~0.4 % of the floats): the selection matters
    other = np.empty_like(x)
    lib.slafb_debug_log1pf(1 - kind, x.ctypes.data, other.ctypes.data, x.size)
    assert 0 < np.count_nonzero(other.view(np.uint32) != got.view(np.uint32)) < x.size // 50
