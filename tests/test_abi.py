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
