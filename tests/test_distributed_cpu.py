"""World-size-2 tests of the multi-GPU plumbing on CPU (gloo): fragment sharding with cell-aligned
ownership, the host feeder per rank, and the per-gene statistics all-reduce contract.  The kernels
themselves need a GPU and are covered by -m gpu; here the per-rank partial statistics come from numpy."""

import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slaf_b200 import synthetic
from slaf_b200.distributed import (RowRangeSource, allreduce_gene_stats, assign_fragments, normalisation_vector, shard_rows,
                                   shard_source)
from slaf_b200.ml.sources import ArrayExpressionSource, SyntheticSLAFArray

N_CELLS, N_GENES, ROWS_PER_FILE = 150, 400, 313


def _data():
    return synthetic.make_batch(N_CELLS, N_GENES, 25, seed=21, min_nnz=2)


def test_assign_fragments_contiguous_and_reference_shuffle():
    parts = assign_fragments(10, 4)
    assert parts == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]                 # remainder to the first ranks
    # shuffle=True is Coordinator.assign_partitions (slaf/distributed/coordinator.py:40-80)
    random.seed(42)
    idx = list(range(10))
    random.shuffle(idx)
    got = assign_fragments(10, 4, seed=42, shuffle=True)
    assert [x for p in got for x in p] == idx and [len(p) for p in got] == [3, 3, 2, 2]
    assert assign_fragments(2, 4) == [[0], [1], [], []]
    with pytest.raises(ValueError):
        assign_fragments(3, 0)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_rows_partition_cells_exactly_once(world):
    b = _data()
    src = ArrayExpressionSource(b.cell, b.gene, b.value, ROWS_PER_FILE)
    rows = [f.count_rows() for f in src.get_fragments()]
    seen = []
    prev_hi = 0
    for r in range(world):
        lo, hi = shard_rows(rows, b.cell_start, world, r)
        if hi > lo:
            assert lo == prev_hi and lo in set(b.cell_start.tolist()) and hi in set(b.cell_start.tolist())
            prev_hi = hi
        shard = RowRangeSource(src, lo, hi)
        cells = np.concatenate([f.to_table().cell for f in shard.get_fragments()]) if hi > lo else np.zeros(0, np.uint32)
        ids, counts = np.unique(cells, return_counts=True)
        assert np.array_equal(counts, np.diff(b.cell_start)[ids])          # every owned cell is complete
        seen.extend(ids.tolist())
    assert sorted(seen) == list(range(N_CELLS))
    assert prev_hi == b.n_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = _data()
        src = ArrayExpressionSource(b.cell, b.gene, b.value, ROWS_PER_FILE)
        shard = shard_source(src, b.cell_start, world, rank)
        # host feeder of this rank over its shard (raw mode: no GPU needed)
        from slaf_b200.ml import SLAFDataLoader

        arr = SyntheticSLAFArray(shard, b.cell_start, N_GENES)
        np.random.seed(100 + rank)
        loader = SLAFDataLoader(arr, raw_mode=True, batch_size=16, n_scanners=2, prefetch_batch_size=1000, verbose=False)
        cells, genes, values = [], [], []
        for batch in loader:
            cells.append(np.asarray(batch["x"]["cell_integer_id"]))
            genes.append(np.asarray(batch["x"]["gene_integer_id"]))
            values.append(np.asarray(batch["x"]["value"]))
        cells, genes, values = (np.concatenate(x) if x else np.zeros(0, np.int64) for x in (cells, genes, values))
        # per-rank partial statistics (numpy stands in for the gene_stats kernel), then the real all-reduce
        buf = torch.zeros((2, N_GENES), dtype=torch.int64)
        buf[0] = torch.from_numpy(np.bincount(genes.astype(np.int64), weights=values.astype(np.float64), minlength=N_GENES).astype(np.int64))
        buf[1] = torch.from_numpy(np.bincount(genes.astype(np.int64), minlength=N_GENES).astype(np.int64))
        allreduce_gene_stats(buf)
        norm = normalisation_vector(buf[0], buf[1])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), cells=np.unique(cells), n_rows=len(cells), sum=buf[0].numpy(),
                 cnt=buf[1].numpy(), norm=norm.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_ranks_shard_feed_and_allreduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    b = _data()
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    # disjoint ownership, complete coverage, no row lost at the shard boundary
    assert not set(res[0]["cells"].tolist()) & set(res[1]["cells"].tolist())
    assert sorted(res[0]["cells"].tolist() + res[1]["cells"].tolist()) == list(range(N_CELLS))
    assert int(res[0]["n_rows"]) + int(res[1]["n_rows"]) == b.n_rows
    assert min(len(res[0]["cells"]), len(res[1]["cells"])) > N_CELLS // 4        # both ranks really worked
    # the reduced statistics equal the single-process ones on every rank, bit for bit
    want_sum = np.bincount(b.gene.astype(np.int64), weights=b.value.astype(np.float64), minlength=N_GENES).astype(np.int64)
    want_cnt = np.bincount(b.gene.astype(np.int64), minlength=N_GENES)
    for r in res:
        assert np.array_equal(r["sum"], want_sum) and np.array_equal(r["cnt"], want_cnt)
    assert np.array_equal(res[0]["norm"], res[1]["norm"])
    want_norm = np.where(want_cnt > 0, want_sum / np.maximum(want_cnt, 1), 1.0).astype(np.float32)
    assert np.array_equal(res[0]["norm"], want_norm)


def test_allreduce_argument_checks_and_single_process_identity():
    t = torch.arange(8, dtype=torch.int64).reshape(2, 4)
    assert allreduce_gene_stats(t.clone()).equal(t)                       # no process group: identity
    with pytest.raises(ValueError):
        allreduce_gene_stats(torch.zeros(4, dtype=torch.int64))
    with pytest.raises(ValueError):
        allreduce_gene_stats(torch.zeros((2, 4), dtype=torch.float32))


def test_coordinator_mirror_matches_reference_assignment():
    """slaf/distributed/coordinator.py:40-80: seeded shuffle of the partition indices, contiguous split, remainder first."""
    from slaf_b200 import distributed as D

    src = ArrayExpressionSource(np.zeros(100, np.uint32), np.zeros(100, np.uint16), np.zeros(100, np.uint16), 10)
    got = D.Coordinator(src, 3).assign_partitions(seed=42)
    random.seed(42)
    idx = list(range(10))
    random.shuffle(idx)
    assert [got[f"worker_{i}"].partition_indices for i in range(3)] == [idx[:4], idx[4:7], idx[7:]]
