#!/usr/bin/env python
"""bench.py -- tokenized cells/sec of the SLAF hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg1|cfg3] [--impl reference]

A "step" is one pass of the hot path (shuffle order -> CSR build -> window -> tokenize ->
collate into the batch tensors) over one prefetch batch of `--cells-per-step` complete cells.

  value  : whole-job cells/s with the COO columns already resident in HBM (3 distinct batches
           cycled, each larger than the 126 MB L2), through the library's pipe (slafb_pipe_push:
           one C call per step, outputs into a pre-allocated ring), NO timing events in any
           stream, CUDA events around the K steps on the launch stream, max over ranks.  The
           pipeline is in steady state: warm-up leaves `depth` batches in flight, every timed
           step submits one batch and completes one (K CSR builds + K tokenizations inside
           the window), the tail is drained after the clock stops.
  e2e    : same steps with HOST (pinned) columns: the H2D copy of every step's input and the
           D2H read of its result (cell ids) are inside the timed region.  Headline route =
           the dataset's cell_start_index (gene + value columns + 8 B/cell cross the bus, the
           loader's default); `e2e.coo_route` = the cell column shipped and scanned too.
  roofline: dominant kernel (K2 tokenize_cells_kernel), algorithmic bytes / CUDA-event time
           (a second pipelined pass that carries per-kernel events, and an isolated pass
           where no two kernels overlap), against MEASURED_PEAKS.json's HBM copy bandwidth
           and the nominal 8 TB/s.
  cpu_baseline: the oracle "port" (numpy window + per-cell Python tokenizer loop = the
           reference's structure, polars replaced by numpy) on a bounded sample, all host cores.

`--impl reference` times that CPU port alone (rank 0), same config / metric / unit.
Nothing here reads /root/reference.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
ALL_CPUS = os.sched_getaffinity(0)

WORKLOADS = {
    # BASELINE.json configs[1]: scGPT binned, n_bins=51, max_genes=1200, 60k genes (the metric's 1-GPU config)
    "cfg2": dict(name="cfg2: scGPT binned tokenization (n_bins=51, max_genes=1200, vocab=65000), synthetic 60k-gene COO, ~2k nnz/cell, uint32/uint16/uint16 rows",
                 mode="scgpt_binned", tokenizer="scgpt", max_genes=1200, n_bins=51, vocab=65000, n_genes=60000, mean_nnz=2000, binned=True),
    # configs[0]/[2]: Geneformer, max_genes=2048
    "cfg1": dict(name="cfg1: Geneformer tokenization (max_genes=2048), synthetic 20k-gene COO, ~2k nnz/cell",
                 mode="geneformer", tokenizer="geneformer", max_genes=2048, n_bins=10, vocab=50000, n_genes=20000, mean_nnz=2000, binned=False),
    "cfg3": dict(name="cfg3: Geneformer tokenization (max_genes=2048), synthetic 62k-gene Tahoe-scale COO, ~2k nnz/cell, sharded by fragment range",
                 mode="geneformer", tokenizer="geneformer", max_genes=2048, n_bins=10, vocab=50000, n_genes=62000, mean_nnz=2000, binned=False),
    # the reference benchmark's "Geneformer with percentile filtering" row (docs/benchmarks/ml_benchmarks.md:51): every cell takes
    # the exact k-th-distinct-value path (aggregators.py:167-196)
    "cfg3p": dict(name="cfg3p: Geneformer tokenization with min_percentile=10 (max_genes=2048), synthetic 62k-gene COO, ~2k nnz/cell",
                  mode="geneformer_pct", tokenizer="geneformer", max_genes=2048, n_bins=10, vocab=50000, n_genes=62000, mean_nnz=2000, binned=False,
                  min_percentile=10.0),
    # configs[3] (EXTENSION, not in the reference): per-gene normalisation vectors computed on device (+ all-reduce over the
    # ranks), feeding Geneformer rank-value tokenization in descending normalised-value order
    "cfg4": dict(name="cfg4 (extension): per-gene nonzero means on device + all-reduce, then Geneformer rank-value tokenization "
                      "(descending value/norm[gene], max_genes=2048), synthetic 62k-gene COO, ~2k nnz/cell",
                 mode="geneformer", tokenizer="geneformer", max_genes=2048, n_bins=10, vocab=50000, n_genes=62000, mean_nnz=2000, binned=False,
                 rank_norm=True),
}


def algorithmic_bytes(w, nnz, value_bytes=2, gene_bytes=2):
    """Bytes the algorithm has to move for cells with `nnz` rows each:
    K1 reads the cell column once and writes the segment table; K2 reads every value (the rank
    filter and the bin maximum need them all), the genes that can reach the output
    (min(n, max_genes) per cell -- rows beyond the truncation point are never needed), the segment
    table + permutation entry, and writes the output tensors once.  This is SURVEY.md section 8(d)'s
    figure minus the gene bytes the truncation makes unnecessary, i.e. the conservative one."""
    nnz = np.asarray(nnz, np.int64)
    n_rows, n_cells = int(nnz.sum()), int(nnz.size)
    scgpt = w["tokenizer"] == "scgpt"
    L = w["max_genes"] + 2 if scgpt else w["max_genes"]
    limit = w["max_genes"] if scgpt else w["max_genes"] - 1
    out_per_cell = (17 * L + 8) if scgpt else (9 * L + 8)
    need_values = scgpt | (nnz > w["max_genes"]) | bool(w.get("rank_norm")) | (w.get("min_percentile") is not None)   # Geneformer cells with n <= k never touch their values
    k1 = 4 * n_rows + 8 * n_cells
    k2 = int(value_bytes * nnz[need_values].sum() + gene_bytes * np.minimum(nnz, limit).sum()) + (12 + out_per_cell) * n_cells
    h2d = (4 + gene_bytes + value_bytes) * n_rows
    return k1, k2, h2d


class ClockSampler:
    """SM clock / power / throttle reasons sampled through NVML while the timed regions run
    (same fields as the nvidia-smi line in B200_PROFILING.md; nvidia-smi's own output is block
    buffered on a pipe, so short runs would record nothing)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int, period_s: float = 0.05):
        self.gpu, self.period = gpu_index, period_s
        self.sm, self.power, self.bits = [], [], 0
        self.sm_max = None
        self._stop = threading.Event()
        self._thr = None
        self.err = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception as e:
                self.err = repr(e)
                return
            self._stop.wait(self.period)

    def stop(self) -> dict:
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
               "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm),
               "reasons": sorted(n for b, n in self.REASONS.items() if self.bits & b), "source": "nvml"}
        if self.err:
            out["error"] = self.err
        return out


# ----------------------------------------------------------------------------- CPU port arm
def _port_worker(args):
    (cell, gene, value, w, seed) = args
    from oracle import py_restatement as pr
    t0 = time.perf_counter()
    if w.get("rank_norm"):
        norm = np.ones(w["n_genes"], np.float32)       # the CPU arm is given the vector; only the tokenization is timed
        ids, mask, cids = pr.geneformer_rank_value(cell, gene, value, w["max_genes"], norm=norm, order="rank", seed=seed)
    else:
        extra = {"min_percentile": w["min_percentile"]} if w.get("min_percentile") is not None else {}
        ids, mask, vals, cids = pr.hot_path(cell, gene, value, w["tokenizer"], w["max_genes"], seed=seed, vocab_size=w["vocab"],
                                            n_expression_bins=w["n_bins"], use_binned_expressions=w["binned"], **extra)
    return len(cids), time.perf_counter() - t0


def port_pool(n_proc: int):
    """Worker processes of the CPU arm.  Spawned, not forked: the GPU arm's process holds CUDA state and helper / feeder threads,
    and a forked child can inherit a lock one of them held (a default run of this file once sat in exactly that until its time
    limit)."""
    import multiprocessing as mp

    return mp.get_context("spawn").Pool(n_proc)


def cpu_port_rate(w, n_cells_total: int, n_proc: int, seed: int = 1234, repeats: int = 1, pool=None):
    """cells/s of the CPU port over `n_proc` processes, each on its own range of cells
    (stands in for polars' thread pool; the reference's tokenizer loop is single-threaded Python)."""
    from slaf_b200 import synthetic

    per = max(1, n_cells_total // n_proc)
    jobs = []
    for i in range(n_proc):
        b = synthetic.make_batch(per, w["n_genes"], w["mean_nnz"], seed=seed + i)
        jobs.append((b.cell, b.gene, b.value, w, 42 + i))
    best = None
    own = pool is None
    if own:
        pool = port_pool(n_proc)
    try:
        pool.map(_port_worker, [(j[0][:1000], j[1][:1000], j[2][:1000], w, 1) for j in jobs])  # warm imports
        for _ in range(repeats):
            t0 = time.perf_counter()
            res = pool.map(_port_worker, jobs)
            dt = time.perf_counter() - t0
            cells = sum(r[0] for r in res)
            rate = cells / dt
            best = rate if best is None or rate > best else best
    finally:
        if own:
            pool.close()
            pool.join()
    return best, per * n_proc


def c_oracle_rate(w, n_cells: int, seed: int = 99):
    from oracle import c_oracle
    from slaf_b200 import synthetic

    b = synthetic.make_batch(n_cells, w["n_genes"], w["mean_nnz"], seed=seed)
    t0 = time.perf_counter()
    c_oracle.tokenize(b.cell, b.gene, b.value, w["mode"], w["max_genes"], n_bins=w["n_bins"], vocab_size=w["vocab"])
    dt = time.perf_counter() - t0
    return n_cells / dt, c_oracle.num_threads()


def bench_config(args, w) -> dict:
    """`config` of the JSON line -- identical in both arms (the reference arm's bounded per-step sample is described in its
    cpu_baseline.sample, not here)."""
    return {"workload": w["name"], "cells_per_step_per_gpu": args.cells_per_step, "value_dtype": args.value_dtype,
            "dataset": "every step is one prefetch batch of the synthetic dataset (cells generated from (seed, rank, batch index) with "
                       "the SURVEY section-8d generator); 3 distinct batches per GPU are cycled",
            "l2": "3 distinct batches of ~530 MiB input each are cycled (each > 126 MB L2)",
            "shuffle": "CPython-exact block shuffle, seed 42+step",
            "timing": "steady-state pipeline: every timed step submits one batch and completes one; barrier + synchronize on both sides"}


def pin_rank_cores(local_rank: int, world: int):
    """Give every rank of a multi-GPU run its own disjoint set of whole physical cores (main thread, the library's shuffle
    helpers, NCCL's threads and the feeder all inherit it), so that eight ranks on one 32-vCPU host do not migrate onto each
    other's hyperthread siblings.  Returns the CPU list, or None when there are fewer cores than ranks."""
    cpus = sorted(os.sched_getaffinity(0))
    groups, seen = [], set()
    for c in cpus:
        if c in seen:
            continue
        sib = {c}
        try:
            txt = open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list").read().strip()
            for part in txt.split(","):
                lo, _, hi = part.partition("-")
                sib |= set(range(int(lo), int(hi or lo) + 1))
        except Exception:
            pass
        sib &= set(cpus)
        groups.append(sorted(sib))
        seen |= sib
    per = len(groups) // world
    if per < 1:
        return None
    mine = sorted(c for g in groups[local_rank * per:(local_rank + 1) * per] for c in g)
    os.sched_setaffinity(0, mine)
    return mine


def run_reference(args, w):
    """`--impl reference`: the reference's CPU path (restated; polars/lance are not installable)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    # a bounded sample per step, shrunk for long runs so that the whole arm stays within a few minutes
    per_core = max(64, int(1024 * min(1.0, 40.0 / max(1, args.steps))))
    sample = per_core * cores
    rates = []
    pool = port_pool(cores)
    try:
        for _ in range(args.warmup):
            cpu_port_rate(w, max(cores * 16, 64), cores, pool=pool)
        t_all = time.perf_counter()
        for _ in range(args.steps):
            r, n = cpu_port_rate(w, sample, cores, pool=pool)
            rates.append(r)
    finally:
        pool.close()
        pool.join()
    rate = statistics.median(rates)
    c_rate, c_thr = c_oracle_rate(w, 2048)
    line = {
        "impl": "reference", "metric": "tokenized cells/sec", "value": rate, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sample / rate, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": bench_config(args, w),
        "cpu_baseline": {"value": rate, "unit": "cells/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {sample} cells ({cores} processes x {per_core} cells), numpy window + per-cell Python tokenizer loop (reference structure; polars not installable)",
                         "c_oracle_cells_per_s": c_rate, "c_oracle_threads": c_thr},
        "e2e": {"value": rate, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


def loader_rate(w, batches, dev, value_dtype):
    """cells/s through the drop-in SLAFIterableDataset (what SLAFDataLoader iterates) (host feeder + boundary handling + fused device
    pass + batch slicing), over an in-memory expression table whose fragments cut cells: the call a user of
    the reference makes.  Per-rank figure (every rank runs its own loader over its own shard)."""
    import torch

    from slaf_b200.ml import GeneformerTokenizer, ScGPTTokenizer, SLAFIterableDataset
    from slaf_b200.ml.sources import ArrayExpressionSource, SyntheticSLAFArray

    cell = np.concatenate([b.cell for b in batches])
    gene = np.concatenate([b.gene for b in batches])
    value = np.concatenate([b.value for b in batches])
    first = int(cell[0])
    starts = np.concatenate([[0]] + [b.cell_start[1:] + off for b, off in zip(batches, np.cumsum([0] + [b.n_rows for b in batches[:-1]]))])
    csi = np.zeros(first + len(starts), np.int64)
    csi[first:] = starts
    src = ArrayExpressionSource(cell, gene, value, max_rows_per_file=1 << 24)      # 16.8 M-row fragments, not cell aligned
    pinned = src.pin()
    arr = SyntheticSLAFArray(src, csi, w["n_genes"])
    n_epochs = 12
    out = {"pinned_source": bool(pinned), "fragments": len(src.get_fragments()), "rows": int(len(cell)), "epochs": n_epochs,
           "mode": "by_fragment (one 16.8 M-row fragment per prefetch batch), boundary cells carried on the host; epoch 0 is warm-up, epochs 1.. are timed",
           "routes": "default: cell boundaries from SLAFArray._cell_start_index, the cell column stays on the host (4 B/row over PCIe); "
                     "*_coo_route: use_cell_start_index=False, the cell column is shipped and scanned (8 B/row)"}
    # csi_mode None = the loader's default (the index is used while it validates), False = the cell column is shipped and scanned
    for bs, csi_mode, mos in ((32, None, False), (1024, None, False), (1024, False, False), (1024, None, True)):
        rates = []
        for _rep in range(2):                         # best of two: several loaders in one process disturb each other's first seconds
            tok = (ScGPTTokenizer(arr, vocab_size=w["vocab"], n_expression_bins=w["n_bins"], max_genes=w["max_genes"], device=dev)
                   if w["tokenizer"] == "scgpt" else GeneformerTokenizer(arr, vocab_size=w["vocab"], max_genes=w["max_genes"], device=dev))
            stage(f"  loader batch_size={bs} cell_start_index={csi_mode} mixture_of_scanners={mos} rep={_rep}")
            np.random.seed(17)                           # the scanner draws of the mixture-of-scanners variant
            loader = SLAFIterableDataset(arr, tok, batch_size=bs, use_binned_expressions=w["binned"], use_mixture_of_scanners=mos,
                                         by_fragment=True, verbose=False, device=dev, max_queue_size=4, n_epochs=n_epochs,
                                         use_cell_start_index=csi_mode)              # mos: the reference's defaults (16 scanners x 4,194,304 rows)
            n = 0
            t0 = t1 = None
            last = None
            for batch in loader:
                # steady state between two points of the stream: the first epoch warms up (context creation, staging and output
                # allocations are one-off costs); the clock runs from the first batch of epoch 1 to the first batch of the last
                # epoch, counting the cells received in between (the prefetch queue is equally full at both ends)
                ep = batch["epoch"]
                last = batch
                if ep == 0 or t1 is not None:
                    continue
                if t0 is None:
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                if ep == n_epochs - 1:
                    _ = batch["cell_ids"].cpu()              # D2H read: everything up to this batch has been computed
                    t1 = time.perf_counter()
                    continue
                n += batch["input_ids"].shape[0]
            torch.cuda.synchronize(dev)
            rates.append(n / (t1 - t0))
            del loader, last, batch
        out[f"cells_per_s_batch{bs}" + ("_coo_route" if csi_mode is False else "") + ("_mixture_of_scanners" if mos else "")] = max(rates)
    src.unpin()
    return out


# ----------------------------------------------------------------------------- sharded loader (BASELINE.json config 3)
def table_layout(w, table_cells: int, chunk_cells: int):
    """Row offsets of the synthetic table: (cells per chunk, rows per chunk, cell start index of the whole table)."""
    from slaf_b200 import synthetic

    sizes = [min(chunk_cells, table_cells - lo) for lo in range(0, table_cells, chunk_cells)]
    nnz = [synthetic.batch_nnz(n, w["n_genes"], w["mean_nnz"], seed=5000 + c) for c, n in enumerate(sizes)]
    csi = np.zeros(table_cells + 1, np.int64)
    np.cumsum(np.concatenate(nnz), out=csi[1:])
    return sizes, [int(x.sum()) for x in nnz], csi


def write_table_files(path: str, w, sizes, chunk_rows, files: range, rows_per_file: int, value_dtype: str, rows_per_batch: int = 1 << 20):
    """Write the Arrow IPC fragments `files` of the table: fragment f holds the table rows [f * rows_per_file, (f + 1) * rows_per_file),
    cut without regard to cells or generator chunks (a file that straddles two chunks is assembled from both: the generator is
    deterministic per chunk, so any rank can produce any file)."""
    import pyarrow as pa

    from slaf_b200 import synthetic

    chunk_row0 = np.concatenate(([0], np.cumsum(chunk_rows)))
    total = int(chunk_row0[-1])
    cell0 = np.concatenate(([0], np.cumsum(sizes)))
    cache: dict[int, object] = {}

    def chunk(c):
        if c not in cache:
            if len(cache) > 1:
                cache.pop(min(cache))
            cache[c] = synthetic.make_batch(sizes[c], w["n_genes"], w["mean_nnz"], seed=5000 + c, value_dtype=value_dtype, first_cell_id=int(cell0[c]))
        return cache[c]

    for f in files:
        lo, hi = f * rows_per_file, min((f + 1) * rows_per_file, total)
        parts = []
        c = int(np.searchsorted(chunk_row0, lo, side="right")) - 1
        while lo < hi:
            b = chunk(c)
            a, e = lo - int(chunk_row0[c]), min(hi, int(chunk_row0[c + 1])) - int(chunk_row0[c])
            parts.append((b.cell[a:e], b.gene[a:e], b.value[a:e]))
            lo = int(chunk_row0[c + 1])
            c += 1
        cols = [np.concatenate([p[i] for p in parts]) if len(parts) > 1 else parts[0][i] for i in range(3)]
        table = pa.table({"cell_integer_id": cols[0], "gene_integer_id": cols[1], "value": cols[2]})
        name = os.path.join(path, f"fragment-{f:06d}.arrow")
        with pa.OSFile(name + ".tmp", "wb") as sink, pa.ipc.new_file(sink, table.schema) as writer:
            for rb in table.to_batches(max_chunksize=rows_per_batch):
                writer.write_batch(rb)
        os.replace(name + ".tmp", name)


def run_sharded(args, w):
    """`--sharded`: BASELINE.json config 3 as the reference shards it (slaf/distributed/coordinator.py:40-80): ONE expression table
    (Arrow IPC fragments in /dev/shm, not cell aligned), every rank opens the whole table and reads its contiguous fragment range
    through shard_source + SLAFIterableDataset (boundary cells owned by the rank that holds their first row); no collective on
    the batch path.  Every-cell-exactly-once is proven with (count, sum, xor, sum of squares mod 2^61) of the emitted cell ids over
    all ranks.  One JSON line: aggregate loader cells/s (host-resident input: an end-to-end number), per-rank rates and H2D GB/s."""
    import shutil

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    my_cpus = pin_rank_cores(local_rank, world) if world > 1 and not os.environ.get("SLAFB_BENCH_NO_PIN") else None
    import torch
    import torch.distributed as dist

    from slaf_b200.distributed import shard_rows, shard_source
    from slaf_b200.ml import GeneformerTokenizer, ScGPTTokenizer, SLAFIterableDataset
    from slaf_b200.ml.sources import ArrowExpressionSource, SyntheticSLAFArray

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allreduce(x, op):
        t = torch.tensor(x, dtype=torch.float64 if isinstance(x[0], float) else torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=op)
        return t.tolist()

    stage(f"sharded: table layout ({args.table_cells} cells)")
    sizes, chunk_rows, csi = table_layout(w, args.table_cells, 32768)
    total_rows = int(csi[-1])
    n_files = -(-total_rows // args.rows_per_file)
    path = os.path.join(args.table_dir, f"slafb_table_{os.environ.get('MASTER_PORT', 'single')}")
    if rank == 0:
        shutil.rmtree(path, ignore_errors=True)
        os.makedirs(path)
    barrier()
    base, rem = divmod(n_files, world)
    f0 = rank * base + min(rank, rem)
    t_w = time.perf_counter()
    write_table_files(path, w, sizes, chunk_rows, range(f0, f0 + base + (1 if rank < rem else 0)), args.rows_per_file, args.value_dtype)
    barrier()
    stage(f"sharded: {n_files} fragments written in {time.perf_counter() - t_w:.1f} s (every rank wrote its share)")

    src = ArrowExpressionSource(path)
    rows = [f.count_rows() for f in src.get_fragments()]
    lo, hi = shard_rows(rows, csi, world, rank)
    shard = shard_source(src, csi, world, rank)
    pinned = bool(src.pin()) if not os.environ.get("SLAFB_BENCH_NO_PIN_SOURCE") else False
    cache_s, source_mode = 0.0, "memory-mapped Arrow files, page-locked in place" if pinned else "memory-mapped Arrow files through the pinned staging ring (one host copy per load)"
    if not pinned and not args.no_cache_pinned:
        # the driver refuses to register the files' mappings: the rank's shard is copied ONCE into page-locked memory and every
        # epoch is fed from there by DMA (resident-dataset mode); --no-cache-pinned measures the staged path instead
        from slaf_b200.ml.sources import PinnedCacheSource

        t_c = time.perf_counter()
        shard = PinnedCacheSource(shard, threads=max(2, len(os.sched_getaffinity(0)) // 2))
        cache_s = time.perf_counter() - t_c
        source_mode = f"this rank's shard cached once in page-locked host memory ({shard.nbytes_pinned() >> 20} MiB gene + value columns, {cache_s:.1f} s)"
    arr = SyntheticSLAFArray(shard, csi, w["n_genes"])
    c_lo, c_hi = int(np.searchsorted(csi, lo, side="left")), int(np.searchsorted(csi, hi, side="left"))
    shard_cells = c_hi - c_lo
    n_epochs = max(4, 3 + -(-args.sharded_cells // max(1, shard_cells)))
    tok = (ScGPTTokenizer(arr, vocab_size=w["vocab"], n_expression_bins=w["n_bins"], max_genes=w["max_genes"], device=dev)
           if w["tokenizer"] == "scgpt" else GeneformerTokenizer(arr, vocab_size=w["vocab"], max_genes=w["max_genes"], device=dev))
    loader = SLAFIterableDataset(arr, tok, batch_size=1024, use_binned_expressions=w["binned"], use_mixture_of_scanners=False, by_fragment=True,
                                 verbose=False, device=dev, max_queue_size=4, n_epochs=n_epochs)
    route = "cell_start_index" if loader.batch_processor._csi_arr is not None else "cell column"
    # epoch 0: warm-up + the every-cell-once proof; epoch 1: second warm-up (the feeder runs ahead of the consumer for the first time:
    # allocator growth); epochs 2 .. n-2: timed between the first batch of epoch 2 and the first of the last epoch
    M61 = (1 << 61) - 1
    cnt = torch.zeros((), dtype=torch.int64, device=dev)
    ssum = torch.zeros((), dtype=torch.int64, device=dev)
    sxor = torch.zeros((), dtype=torch.int64, device=dev)
    ssq = torch.zeros((), dtype=torch.int64, device=dev)
    lo_seen = torch.full((), 1 << 62, dtype=torch.int64, device=dev)
    hi_seen = torch.full((), -1, dtype=torch.int64, device=dev)
    n_timed, t0, t1, started = 0, None, None, False
    epoch_marks: dict[int, float] = {}            # wall time at the first batch of every epoch (per-epoch rates in the line)
    stage(f"sharded: rank {rank} reads rows [{lo}, {hi}) = cells [{c_lo}, {c_hi}), {n_epochs} epochs, route {route}, source page-locked: {pinned}")
    for batch in loader:
        ep = batch["epoch"]
        ids = batch["cell_ids"]
        if ep not in epoch_marks:
            epoch_marks[ep] = time.perf_counter()
        if ep <= 1:
            if ep == 1:
                continue                                # second warm-up epoch: the feeder runs ahead for the first time (allocator growth)
            cnt += ids.numel()
            ssum += ids.sum()
            ssq += ((ids % M61) * (ids % M61) % M61).sum()
            for part in ids.split(65536):
                sxor ^= int(np.bitwise_xor.reduce(part.cpu().numpy()))
            lo_seen = torch.minimum(lo_seen, ids.min())
            hi_seen = torch.maximum(hi_seen, ids.max())
            continue
        if t1 is not None:
            continue
        if not started:
            barrier()                                   # all ranks start their timed epochs together
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            started = True
        if ep == n_epochs - 1:
            _ = ids.cpu()                               # D2H read: everything up to this batch has been computed
            t1 = time.perf_counter()
            continue
        n_timed += ids.numel()
    torch.cuda.synchronize(dev)
    dt = t1 - t0
    got = allreduce([int(cnt), int(ssum), int(ssq) % M61], dist.ReduceOp.SUM)
    xors = [torch.zeros((), dtype=torch.int64, device=dev) for _ in range(world)]
    if world > 1:
        dist.all_gather(xors, sxor)
    else:
        xors = [sxor]
    gx = 0
    for x in xors:
        gx ^= int(x)
    n = args.table_cells
    allc = np.arange(n, dtype=np.int64)
    want = [n, n * (n - 1) // 2, int(((allc % M61) * (allc % M61) % M61).sum()), int(np.bitwise_xor.reduce(allc))]
    once = (got[0] == want[0] and got[1] == want[1] and got[2] % M61 == want[2] % M61 and gx == want[3]
            and int(lo_seen) == c_lo and int(hi_seen) == c_hi - 1 and int(cnt) == shard_cells)
    ok_all = allreduce([1 if once else 0], dist.ReduceOp.MIN)[0]
    rows_timed = int(csi[c_hi] - csi[c_lo]) * (n_timed // max(1, shard_cells)) if shard_cells else 0
    b_row = (batch_gene_bytes := 2 if w["n_genes"] <= 65535 else 4) + (2 if args.value_dtype == "uint16" else 4) + (0 if route == "cell_start_index" else 4)
    rates = allreduce([0.0] * rank + [n_timed / dt] + [0.0] * (world - rank - 1), dist.ReduceOp.SUM)
    h2d = allreduce([0.0] * rank + [rows_timed * b_row / dt / 1e9] + [0.0] * (world - rank - 1), dist.ReduceOp.SUM)
    tot = allreduce([float(n_timed)], dist.ReduceOp.SUM)[0]
    tmax = allreduce([float(dt)], dist.ReduceOp.MAX)[0]
    del loader
    barrier()
    src.unpin()
    if rank == 0:
        shutil.rmtree(path, ignore_errors=True)
        if not ok_all:
            raise SystemExit(f"SHARDING FAILURE: the ranks did not emit every cell exactly once (got {got}, xor {gx}; want {want})")
        line = {"metric": "tokenized cells/sec", "value": tot / tmax, "unit": "cells/s", "n_gpus": world, "steps": n_epochs - 3, "warmup": 2,
                "ms_per_step": 1e3 * tmax / max(1, n_epochs - 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
                "data": "synthetic",
                "config": {"workload": w["name"] + " -- ONE table sharded by contiguous fragment range over the ranks (shard_source + SLAFIterableDataset)",
                           "table_cells": n, "table_rows": total_rows, "fragments": n_files, "rows_per_file": args.rows_per_file,
                           "fragments_cell_aligned": False, "value_dtype": args.value_dtype, "route": route, "source": "Arrow IPC files in " + args.table_dir,
                           "source_page_locked": pinned, "source_mode": source_mode, "epochs_timed": n_epochs - 3, "cpus_of_rank0": my_cpus,
                           "note": "a step is one epoch of a rank's shard; host-resident input, so `value` is an end-to-end loader number (H2D inside)"},
                "every_cell_exactly_once": {"checked": "count, sum, sum of squares mod 2^61-1 and xor of the emitted cell ids over all ranks; min / max / count per rank against its shard",
                                            "ok": bool(ok_all), "cells": got[0]},
                "per_rank": {"cells_per_s": [round(x, 1) for x in rates], "h2d_gbs": [round(x, 2) for x in h2d], "seconds": tmax,
                             "rank0_cells_per_s_by_epoch": [round(shard_cells / (epoch_marks[e + 1] - epoch_marks[e]), 1) for e in sorted(epoch_marks)[:-1] if e + 1 in epoch_marks]},
                "e2e": {"value": tot / tmax, "unit": "cells/s", "h2d_bytes_per_step": int(rows_timed * b_row // max(1, n_epochs - 3)), "d2h_bytes_per_step": 8 * 1024}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- B200 arm
_T0 = time.perf_counter()


def stage(msg: str) -> None:
    """progress line on stderr (rank 0): where a run was when it stopped"""
    if os.environ.get("RANK", "0") == "0":
        print(f"[bench {time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cells-per-step", type=int, default=32768)
    ap.add_argument("--value-dtype", default="uint16", choices=["uint16", "float32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-loader", action="store_true")
    ap.add_argument("--no-packed", action="store_true", help="skip the packed-rows leg of e2e")
    ap.add_argument("--verify-cells", type=int, default=2048)
    ap.add_argument("--sharded", action="store_true", help="config 3 as the reference shards it: one table in /dev/shm, every rank reads its fragment range through the loader")
    ap.add_argument("--table-cells", type=int, default=1 << 20, help="--sharded: cells of the table")
    ap.add_argument("--rows-per-file", type=int, default=1 << 24, help="--sharded: rows per fragment file (not cell aligned)")
    ap.add_argument("--sharded-cells", type=int, default=4_000_000, help="--sharded: cells each rank tokenizes in its timed epochs (at least)")
    ap.add_argument("--table-dir", default="/dev/shm", help="--sharded: where the Arrow fragments are written")
    ap.add_argument("--no-cache-pinned", action="store_true", help="--sharded: feed from the memory-mapped files through the staging ring instead of caching the shard in page-locked memory")
    ap.add_argument("--watchdog", type=int, default=None,
                    help="seconds after which every thread's stack is dumped to stderr and the run aborts (default: 780 s, more for very long runs; 0 = off)")
    args = ap.parse_args()
    import faulthandler

    faulthandler.enable()
    if args.watchdog is None:
        args.watchdog = int(max(780, 300 + 0.15 * (args.steps + args.warmup)))
    if args.watchdog > 0:       # a hang must cost a bounded amount of box time and leave a trace of where it sat
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    w = WORKLOADS[args.workload]
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args, w)
    if args.sharded:
        return run_sharded(args, w)

    cpu_baseline = None
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 and not args.no_cpu_baseline:
        # first, while this process is still single-threaded and has no CUDA context, on every host core
        cores = len(os.sched_getaffinity(0))
        stage(f"cpu baseline on {cores} cores")
        rate, n = cpu_port_rate(w, 768 * cores, cores)
        c_rate, c_thr = c_oracle_rate(w, 2048)
        cpu_baseline = {"value": rate, "unit": "cells/s", "cores": cores, "kind": "port",
                        "sample": f"{n} cells ({cores} processes x {n // cores}), numpy window + per-cell Python tokenizer loop (reference structure; polars not installable)",
                        "c_oracle_cells_per_s": c_rate, "c_oracle_threads": c_thr,
                        "published_reference": "5,350 cells/s scGPT / 6,494 cells/s Geneformer on M1 Max (docs/benchmarks/ml_benchmarks.md:49-52)"}

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # one disjoint set of whole cores per rank, taken BEFORE any thread exists (torch, NCCL, the library's helpers inherit it)
    my_cpus = pin_rank_cores(local_rank, world) if world > 1 and not os.environ.get("SLAFB_BENCH_NO_PIN") else None
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ.setdefault(var, "1")

    import torch
    import torch.distributed as dist

    from slaf_b200 import synthetic
    from slaf_b200.engine import HotPathEngine, PreparedInput

    torch.set_num_threads(1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def gather_ranks(xs: list[float]) -> list[list[float]]:
        if world == 1:
            return [xs]
        mine = torch.tensor(xs, dtype=torch.float64, device=dev)
        out = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(out, mine)
        return [[round(float(v), 5) for v in t.tolist()] for t in out]

    # ---- synthetic shard of this rank: NB distinct prefetch batches, cycled (weak scaling: fixed work per GPU)
    NB = 3
    cps = args.cells_per_step
    batches = [synthetic.make_batch(cps, w["n_genes"], w["mean_nnz"], seed=1000 * (rank + 1) + i, value_dtype=args.value_dtype,
                                    first_cell_id=(rank * NB + i) * cps) for i in range(NB)]
    max_rows = max(b.n_rows for b in batches)
    eng = HotPathEngine(dev, max_rows=max_rows + 1024, max_cells=cps + 1024)
    kw = dict(mode=w["mode"], max_genes=w["max_genes"], n_bins=w["n_bins"], vocab_size=w["vocab"])
    if w.get("min_percentile") is not None:
        kw["min_percentile"] = w["min_percentile"]

    dev_cols = [(torch.from_numpy(b.cell).to(dev), torch.from_numpy(b.gene).to(dev), torch.from_numpy(b.value).to(dev)) for b in batches]
    stats = None
    if w.get("rank_norm"):
        # config 4: per-gene nonzero sum / count over this rank's rows (K5 accumulates straight into the reduce buffer),
        # ONE all-reduce over NVLink, mean = sum / count in f64 -> the float32 normaliser every rank uses
        from slaf_b200.distributed import gene_stats, normalisation_vector

        gene_stats(eng, dev_cols[:1], w["n_genes"])        # warm-up (NCCL communicator, kernel load)
        barrier()
        s0, s1, s2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        s0.record()
        total, count = gene_stats(eng, dev_cols, w["n_genes"], mark=s1)
        s2.record()
        barrier()
        norm_dev = normalisation_vector(total, count).contiguous()
        stats = {"ms_stats_plus_allreduce": max_over_ranks(s0.elapsed_time(s2)), "ms_stats_kernels": max_over_ranks(s0.elapsed_time(s1)),
                 "ms_allreduce": max_over_ranks(s1.elapsed_time(s2)), "rows_per_gpu": int(sum(b.n_rows for b in batches)),
                 "allreduce_bytes": 2 * 8 * w["n_genes"], "collective": "NCCL all-reduce (sum, int64) of [2, n_genes]" if world > 1 else "none (1 rank)"}
        kw.update(order="rank", norm=norm_dev)

    # ---- parity gate (bit-exact against the CPU oracle) before any number counts
    stage("parity gate")
    from oracle import c_oracle
    import random as _random

    def oracle_for(vb, n_cells, seed):
        rows_v = int(vb.cell_start[n_cells])
        _random.seed(seed)
        perm = list(range(n_cells)); _random.shuffle(perm)
        if w.get("rank_norm"):
            from oracle import py_restatement as _pr

            w_ids, w_mask, w_cids = _pr.geneformer_rank_value(vb.cell[:rows_v], vb.gene[:rows_v], vb.value[:rows_v], w["max_genes"],
                                                              norm=norm_dev.cpu().numpy(), order="rank", seed=seed)
            return rows_v, (w_ids, w_mask, None, w_cids)
        return rows_v, c_oracle.tokenize(vb.cell[:rows_v], vb.gene[:rows_v], vb.value[:rows_v], w["mode"], w["max_genes"], perm=perm,
                                         n_bins=w["n_bins"], vocab_size=w["vocab"], min_percentile=w.get("min_percentile"))

    def same(got, want) -> bool:
        return (np.array_equal(got.input_ids.cpu().numpy(), want[0]) and np.array_equal(got.attention_mask.cpu().numpy(), want[1])
                and (want[2] is None or np.array_equal(got.values.cpu().numpy(), want[2])) and np.array_equal(got.cell_ids.cpu().numpy(), want[3]))

    nv = min(args.verify_cells, cps, 512 if w.get("rank_norm") else cps)
    rows_v, want = oracle_for(batches[0], nv, 42)
    got = eng.tokenize(batches[0].cell[:rows_v], batches[0].gene[:rows_v], batches[0].value[:rows_v], shuffle_seed=42, **kw)
    if not same(got, want):
        raise SystemExit("PARITY FAILURE: CUDA path differs from the oracle; refusing to report throughput")
    del got

    host_cols = [tuple(torch.from_numpy(a).pin_memory() for a in (b.cell, b.gene, b.value)) for b in batches]
    segs = [(b.cell_start.astype(np.uint32), b.cell[b.cell_start[:-1]].astype(np.uint32)) for b in batches]
    tot_rows = [b.n_rows for b in batches]
    vbytes = 2 if args.value_dtype == "uint16" else 4
    gbytes = batches[0].gene.dtype.itemsize
    per_batch = [algorithmic_bytes(w, np.diff(b.cell_start), vbytes, gbytes) for b in batches]

    def bytes_over(n_steps):   # (K1, K2, H2D) algorithmic bytes of n_steps steps (batches are cycled)
        return tuple(sum(per_batch[i % NB][j] for i in range(n_steps)) for j in range(3))

    DEPTH = int(os.environ.get("SLAFB_BENCH_DEPTH", "4"))
    pipe = eng.pipe(ring=3, capacity_cells=cps + 1024, depth=DEPTH, **kw)
    in_dev = [PreparedInput([c]) for c in dev_cols]
    in_host = [PreparedInput([c]) for c in host_cols]
    in_host_csi = [PreparedInput([c], seg_start=sg[0], seg_cell=sg[1]) for c, sg in zip(host_cols, segs)]
    host_ids = torch.empty(cps + 1024, dtype=torch.int64).pin_memory()

    class Stream:
        """The steady pipeline over cycled inputs: `prime` leaves DEPTH batches in flight, every `step` submits one batch and
        completes the oldest (one slafb_pipe_push), `drain` completes the tail."""

        def __init__(self, inputs):
            self.inputs, self.i, self.cells, self.last = inputs, 0, 0, None

        def prime(self):
            for _ in range(DEPTH):
                assert pipe.push(self.inputs[self.i % NB], 42 + self.i)[0] < 0
                self.i += 1

        def steps(self, n, read_result=False):
            cells = d2h = 0
            push, inputs = pipe.push, self.inputs
            for _ in range(n):
                ri, nc, seq = push(inputs[self.i % NB], 42 + self.i)
                self.i += 1
                cells += nc
                if read_result:                    # D2H read of the step's result: the cell ids, in shuffled order
                    host_ids[:nc].copy_(pipe.tensors[ri][3][:nc])
                    d2h += 8 + 8 * nc
            self.last = (ri, nc, seq)
            return cells, d2h

        def drain(self):
            while pipe.pop()[0] >= 0:
                pass

    def timed(inputs, n_warm, n_steps, read_result, wall, keep_last=False):
        """(cells, seconds [max over ranks], this rank's seconds, d2h bytes, last completed (ring, n, seq))"""
        st = Stream(inputs)
        st.prime()
        st.steps(n_warm, read_result)
        eng.timing()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        cells, d2h = st.steps(n_steps, read_result)
        e1.record()
        torch.cuda.synchronize(dev)
        dt_wall = time.perf_counter() - t0
        tm_ = eng.timing()                  # counters of exactly the timed steps (the drain below is not part of them)
        barrier()
        local = dt_wall if wall else e0.elapsed_time(e1) / 1e3
        ri, nc, seq = st.last
        last = None
        if keep_last:                       # the batch the last timed step completed (the drain below reuses its ring entry)
            out = pipe.batch(ri, nc)
            last = (seq, type(out)(*(t.clone() if t is not None else None for t in (out.input_ids, out.attention_mask, out.values, out.cell_ids)), nc))
        st.drain()
        return cells, max_over_ranks(local), local, d2h, last, tm_

    # ---- value: inputs resident in HBM, no timing events in any stream
    stage("value: warm-up + timed steps, inputs in HBM")
    eng.set_timing_interval(0)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    cells, sec, sec_local, _, last, tm = timed(in_dev, args.warmup, args.steps, False, wall=False, keep_last=(rank == 0))
    ms = sec * 1e3
    total_cells = sum_over_ranks(cells)
    value = total_cells / sec
    per_rank = gather_ranks([1e3 * sec_local / args.steps] + [1000.0 * tm[f"host_{k}_ms"] / max(1, tm["batches"]) for k in ("submit", "collect", "wait")])
    # the batch the last timed step completed, checked in full against the oracle (rank 0; the ring still holds it)
    if rank == 0 and not w.get("rank_norm"):
        seq, got_last = last
        _, want_full = oracle_for(batches[seq % NB], got_last.n_cells, 42 + seq)
        if not same(got_last, want_full):
            raise SystemExit("PARITY FAILURE: a batch of the timed pass differs from the oracle")
        del got_last, want_full
    last = None

    # ---- the same pipelined steps carrying per-kernel CUDA events (what roofline.frac is computed from)
    stage("pipelined pass with per-kernel events")
    eng.set_timing_interval(1)
    n_ev = max(12, min(args.steps, 48))
    tev = timed(in_dev, 3, n_ev, False, wall=False)[5]

    stage("isolated-kernel pass")
    # ---- isolated-kernel pass: same steps, CSR build on the launch stream so no two kernels overlap and every
    #      kernel's CUDA-event time is its own (in the pipelined run above K1 of a later batch shares HBM with K2)
    eng.set_front_on_caller(True)
    n_iso = 12
    for i in range(3 + n_iso):
        if i == 3:
            torch.cuda.synchronize(dev)
            eng.timing()
        eng.collect(eng.submit(*dev_cols[i % NB], shuffle_seed=42 + i), shuffle_seed=42 + i, **kw)
    torch.cuda.synchronize(dev)
    tiso = eng.timing()
    eng.set_front_on_caller(False)
    eng.set_timing_interval(0)

    # ---- e2e: host (pinned) columns, H2D + result read inside the timed region
    e2e = None
    if not args.no_e2e:
        # the two routes give identical tensors: checked once before either is timed
        a = eng.collect(eng.submit(*host_cols[0], shuffle_seed=7), shuffle_seed=7, **kw)
        a_t = [t.clone() for t in (a.input_ids, a.attention_mask, a.cell_ids)] + ([a.values.clone()] if a.values is not None else [])
        b_ = eng.collect(eng.submit_segments([host_cols[0][1:]], *segs[0], shuffle_seed=7), shuffle_seed=7, **kw)
        b_t = [b_.input_ids, b_.attention_mask, b_.cell_ids] + ([b_.values] if b_.values is not None else [])
        if not all(torch.equal(x, y) for x, y in zip(a_t, b_t)):
            raise SystemExit("PARITY FAILURE: cell_start_index route differs from the COO route")
        del a, a_t, b_, b_t
        e2e_steps, e2e_warm = args.steps, max(3, args.warmup)

        stage("e2e: host columns, cell_start_index route (headline)")
        eng.set_timing_interval(8)
        cells_s, dts, _, d2h_s, _, tms = timed(in_host_csi, e2e_warm, e2e_steps, True, wall=True)
        seg_bytes = sum((per_batch[i % NB][2] - 4 * tot_rows[i % NB] + 8 * cps + 4) for i in range(e2e_steps))
        e2e = {"value": sum_over_ranks(cells_s) / dts, "unit": "cells/s", "h2d_bytes_per_step": seg_bytes // e2e_steps,
               "d2h_bytes_per_step": d2h_s // e2e_steps, "ms_per_step": 1000.0 * dts / e2e_steps, "h2d_gbs_per_gpu": seg_bytes / dts / 1e9,
               "h2d_copy_gbs": (seg_bytes * tms["timed_batches"] / max(1, tms["batches"]) / (tms["h2d_ms"] / 1e3) / 1e9) if tms["h2d_ms"] else None,
               "route": "cell_start_index (the loader's default when SLAFArray._cell_start_index validates): gene + value columns and an "
                        "8 B/cell segment table cross the bus, the cell column does not and no CSR build runs; outputs identical to the "
                        "COO route (checked in this run)"}

        stage("e2e: host columns, COO route")
        cells_e, dt, _, d2h_e, _, tme = timed(in_host, e2e_warm, e2e_steps, True, wall=True)
        h2d_bytes = bytes_over(e2e_steps)[2]
        e2e["coo_route"] = {"value": sum_over_ranks(cells_e) / dt, "unit": "cells/s", "h2d_bytes_per_step": h2d_bytes // e2e_steps,
                            "d2h_bytes_per_step": d2h_e // e2e_steps, "ms_per_step": 1000.0 * dt / e2e_steps, "h2d_gbs_per_gpu": h2d_bytes / dt / 1e9,
                            "h2d_copy_gbs": (h2d_bytes * tme["timed_batches"] / max(1, tme["batches"]) / (tme["h2d_ms"] / 1e3) / 1e9) if tme["h2d_ms"] else None,
                            "route": "the reference's COO contract: all three columns cross the bus (8 B/row), K1 builds the segment table on the device"}
        eng.set_timing_interval(0)

        # ---- packed wire format: the batches packed ONCE (as a converter / ingest step would), ~2 B per row on the bus
        if args.value_dtype == "uint16" and not w.get("rank_norm") and not args.no_packed:
            from slaf_b200.engine import pack_rows

            stage("e2e: packed rows (f4)")
            packed, pack_s, pack_in = [], 0.0, 0
            for b in batches:
                ids_b = b.cell[b.cell_start[:-1]]
                t0p = time.perf_counter()
                packed.append(pack_rows(b.gene, b.value, b.cell_start, ids_b, pin=True))
                pack_s += time.perf_counter() - t0p
                pack_in += b.n_rows * (gbytes + vbytes)
            in_packed = [PreparedInput([pk]) for pk in packed]
            pa = eng.collect(eng.submit_packed([packed[0]], shuffle_seed=7), shuffle_seed=7, **kw)
            pb = eng.collect(eng.submit(*host_cols[0], shuffle_seed=7), shuffle_seed=7, **kw)
            pairs = [(pa.input_ids, pb.input_ids), (pa.attention_mask, pb.attention_mask), (pa.cell_ids, pb.cell_ids)] + ([(pa.values, pb.values)] if pa.values is not None else [])
            if not all(torch.equal(x, y) for x, y in pairs):
                raise SystemExit("PARITY FAILURE: packed route differs from the COO route")
            del pa, pb, pairs
            eng.set_timing_interval(8)
            cells_p, dtp, _, d2h_p, _, tmp_ = timed(in_packed, e2e_warm, e2e_steps, True, wall=True)
            eng.set_timing_interval(0)
            pk_bytes = sum(packed[i % NB].nbytes + 12 * cps + 8 for i in range(e2e_steps))
            # the same packed blocks resident in HBM: what the decoding kernel itself sustains
            dev_packed = []
            for pk in packed:
                from slaf_b200.engine import PackedRows

                dev_packed.append(PreparedInput([PackedRows(torch.from_numpy(np.asarray(pk.bytes)).to(dev), pk.blk_off, pk.row_start, pk.seg_cell, pk.gene_dtype)]))
            cells_pd, sec_pd, _, _, _, _ = timed(dev_packed, 3, max(12, min(args.steps, 40)), False, wall=False)
            e2e["packed_route"] = {"value": sum_over_ranks(cells_p) / dtp, "unit": "cells/s", "h2d_bytes_per_step": pk_bytes // e2e_steps,
                                   "d2h_bytes_per_step": d2h_p // e2e_steps, "ms_per_step": 1000.0 * dtp / e2e_steps, "h2d_gbs_per_gpu": pk_bytes / dtp / 1e9,
                                   "bytes_per_row": sum(pk.nbytes for pk in packed) / sum(tot_rows),
                                   "device_resident_cells_per_s": sum_over_ranks(cells_pd) / sec_pd,
                                   "packer": {"gbs_of_input_per_core": pack_in / pack_s / 1e9, "rows_per_s_per_core": sum(tot_rows) / pack_s,
                                              "note": "slafb_pack_rows, one thread, gene + value columns in, blocks out; run once per batch here "
                                                      "(a converter / ingest step), NOT inside the timed region"},
                                   "route": "packed wire format (include/slaf_b200.h): one block per cell, byte deltas of the ascending gene ids + byte "
                                            "values + exception lists, decoded in the tokenizing kernel's shared-memory stage; outputs identical to "
                                            "the COO route (checked in this run)"}
            del dev_packed, in_packed, packed

        if not args.no_loader and not w.get("rank_norm"):
            stage("loader")
            e2e["loader"] = loader_rate(w, batches, dev, args.value_dtype)
            stage("loader done")

    clocks = sampler.stop() if rank == 0 else None
    pipe.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    NOMINAL = 8000.0
    k1_b, k2_b, _ = bytes_over(args.steps)
    k1_e, k2_e, _ = bytes_over(n_ev)
    nt = max(1, tev["timed_batches"])
    k2_ms, k1_ms, gen_ms = tev["tokenize_ms"] / nt, tev["csr_ms"] / nt, tev["slowpath_ms"] / nt
    slow_all = w.get("min_percentile") is not None and args.value_dtype == "float32" and bool(os.environ.get("SLAFB_F32_CTA"))     # percentile on float32 without the warp-per-cell kernel: the general kernel IS the tokenizer
    if slow_all:
        k2_ms, gen_ms = gen_ms, 0.0
        tiso["tokenize_ms"] = tiso["slowpath_ms"]
    k2_gbs = (k2_e / n_ev) / (k2_ms / 1e3) / 1e9 if k2_ms else None
    k1_i, k2_i, _ = bytes_over(n_iso)
    nti = max(1, tiso["timed_batches"])
    iso_k2 = k2_i / n_iso / (tiso["tokenize_ms"] / nti / 1e3) / 1e9
    iso_k1 = k1_i / n_iso / (tiso["csr_ms"] / nti / 1e3) / 1e9
    iso = {"note": "same steps with the CSR build on the launch stream: no two kernels overlap, each CUDA-event time is that kernel alone",
           "launches": int(tiso["timed_batches"]),
           "tokenize_cells_kernel": {"avg_launch_ms": tiso["tokenize_ms"] / nti, "achieved": iso_k2, "frac": iso_k2 / peak, "frac_of_nominal_8tbs": iso_k2 / NOMINAL},
           "csr_build_kernel": {"avg_launch_ms": tiso["csr_ms"] / nti, "achieved": iso_k1, "frac": iso_k1 / peak, "frac_of_nominal_8tbs": iso_k1 / NOMINAL}}
    # which kernel tokenizes this workload (slaf_b200.cu: dispatch_fast): one warp per cell for the percentile mode on uint16 values
    # and for float32 values, one CTA per cell otherwise
    if w.get("rank_norm"):
        k2_name = "tokenize_rank_kernel (per-cell sort)"
    elif slow_all:
        k2_name = "tokenize_general_kernel (exact k-th distinct value for every cell)"
    elif w.get("min_percentile") is not None and args.value_dtype != "float32" and not os.environ.get("SLAFB_PCT_CTA"):
        k2_name = "tokenize_pct_warp_kernel"
    elif args.value_dtype == "float32" and not os.environ.get("SLAFB_F32_CTA"):
        k2_name = "tokenize_f32_warp_kernel"
    else:
        k2_name = "tokenize_cells_kernel"
    iso["kernel"] = k2_name + " (reported under the key tokenize_cells_kernel)"
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of K2 from the committed ncu --set full capture, per cell
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}:{args.value_dtype}")
        traffic = int(tj["k2_dram_bytes_per_cell"] * cps) if tj else None
    except Exception:
        pass
    cfg = bench_config(args, w)
    cfg.update({"rows_per_step_per_gpu": int(np.mean(tot_rows)), "pipeline_depth": DEPTH,
                "parity": f"bit-exact vs oracle on {nv} cells before timing and on the last batch of the timed pass ({cps} cells)",
                "cpus_of_rank0": my_cpus, **({"gene_stats": stats} if stats else {})})
    whole = (k1_b + k2_b) / (ms / 1e3) / 1e9 if world == 1 else None
    line = {
        "metric": "tokenized cells/sec", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "ms_per_step_per_rank": [r[0] for r in per_rank], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": cfg,
        "roofline": {"bound": "hbm", "kernel": k2_name, "achieved": k2_gbs, "peak": peak, "unit": "GB/s",
                     "frac": (k2_gbs / peak) if k2_gbs else None, "frac_isolated": iso["tokenize_cells_kernel"]["frac"],
                     "frac_of_nominal_8tbs": (k2_gbs / NOMINAL) if k2_gbs else None, "traffic": traffic,
                     "traffic_source": "static: profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, per cell x cells per launch), not measured in this run",
                     "note": "frac: a pipelined pass of the same steps that carries per-kernel CUDA events (the throughput pass above carries none); the CSR "
                             "build of a later batch executes inside this kernel's event window on the side stream and shares HBM with it; "
                             "frac_isolated: same steps with no two kernels overlapping",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s", "nominal_peak": NOMINAL,
                     "algorithmic_bytes_per_launch": int(k2_e / n_ev), "avg_launch_ms": k2_ms, "timed_launches": int(tev["timed_batches"]),
                     "csr_build_kernel": {"achieved": iso["csr_build_kernel"]["achieved"], "frac": iso["csr_build_kernel"]["frac"],
                                          "algorithmic_bytes_per_launch": int(k1_b / args.steps), "avg_launch_ms": iso["csr_build_kernel"]["avg_launch_ms"],
                                          "in_pipeline": {"event_window_ms": k1_ms, "note": "on the side stream the kernel queues behind the persistent tokenizing "
                                                          "kernel: this CUDA-event window includes waiting for free SMs, it is not a kernel duration"}},
                     "general_path": {"cells_per_step": tev["slow_cells"] / max(1, tev["batches"]), "avg_launch_ms": gen_ms},
                     "whole_step_hbm_frac": (whole / peak) if whole else None, "whole_step_frac_of_nominal_8tbs": (whole / NOMINAL) if whole else None,
                     "isolated": iso},
        "gpu_launches": int(tm["kernel_launches"]),
        "host_us_per_step": {k: round(1000.0 * tm[f"host_{k}_ms"] / max(1, tm["batches"]), 1) for k in ("submit", "collect", "wait", "shuffle")},
        "per_rank": {"columns": ["ms_per_step", "host_submit_us", "host_collect_us", "host_wait_us"], "rows": per_rank},
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = e2e
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
