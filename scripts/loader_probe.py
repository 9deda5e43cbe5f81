"""Where the host feeder's time goes: per-stage wall times of PrefetchBatchProcessor.load_prefetch_batch
over an in-memory table (run on a GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from slaf_b200 import synthetic
from slaf_b200.ml import PrefetchBatchProcessor, RandomShuffle, ScGPTTokenizer
from slaf_b200.ml.sources import ArrayExpressionSource, SyntheticSLAFArray

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
frag_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
pin = (sys.argv[3] != "nopin") if len(sys.argv) > 3 else True
b = synthetic.make_batch(cells, 60000, 2000, seed=3)
src = ArrayExpressionSource(b.cell, b.gene, b.value, max_rows_per_file=frag_rows)
print("pinned:", src.pin() if pin else False, "rows", b.n_rows, "fragments", len(src.get_fragments()))
arr = SyntheticSLAFArray(src, b.cell_start, 60000)
tok = ScGPTTokenizer(arr, vocab_size=65000, n_expression_bins=51, max_genes=1200)

stages = {}


def timed(name, fn):
    def wrap(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        stages.setdefault(name, []).append(time.perf_counter() - t0)
        return r
    return wrap


profile = os.environ.get("SLAFB_PROBE_PROFILE") == "1"       # cProfile of the feeder thread's work instead of the stage timers
modes = tuple(os.environ.get("SLAFB_PROBE_MODES", "fragment,mos,fragment+csi,mos+csi").split(","))
for epoch_mode in modes:
    proc = PrefetchBatchProcessor(arr, RandomShuffle(), tok, n_expression_bins=51, use_binned_expressions=True, n_epochs=3,
                                  use_mixture_of_scanners=epoch_mode.startswith("mos"), by_fragment=True, verbose=False, n_scanners=4,
                                  prefetch_batch_size=4194304, use_cell_start_index=epoch_mode.endswith("csi"))
    if not profile:
        proc._next_pieces = timed("read", proc._next_pieces)
        proc._submit = timed("submit", proc._submit)
        proc._stash_partials = timed("stash", proc._stash_partials)
        proc._complete_mask = timed("complete", proc._complete_mask)
    else:
        import cProfile
        import pstats

        proc.load_prefetch_batch()         # context creation outside the profile
        torch.cuda.synchronize()
        prof = cProfile.Profile()
        prof.enable()
    stages.clear()
    n = 0
    t_first = None
    t0 = time.perf_counter()
    try:
        while True:
            e = proc._engine
            if e is not None and not hasattr(e, "_wrapped") and not profile:
                e.segments = timed("segments", e.segments)
                e.collect = timed("collect", e.collect)
                e._wrapped = True
            out = proc.load_prefetch_batch()
            if t_first is None:
                torch.cuda.synchronize()
                t_first = time.perf_counter()
                n0 = out.n_cells
            n += out.n_cells
    except StopIteration:
        pass
    if profile:
        prof.disable()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if profile:
        pstats.Stats(prof).sort_stats("tottime").print_stats(22)
    print(f"[{epoch_mode}] {n} cells in {t1 - t0:.3f}s ({n / (t1 - t0):,.0f} cells/s incl. start-up; steady {(n - n0) / (t1 - t_first):,.0f} cells/s)")
    for k, v in stages.items():
        print(f"   {k:10s} n={len(v):3d} mean={1e3 * np.mean(v):8.3f} ms  first={1e3 * v[0]:8.3f} ms  total={np.sum(v):.3f}s")
    if proc._engine is not None:
        t = proc._engine.timing()
        print("   library: host_submit %.3f ms/batch, host_collect %.3f ms/batch, wait %.3f ms/batch over %d batches" % (
            t["host_submit_ms"] / max(1, t["batches"]), t["host_collect_ms"] / max(1, t["batches"]), t["host_wait_ms"] / max(1, t["batches"]), t["batches"]))
    proc.close()
