"""What bounds the full loader (SLAFIterableDataset: producer thread + consumer) in cell_start_index mode: per-load time in the
producer thread, time the producer spends blocked on the queue, under a few interpreter / queue settings.  Run on a GPU box."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from slaf_b200 import synthetic
from slaf_b200.ml import ScGPTTokenizer, SLAFIterableDataset
from slaf_b200.ml import datasets as D
from slaf_b200.ml.sources import ArrayExpressionSource, SyntheticSLAFArray

cells = 98304
b = synthetic.make_batch(cells, 60000, 2000, seed=3)
src = ArrayExpressionSource(b.cell, b.gene, b.value, max_rows_per_file=1 << 24)
src.pin()
arr = SyntheticSLAFArray(src, b.cell_start, 60000)
dev = torch.device("cuda:0")

orig = D.PrefetchBatchProcessor.load_prefetch_batch
acc = {"load": 0.0, "n": 0}


def timed_load(self):
    t0 = time.perf_counter()
    try:
        return orig(self)
    finally:
        acc["load"] += time.perf_counter() - t0
        acc["n"] += 1


D.PrefetchBatchProcessor.load_prefetch_batch = timed_load

CONFIGS = {"csi": (True, 1024, 4, None), "csi_sw": (True, 1024, 4, 0.0002), "coo": (False, 1024, 4, None), "coo_sw": (False, 1024, 4, 0.0002),
           "coo32": (False, 32, 4, None), "csi32": (True, 32, 4, None)}
for label in (sys.argv[1:] or ["csi", "coo", "csi", "coo"]):
    csi, bs, qs, sw = CONFIGS[label]
    sys.setswitchinterval(sw if sw else 0.005)
    tok = ScGPTTokenizer(arr, vocab_size=65000, n_expression_bins=51, max_genes=1200, device=dev)
    n_epochs = 8
    loader = SLAFIterableDataset(arr, tok, batch_size=bs, use_binned_expressions=True, use_mixture_of_scanners=False, by_fragment=True,
                                 verbose=False, device=dev, max_queue_size=qs, n_epochs=n_epochs, use_cell_start_index=csi)
    n, t0, t1 = 0, None, None
    for batch in loader:
        ep = batch["epoch"]
        if ep == 0 or t1 is not None:
            continue
        if t0 is None:
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            acc["load"], acc["n"] = 0.0, 0
        if ep == n_epochs - 1:
            _ = batch["cell_ids"].cpu()
            t1 = time.perf_counter()
            snap = dict(acc)
            continue
        n += batch["input_ids"].shape[0]
    print(f"[{label}] {n / (t1 - t0):,.0f} cells/s; producer: {1e3 * snap['load'] / max(1, snap['n']):.3f} ms per load over {snap['n']} loads "
          f"= {100 * snap['load'] / (t1 - t0):.0f}% of the wall time")
    del loader
