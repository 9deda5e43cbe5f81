"""Small fixed workload for ncu: a few steps of one bench workload, one kernel sequence per step
(csr_build_kernel, tokenize_cells_kernel, tokenize_general_kernel | tokenize_rank_kernel), no pipelining.

    python scripts/prof_driver.py <cfg2|cfg1|cfg3|cfg3p|cfg4> <steps> <cells> [uint16|float32]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from slaf_b200 import synthetic
from slaf_b200.engine import HotPathEngine

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cells = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
vdt = sys.argv[4] if len(sys.argv) > 4 else "uint16"
G = 62000
if workload == "cfg2":
    kw = dict(mode="scgpt_binned", max_genes=1200, n_bins=51, vocab_size=65000)
    G = 60000
elif workload == "cfg3p":
    kw = dict(mode="geneformer_pct", max_genes=2048, min_percentile=10.0)
elif workload == "cfg4":
    kw = dict(mode="geneformer", max_genes=2048, order="rank")
else:
    kw = dict(mode="geneformer", max_genes=2048)
    G = 20000 if workload == "cfg1" else 62000
bs = [synthetic.make_batch(cells, G, 2000, seed=10 + i, value_dtype=vdt) for i in range(2)]
dev = [tuple(torch.from_numpy(a).cuda() for a in (b.cell, b.gene, b.value)) for b in bs]
eng = HotPathEngine(0, max_rows=max(b.n_rows for b in bs) + 1024, max_cells=cells + 1024)
if workload == "cfg4":
    from slaf_b200.distributed import gene_stats, normalisation_vector

    total, count = gene_stats(eng, dev, G)
    kw["norm"] = normalisation_vector(total, count).contiguous()
for i in range(2):  # warm-up (module load, allocator)
    eng.tokenize(*dev[i % 2], shuffle_seed=42 + i, **kw)
torch.cuda.synchronize()
eng.timing()
for i in range(steps):
    out = eng.tokenize(*dev[i % 2], shuffle_seed=42 + i, **kw)
    torch.cuda.synchronize()
t = eng.timing()
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in t.items()})
print("per-step ms: csr %.4f tok %.4f gen %.4f" % (t["csr_ms"] / steps, t["tokenize_ms"] / steps, t["slowpath_ms"] / steps))
