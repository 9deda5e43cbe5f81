"""K1 (csr_build_kernel) back to back on one stream: true per-launch time without per-launch event overhead."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slaf_b200 import synthetic
from slaf_b200.engine import HotPathEngine

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
bs = [synthetic.make_batch(cells, 60000, 2000, seed=10 + i) for i in range(3)]
dev = [tuple(torch.from_numpy(a).cuda() for a in (b.cell, b.gene, b.value)) for b in bs]
eng = HotPathEngine(0, max_rows=max(b.n_rows for b in bs) + 1024, max_cells=cells + 1024)
eng.set_front_on_caller(True)
eng.set_timing_interval(1000000)
for rep in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    slots = [eng.submit(*dev[i]) for i in range(3)]
    e1.record()
    torch.cuda.synchronize()
    for s in slots:
        eng.release(s)
    rows = sum(b.n_rows for b in bs)
    ms = e0.elapsed_time(e1)
    print(f"rep {rep}: 3 x K1 in {ms:.4f} ms -> {ms / 3 * 1000:.1f} us per launch, {4 * rows / (ms / 1e3) / 1e12:.2f} TB/s")
