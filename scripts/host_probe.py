"""Host-side cost of one pipelined step (device-resident inputs): wall time of submit() / collect() seen from Python next to the
library's own host counters, with and without the prepared shuffle.  Run on a GPU box."""
import os
import sys
import time
from collections import deque

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from slaf_b200 import synthetic
from slaf_b200.engine import HotPathEngine

cps, steps, depth = 32768, int(sys.argv[1]) if len(sys.argv) > 1 else 300, 4
dev = torch.device("cuda:0")
bs = [synthetic.make_batch(cps, 60000, 2000, seed=5 + i) for i in range(3)]
cols = [tuple(torch.from_numpy(x).to(dev) for x in (b.cell, b.gene, b.value)) for b in bs]
eng = HotPathEngine(dev, max_rows=max(b.n_rows for b in bs) + 1024, max_cells=cps + 1024)
kw = dict(mode="scgpt_binned", max_genes=1200, n_bins=51, vocab_size=65000)
eng.set_timing_interval(8)
for shuffle in (True, False):
    t_sub = t_col = 0.0
    q, nxt = deque(), 0
    for phase, n in (("warm", 20), ("timed", steps)):
        t_sub = t_col = 0.0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            while nxt < 10 ** 9 and len(q) <= depth:
                a = time.perf_counter()
                q.append(eng.submit(*cols[nxt % 3], shuffle_seed=(42 + nxt) if shuffle else None))
                t_sub += time.perf_counter() - a
                nxt += 1
            a = time.perf_counter()
            out = eng.collect(q.popleft(), shuffle_seed=(42 + nxt - len(q) - 1) if shuffle else None, **kw)
            t_col += time.perf_counter() - a
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
    print(f"shuffle={shuffle}: {1e6 * t_all / steps:.1f} us/step wall ({1e6 * t_host / steps:.1f} before the final sync), python submit {1e6 * t_sub / steps:.1f} us, "
          f"python collect {1e6 * t_col / steps:.1f} us", eng.timing())
    while q:
        eng.release(q.popleft())
eng.close()
