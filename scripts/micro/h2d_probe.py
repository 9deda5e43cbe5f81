"""H2D copy rate from one large page-locked allocation, by offset: does DMA from a multi-GB pinned table hold the rate of the
small buffers bench.py cycles?  python scripts/micro/h2d_probe.py [GiB]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from slaf_b200.engine import pinned_empty

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
n = int(gib * (1 << 30))
t0 = time.perf_counter()
host = pinned_empty(n, np.uint8)
host[::4096] = 1                      # touch every page
print(f"pinned {gib} GiB in {time.perf_counter() - t0:.1f} s")
dev = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ht = torch.from_numpy(host)
chunk = 64 << 20
torch.cuda.synchronize()
for label, offsets in (("first GiB", range(0, 1 << 30, chunk)), ("last GiB", range(n - (1 << 30), n, chunk)), ("whole allocation, sequential", range(0, n - chunk + 1, chunk)),
                       ("whole allocation, 64 MiB strided x7", [(i * 7 * chunk) % (n - chunk) // chunk * chunk for i in range(n // chunk)])):
    offs = list(offsets)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for o in offs:
        dev[:chunk].copy_(ht[o:o + chunk], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"{label}: {len(offs) * chunk / (e0.elapsed_time(e1) / 1e3) / 1e9:.1f} GB/s over {len(offs)} copies of 64 MiB")
t2 = pinned_empty(512 << 20, np.uint8); t2[::4096] = 1
h2 = torch.from_numpy(t2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(32):
    dev[:chunk].copy_(h2[(i % 8) * chunk:(i % 8 + 1) * chunk], non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"separate 512 MiB allocation, cycled: {32 * chunk / (e0.elapsed_time(e1) / 1e3) / 1e9:.1f} GB/s")
