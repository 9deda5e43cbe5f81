"""BASELINE.json config 5: batch-size x nnz-per-cell sweep (1 GPU per process).  For every point:
device-resident throughput (cells/s, algorithmic HBM GB/s over K1+K2) and end-to-end throughput from pinned
host columns (cells/s, H2D GB/s, and the fraction of the H2D copy rate the pipeline sustains = overlap).
Writes one JSON line per point and a markdown table.

    python scripts/sweep.py [--out gpurun_out/sweep.json] [--quick]
"""
import argparse
import json
import os
import sys
import time
from collections import deque

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from slaf_b200 import synthetic
from slaf_b200.engine import HotPathEngine


def run_point(eng, w, kw, cells, nnz, steps, dev):
    NB = 3
    batches = [synthetic.make_batch(cells, w["n_genes"], nnz, seed=50 + i, first_cell_id=i * cells) for i in range(NB)]
    dcols = [tuple(torch.from_numpy(a).to(dev) for a in (b.cell, b.gene, b.value)) for b in batches]
    hcols = [tuple(torch.from_numpy(a).pin_memory() for a in (b.cell, b.gene, b.value)) for b in batches]
    per = [bench.algorithmic_bytes(w, np.diff(b.cell_start)) for b in batches]

    def loop(cols, n, read):
        q = deque()
        nxt = 0
        for i in range(n):
            while nxt < n and len(q) <= 2:
                q.append(eng.submit(*cols[nxt % NB], shuffle_seed=42 + nxt))
                nxt += 1
            out = eng.collect(q.popleft(), shuffle_seed=42 + i, **kw)
            if read:
                out.cell_ids.to("cpu")
        return out

    loop(dcols, 5, False)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop(dcols, steps, False)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    hbm_bytes = sum(per[i % NB][0] + per[i % NB][1] for i in range(steps)) / steps
    loop(hcols, 3, True)
    torch.cuda.synchronize(dev)
    eng.timing()
    t0 = time.perf_counter()
    loop(hcols, steps, True)
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / steps
    tm = eng.timing()
    h2d = sum(per[i % NB][2] for i in range(steps)) / steps
    copy_gbs = (h2d * tm["timed_batches"]) / (tm["h2d_ms"] / 1e3) / 1e9 if tm["h2d_ms"] else None
    return {"cells_per_step": cells, "mean_nnz": nnz, "rows_per_step": int(np.mean([b.n_rows for b in batches])),
            "value_cells_per_s": cells / (ms / 1e3), "ms_per_step": ms, "hbm_gbs": hbm_bytes / (ms / 1e3) / 1e9,
            "e2e_cells_per_s": cells / dt, "e2e_ms_per_step": dt * 1e3, "h2d_gbs": h2d / dt / 1e9, "h2d_copy_gbs": copy_gbs,
            "h2d_overlap": (h2d / dt / 1e9) / copy_gbs if copy_gbs else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    peak = 6533.5
    try:
        peak = float(json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    batches = [32, 256, 1024, 4096] if not args.quick else [32, 4096]
    nnzs = [500, 2000, 8000] if not args.quick else [2000]
    eng = HotPathEngine(dev, max_rows=4096 * 8000 * 2, max_cells=8192)
    eng.set_timing_interval(2)
    rows = []
    for wl in ("cfg3", "cfg2"):
        w = bench.WORKLOADS[wl]
        kw = dict(mode=w["mode"], max_genes=w["max_genes"], n_bins=w["n_bins"], vocab_size=w["vocab"])
        for nnz in nnzs:
            for cells in batches:
                steps = 200 if cells <= 256 else (60 if cells <= 1024 else 24)
                r = run_point(eng, w, kw, cells, nnz, steps, dev)
                r["workload"] = wl
                r["hbm_frac"] = r["hbm_gbs"] / peak
                rows.append(r)
                print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    md = ["| workload | cells/step | nnz/cell | value cells/s | HBM GB/s (alg.) | of measured peak | e2e cells/s | H2D GB/s | H2D overlap |", "|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        md.append(f"| {r['workload']} | {r['cells_per_step']} | {r['mean_nnz']} | {r['value_cells_per_s'] / 1e6:.2f} M | {r['hbm_gbs']:.0f} | {r['hbm_frac']:.2f} | "
                  f"{r['e2e_cells_per_s'] / 1e6:.3f} M | {r['h2d_gbs']:.1f} | {(r['h2d_overlap'] or 0):.2f} |")
    open(args.out.replace(".json", ".md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
