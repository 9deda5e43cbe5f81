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

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from slaf_b200 import synthetic
from slaf_b200.engine import HotPathEngine, PreparedInput


def run_point(eng, w, kw, cells, nnz, steps, dev, depth=4):
    """One sweep point through the library's pipe (one C call per step, outputs into a ring), steady state: `depth` batches stay in
    flight, every timed step submits one batch and completes one."""
    NB = 3
    batches = [synthetic.make_batch(cells, w["n_genes"], nnz, seed=50 + i, first_cell_id=i * cells) for i in range(NB)]
    dcols = [tuple(torch.from_numpy(a).to(dev) for a in (b.cell, b.gene, b.value)) for b in batches]
    hcols = [tuple(torch.from_numpy(a).pin_memory() for a in (b.cell, b.gene, b.value)) for b in batches]
    segs = [(b.cell_start.astype(np.uint32), b.cell[b.cell_start[:-1]].astype(np.uint32)) for b in batches]
    per = [bench.algorithmic_bytes(w, np.diff(b.cell_start)) for b in batches]
    pipe = eng.pipe(ring=3, capacity_cells=cells + 64, depth=depth, **kw)
    host_ids = torch.empty(cells + 64, dtype=torch.int64).pin_memory()

    def timed(inputs, n, read, wall):
        i = 0
        for _ in range(depth):
            pipe.push(inputs[i % NB], 42 + i); i += 1
        for _ in range(5):
            ri, nc, _s = pipe.push(inputs[i % NB], 42 + i); i += 1
        torch.cuda.synchronize(dev)
        eng.timing()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            ri, nc, _s = pipe.push(inputs[i % NB], 42 + i); i += 1
            if read:
                host_ids[:nc].copy_(pipe.tensors[ri][3][:nc])
        e1.record()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        tm = eng.timing()
        while pipe.pop()[0] >= 0:
            pass
        return (dt if wall else e0.elapsed_time(e1) / 1e3) / n, tm

    eng.set_timing_interval(0)
    sec, tm = timed([PreparedInput([c]) for c in dcols], steps, False, False)
    ms = sec * 1e3
    hbm_bytes = sum(per[i % NB][0] + per[i % NB][1] for i in range(steps)) / steps
    eng.set_timing_interval(2)
    dt_coo, tmc = timed([PreparedInput([c]) for c in hcols], steps, True, True)
    dt_csi, tms = timed([PreparedInput([c], seg_start=sg[0], seg_cell=sg[1]) for c, sg in zip(hcols, segs)], steps, True, True)
    h2d = sum(per[i % NB][2] for i in range(steps)) / steps
    h2d_csi = h2d - 4 * float(np.mean([b.n_rows for b in batches])) + 8 * cells
    copy_gbs = h2d * tmc["timed_batches"] / (tmc["h2d_ms"] / 1e3) / 1e9 if tmc["h2d_ms"] else None             # the copies alone (CUDA events)
    copy_gbs_csi = h2d_csi * tms["timed_batches"] / (tms["h2d_ms"] / 1e3) / 1e9 if tms["h2d_ms"] else None
    pipe.close()
    return {"cells_per_step": cells, "mean_nnz": nnz, "rows_per_step": int(np.mean([b.n_rows for b in batches])),
            "value_cells_per_s": cells / sec, "ms_per_step": ms, "hbm_gbs": hbm_bytes / sec / 1e9,
            "host_us_per_step": round(1000.0 * (tm["host_submit_ms"] + tm["host_collect_ms"] - tm["host_wait_ms"]) / max(1, tm["batches"]), 1),
            "e2e_cells_per_s": cells / dt_csi, "e2e_ms_per_step": dt_csi * 1e3, "h2d_gbs": h2d_csi / dt_csi / 1e9, "h2d_copy_gbs": copy_gbs_csi,
            "h2d_overlap": (h2d_csi / dt_csi / 1e9) / copy_gbs_csi if copy_gbs_csi else None,
            "e2e_coo_cells_per_s": cells / dt_coo, "h2d_gbs_coo": h2d / dt_coo / 1e9,
            "h2d_overlap_coo": (h2d / dt_coo / 1e9) / copy_gbs if copy_gbs else None}


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
    rows = []
    for wl in ("cfg3", "cfg2"):
        w = bench.WORKLOADS[wl]
        kw = dict(mode=w["mode"], max_genes=w["max_genes"], n_bins=w["n_bins"], vocab_size=w["vocab"])
        for nnz in nnzs:
            for cells in batches:
                steps = 300 if cells <= 256 else (100 if cells <= 1024 else 40)
                r = run_point(eng, w, kw, cells, nnz, steps, dev)
                r["workload"] = wl
                r["hbm_frac"] = r["hbm_gbs"] / peak
                rows.append(r)
                print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    md = ["| workload | cells/step | nnz/cell | value cells/s | HBM GB/s (alg.) | of measured peak | host us/step | e2e cells/s (cell-start-index route) | H2D GB/s | H2D overlap | e2e cells/s (COO route) | H2D overlap |",
          "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        md.append(f"| {r['workload']} | {r['cells_per_step']} | {r['mean_nnz']} | {r['value_cells_per_s'] / 1e6:.2f} M | {r['hbm_gbs']:.0f} | {r['hbm_frac']:.2f} | {r['host_us_per_step']} | "
                  f"{r['e2e_cells_per_s'] / 1e6:.3f} M | {r['h2d_gbs']:.1f} | {(r['h2d_overlap'] or 0):.2f} | {r['e2e_coo_cells_per_s'] / 1e6:.3f} M | {(r['h2d_overlap_coo'] or 0):.2f} |")
    open(args.out.replace(".json", ".md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
