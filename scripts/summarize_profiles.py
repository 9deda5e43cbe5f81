"""Markdown summary of bench.py JSON lines: python scripts/summarize_profiles.py profiles/r2z_bench_*.json ..."""
import json
import os
import sys


def load(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None


rows = []
for path in sys.argv[1:]:
    d = load(path)
    if not d or "value" not in d:
        continue
    r, e = d.get("roofline") or {}, d.get("e2e") or {}
    w = d["config"]["workload"].split(":")[0]
    f = lambda x, s=1e6, n=2: "—" if x in (None, 0) else f"{x / s:.{n}f}"
    rows.append("| {} | {} | {} | {} | {} | {} / {} | {} | {} | {} | {} | {} | `{}` |".format(
        w + (" f32" if d["config"].get("value_dtype") == "float32" else ""), d["n_gpus"], d["steps"], f(d["value"], 1e6, 1), f"{d['ms_per_step']:.4f}",
        f(r.get("frac"), 1, 3), f(r.get("frac_isolated"), 1, 3), f(r.get("whole_step_hbm_frac"), 1, 3),
        f(e.get("value"), 1e6, 2) + (f" ({e.get('h2d_gbs_per_gpu', 0):.1f})" if e.get("value") else ""),
        f((e.get("coo_route") or {}).get("value"), 1e6, 2), f((e.get("packed_route") or {}).get("value"), 1e6, 2),
        f((d.get("cpu_baseline") or {}).get("value"), 1e3, 1) + (" k" if d.get("cpu_baseline") else ""), os.path.basename(path)))
print("| workload | GPUs | steps | value M cells/s | ms/step | K2 frac pipelined / isolated | whole step of HBM peak | e2e M cells/s (H2D GB/s per GPU) | COO route | packed route | CPU port | file |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
print("\n".join(rows))
