"""Join an `ncu --page source --csv` export (per-SASS-instruction counters) with the line table of the built library
(nvdisasm -g), and print executed warp instructions / stall samples per source line of the kernel.

    python scripts/ncu_by_line.py gpurun_out/ncu_source_<...>.csv <mangled kernel name substring> [top N]
"""
import csv
import re
import subprocess
import sys
import tempfile
import os
from collections import defaultdict

src_csv, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("SLAFB_SO") or os.path.join(root, "slaf_b200", "_lib", "libslaf_b200.so")
src_dir = os.environ.get("SLAFB_SRC") or os.path.join(root, "slaf_b200", "csrc")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instructions of the kernel in order, each with the last seen line marker
lines, cur, inside = [], None, False
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
# an export can hold several kernels ("Kernel Name" row, header row, one row per instruction): take the LAST section whose
# (demangled) kernel name contains $KSEL, or simply the last section
heads = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
ksel = os.environ.get("KSEL", "")
pick = [i for i in heads if ksel in rows[i][1]] or heads
h0 = pick[-1]
h1 = min([i for i in heads if i > h0] + [len(rows)])
hdr, data = rows[h0 + 1], rows[h0 + 2:h1]
ii, isamp, ib = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("stall_barrier")
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
if len(lines) != len(data):
    print(f"warning: {len(lines)} instructions in the disassembly, {len(data)} rows in the csv", file=sys.stderr)
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
for loc, r in zip(lines, data):
    a = agg[loc]
    a[0] += int(r[ii] or 0)
    a[1] += int(r[isamp] or 0)
    for nme in names:
        v = r[hdr.index(nme)]
        if v and v != "0":
            a[2][nme] += int(v)
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total warp instructions {tot_i}  samples {tot_s}")
src_cache = {}
def text(loc):
    if loc is None: return ""
    f, l = loc
    for cand in (os.path.join(src_dir, f),):
        if os.path.exists(cand):
            if cand not in src_cache: src_cache[cand] = open(cand).read().splitlines()
            return src_cache[cand][l - 1].strip()[:100]
    return ""
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1 if os.environ.get("BY_SAMPLES") else 0])[:top]:
    st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*a[0]/tot_i:5.1f}% inst {100*a[1]/max(1,tot_s):5.1f}% samp  {loc}  {text(loc)}   [{st}]")
