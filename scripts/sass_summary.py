"""profiles/sass_summary.txt: what the built library contains -- per kernel the resource usage (cuobjdump -res-usage) and a histogram
of the SASS opcodes that matter for this path (bulk-copy / mbarrier, 128-bit global loads and stores, shared-memory atomics, shuffles)."""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "slaf_b200", "_lib", "libslaf_b200.so")
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print(f"# {os.path.relpath(so, root)}  ({os.path.getsize(so)} bytes)")
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print("# cubin architectures:", ", ".join(archs))
usage = {}
cur = None
for ln in res.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in ln:
        usage[cur] = ln.strip()
        cur = None
ops = collections.defaultdict(collections.Counter)
cur = None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if cur and m:
        ops[cur][m.group(1)] += 1
KEY = ["UBLKCP", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "LDG.E.128", "STG.E.128", "LDS.128", "LDS.64", "STS.128", "ATOMS", "ATOMG", "RED", "SHFL", "BAR.SYNC", "DMUL", "DFMA", "VIMNMX", "UTMALDG", "UTCHMMA", "HMMA"]
tot = collections.Counter()
for fn in sorted(ops, key=lambda f: demangle(f)):
    c = ops[fn]
    n = sum(c.values())
    fam = collections.Counter()
    for op, k in c.items():
        for key in KEY:
            if op.startswith(key):
                fam[key] += k
                tot[key] += k
    print(f"\n{demangle(fn)[:150]}\n  {usage.get(fn, '')}\n  {n} SASS instructions; " + ", ".join(f"{k}={v}" for k, v in fam.most_common()))
print("\n# whole library: " + ", ".join(f"{k}={v}" for k, v in tot.most_common()))
print("# (UBLKCP = cp.async.bulk, the 1-D TMA bulk copy; SYNCS.* = mbarrier arrive / try_wait with transaction counts; no UTMALDG / UTCHMMA / HMMA: the path has no tensor-shaped tile and no contraction)")
