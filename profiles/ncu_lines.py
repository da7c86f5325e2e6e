#!/usr/bin/env python
"""Per-source-line digest of an ncu report: executed warp instructions and stall samples per CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> <warp_tiles> [min_inst_per_tile] [demangled-name-hint]
Joins `ncu --page source --print-source sass` (per-SASS counters) with `nvdisasm -g` line info of the same cubin.
"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, lib, kname, wt = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 1.5
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sass_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass_csv)))
# a report may hold several kernels: take the first one whose name contains the demangled hint in argv[6] (default: the first)
kidx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
hint = sys.argv[6] if len(sys.argv) > 6 else ""
k = next(j for j in range(len(kidx) - 1) if hint in rows[kidx[j]][1])
rows = rows[kidx[k]:kidx[k + 1]]
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
    stem = os.environ.get("NCU_LINES_CUBIN", "kernels")  # which translation unit holds the kernel
    cub = next(os.path.join(td, f) for f in os.listdir(td) if f.startswith(stem))
    dis = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l)
cur, seq = None, []
for l in dis[start + 1:]:
    if l.startswith(".text.") and seq:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append(cur)
assert len(seq) == len(data), (len(seq), len(data))
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
ts = sum(int(r[isamp]) for r in data)
for key, r in zip(seq, data):
    a = agg[key]
    a[0] += int(r[ia]); a[1] += int(r[isamp])
    for c in stall_cols:
        if r[c].isdigit(): a[2][hdr[c]] += int(r[c])
src = {}
print(f"total {sum(a[0] for a in agg.values()) / wt:.0f} warp-instructions per warp-tile, {ts} samples")
for key in sorted(agg, key=lambda k: (k is None, k)):
    n, s, st = agg[key]
    if n / wt < thr and s / ts < 0.008: continue
    text = ""
    if key:
        path = os.path.join(root, "gtars_b200/csrc/cuda", key[0])
        if key[0] not in src and os.path.exists(path): src[key[0]] = open(path).read().split("\n")
        if key[0] in src and key[1] - 1 < len(src[key[0]]): text = src[key[0]][key[1] - 1].strip()[:90]
    top = ",".join(f"{k[6:]}={v * 100 // max(s, 1)}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(f"{str(key):26s} {n / wt:6.1f} {s / ts * 100:5.1f}%  {top:34s} {text}")
