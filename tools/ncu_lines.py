#!/usr/bin/env python
"""Attribute ncu per-instruction samples to CUDA source lines (needs -lineinfo).

    python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> <launch-skip> [top]

Joins `ncu --page source --csv` (SASS view: samples / executed counts per instruction, in address order) with the
line table printed by `nvdisasm -g` for the same function of the in-tree .so."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kre, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr, data = rows[1], rows[2:]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
isrc = hdr.index("Source")
base = int(data[0][ia], 16)
so = os.path.join(ROOT, "molkgnn_b200", "libmolkgnn_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
# mangled name: find the function whose demangled name matches
line_of = {}
for cb in cubin:
    txt = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
    cur, fn, lineno, f = None, None, None, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            f, lineno = os.path.basename(m.group(1)), int(m.group(2))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m and fn:
            line_of.setdefault(fn, {})[int(m.group(1), 16)] = (f, lineno)
want = re.sub(r"[^A-Za-z0-9_]", "", kname.split("(")[0].split("::")[-1].split("<")[0])
cands = [fn for fn in line_of if want in fn]
best = max(cands, key=lambda fn: len(line_of[fn])) if cands else None
# choose the candidate with the same number of instructions
for fn in cands:
    if len(line_of[fn]) == len(data):
        best = fn
tab = line_of.get(best, {})
agg_s, agg_e = defaultdict(int), defaultdict(int)
tot_s = tot_e = 0
for r in data:
    off = int(r[ia], 16) - base
    key = tab.get(off, ("?", 0))
    s, e = int(r[isamp] or 0), int(r[iex] or 0)
    agg_s[key] += s
    agg_e[key] += e
    tot_s += s
    tot_e += e
print(f"kernel {kname[:60]}  function {best}  instrs {len(data)} (table {len(tab)})  samples {tot_s}  executed {tot_e}")
src_cache = {}
for key in sorted(agg_s, key=lambda k: -agg_s[k])[:top]:
    f, l = key
    txt = ""
    p = os.path.join(ROOT, "molkgnn_b200", "csrc", f)
    if os.path.exists(p):
        src_cache.setdefault(p, open(p).read().splitlines())
        if 0 < l <= len(src_cache[p]):
            txt = src_cache[p][l - 1].strip()[:90]
    print(f"{agg_s[key] / max(tot_s, 1) * 100:5.1f}% smp {agg_e[key] / max(tot_e, 1) * 100:5.1f}% exe  {f}:{l:<4d} {txt}")
