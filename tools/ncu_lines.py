#!/usr/bin/env python
"""Attribute an ncu --set full capture's per-instruction counts to CUDA source lines (needs -lineinfo).
usage: tools/ncu_lines.py <rep> <kernel-mangled-substring> [top]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import Counter
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "scribe_py_b200/lib/libscribe_b200.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
amap, cur, infn = {}, None, False
for ln in dis.split("\n"):
    if ln.startswith("//--------------------- .text."):
        infn = sub in ln; cur = None; continue
    if not infn:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", ln)
    if m and cur is not None:
        amap[int(m.group(1), 16)] = cur
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
h = next(i for i, r in enumerate(src) if r and r[0] == "Address"); cols = src[h]
ia, iex, ismp = cols.index("Address"), cols.index("Instructions Executed"), cols.index("# Samples")
rows = [r for r in src[h + 1:] if len(r) > iex and r[iex].isdigit()]
base = int(rows[0][ia], 16)
ex, sm = Counter(), Counter()
for r in rows:
    a = int(r[ia], 16) - base
    ex[amap.get(a, ("?", 0))] += int(r[iex]); sm[amap.get(a, ("?", 0))] += int(r[ismp] or 0)
tot, ts = sum(ex.values()), sum(sm.values()) or 1
files = {}
def text(f, l):
    if f not in files:
        q = os.path.join(ROOT, "scribe_py_b200/csrc", f)
        files[f] = open(q).read().split("\n") if os.path.exists(q) else None
    return files[f][l - 1].strip()[:110] if files[f] and 0 < l <= len(files[f]) else "?"
print("total warp instructions %.4g; mapped %d addresses" % (tot, len(amap)))
for (f, l), n in ex.most_common(top):
    print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100.0 * n / tot, 100.0 * sm[(f, l)] / ts, f, l, text(f, l)))
