#!/usr/bin/env python
"""Aggregate an ncu source page by the function (of ecuda_phases.cuh) each SASS instruction came from.
usage: ncu_by_func.py <report.ncu-rep> <libecuda.so the report was taken with> <mangled kernel substring>"""
import csv, re, subprocess, sys, collections, os, tempfile
rep, lib, kern = sys.argv[1:4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[hi + 1:]:
    if r and r[0] == "Address":  # next launch in the report
        break
    if len(r) >= len(hdr):
        data.append(r)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "host" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
seq, cur, inside = [], ("?", 0), False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        inside = kern in ln; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln): seq.append(cur)
assert len(seq) == len(data), (len(seq), len(data))
# function line ranges of the current source
funcs = []
for fn in ("ecuda_phases.cuh", "ecuda_models.cuh", "ecuda_api.cu", "ecuda_fast.cuh", "ecuda_internal.hpp"):
    lines = open(os.path.join(ROOT, "etol_b200", "csrc", fn)).read().splitlines()
    for i, l in enumerate(lines, 1):
        m = re.match(r"^(?:ECUDA_HD|__device__ __forceinline__|__global__)\s+[\w:<>\*& ]*?(\w+)\(", l)
        if m: funcs.append((fn, i, m.group(1)))
        m = re.match(r"^\s+ECUDA_HD static \w+[\s\*&]*(\w+)\(", l)
        if m: funcs.append((fn, i, "Model::" + m.group(1)))
        m = re.match(r"^\s*k_eval\w*\(const", l)
        if m: funcs.append((fn, i, "kernel_body"))
def owner(c):
    f, l = c
    best = None
    for fn, i, name in funcs:
        if fn == f and i <= l: best = name
    return best or f
inst, samp = collections.Counter(), collections.Counter()
for c, r in zip(seq, data):
    o = owner(c)
    inst[o] += int(r[ie] or 0); samp[o] += int(r[isamp] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti}, samples {ts}")
for o, c in samp.most_common(30):
    print(f"{100*c/ts:5.1f}% samp {100*inst[o]/ti:5.1f}% inst  {o}")
