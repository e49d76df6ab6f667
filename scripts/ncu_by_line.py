#!/usr/bin/env python
"""Join an ncu SASS source page (csv) with nvdisasm -g line info and aggregate executed
instructions / stall samples by CUDA source line.
usage: ncu_by_line.py <src.csv from `ncu -i rep --page source --csv`> <dis.txt from `nvdisasm -g -c cubin`> <mangled kernel substring> [top]"""
import collections
import csv
import re
import sys

src_csv, dis_txt, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# --- nvdisasm: instruction sequence of the kernel with (file,line)
lines = open(dis_txt).read().splitlines()
seq, cur, inside = [], ("?", 0), False
for ln in lines:
    if ln.startswith(".text."):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        seq.append((int(m.group(1), 16), m.group(2).strip(), cur))
# --- ncu rows
rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ie, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
print(f"nvdisasm instrs {len(seq)}  ncu rows {len(data)}", file=sys.stderr)
n = min(len(seq), len(data))
byline, samp = collections.Counter(), collections.Counter()
tot = tots = 0
for (addr, txt, cur), r in zip(seq[:n], data[:n]):
    e, s = int(r[ie] or 0), int(r[isamp] or 0)
    byline[cur] += e
    samp[cur] += s
    tot += e
    tots += s
print(f"total warp instructions {tot}, samples {tots}")
srcs = {}
import os
order = samp.most_common(top) if os.environ.get("SORT") == "samp" else byline.most_common(top)
for (f, l), _ in order:
    c = byline[(f, l)]
    if f not in srcs:
        try:
            srcs[f] = open(f"/root/repo/etol_b200/csrc/{f}").read().splitlines()
        except OSError:
            try:
                srcs[f] = open(f"/root/repo/include/{f}").read().splitlines()
            except OSError:
                srcs[f] = []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print(f"{100 * c / tot:5.1f}% inst {100 * samp[(f, l)] / max(tots, 1):5.1f}% samp  {f}:{l:<4d} {text}")
