"""Attribute ncu warp-stall samples (SASS source page CSV) to CUDA source lines using
nvdisasm -g line markers.  usage: attribute_samples.py <source_page.csv> <nvdisasm_g.txt> <kernel substring>"""
import collections
import csv
import re
import sys

csv_path, dis_path, kname = sys.argv[1:4]
rows = list(csv.reader(open(csv_path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[hi], rows[hi + 1:]
ismp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
lines = open(dis_path).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
cur, insts = ("?", 0), []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section"):
        if insts:
            break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        insts.append(cur)
print("sass rows", len(data), "disasm insts", len(insts))
agg = collections.Counter()
ex = collections.Counter()
cnt = collections.Counter()
for r, loc in zip(data, insts):
    agg[loc] += int(r[ismp] or 0)
    ex[loc] += int(r[iex] or 0)
    cnt[loc] += 1
tot = sum(agg.values())
print("total samples", tot)
for loc, s in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 40):
    print(f"{loc[0]}:{loc[1]:<5d} samples {s:7d} ({100*s/tot:5.1f}%)  sass {cnt[loc]:5d}  executed {ex[loc]}")

# ---- phase summary: samples grouped by (file, function-ish line range) given on the command line as
#      name=file:lo-hi ... after the count argument
if len(sys.argv) > 5:
    print("---- phases")
    for spec in sys.argv[5:]:
        name, rng = spec.split("=")
        f, lr = rng.split(":")
        lo, hi = (int(v) for v in lr.split("-"))
        s = sum(v for (ff, ln), v in agg.items() if ff == f and lo <= ln <= hi)
        print(f"{name:24s} {s:7d} samples ({100*s/tot:5.1f}%)")
