"""Aggregate ncu stall samples of the panel kernel by kernel phase.
usage: phase_samples.py <source_page.csv> <nvdisasm_g.txt> <kernel substring> name=lo-hi ..."""
import csv, re, collections, sys
csv_path, dis_path, kname = sys.argv[1:4]
ranges = []
for spec in sys.argv[4:]:
    nm, lr = spec.split("=")
    lo, hi = (int(v) for v in lr.split("-"))
    ranges.append((nm, lo, hi))
rows = list(csv.reader(open(csv_path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[hi], rows[hi + 1:]
ismp = hdr.index("# Samples"); ib = hdr.index("stall_barrier"); iex = hdr.index("Instructions Executed")
lines = open(dis_path).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
cur, insts = ("?", 0), []
for l in lines[start + 1:]:
    if (l.startswith(".text.") or l.startswith("\t.section")) and insts: break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): insts.append(cur)
agg = collections.Counter(); aggb = collections.Counter(); aggx = collections.Counter()
for r, loc in zip(data, insts):
    s = int(r[ismp] or 0); b = int(r[ib] or 0); x = int(r[iex] or 0)
    name = loc[0]
    if loc[0] == 'vbfem_panel.cuh':
        name = "panel.cuh other l%d" % loc[1]
        for nm, lo, hi_ in ranges:
            if lo <= loc[1] <= hi_: name = nm
    agg[name] += s; aggb[name] += b; aggx[name] += x
tot = sum(agg.values())
print("total samples", tot, " per warp (8 warps):", tot // 8)
for k, v in agg.most_common(16):
    print(f"{k:32s} {v:8d} {100*v/tot:5.1f}%   barrier {aggb[k]:7d}  warp-instrs {aggx[k]}")
