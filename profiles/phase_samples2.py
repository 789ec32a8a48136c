"""Aggregate ncu stall samples by source-line ranges of ONE file.
usage: phase_samples2.py <source_page.csv> <nvdisasm_g.txt> <kernel substring> <file.cuh> name=lo-hi ..."""
import csv, re, collections, sys
csv_path, dis_path, kname, fname = sys.argv[1:5]
ranges = [(s.split("=")[0],) + tuple(int(v) for v in s.split("=")[1].split("-")) for s in sys.argv[5:]]
rows = list(csv.reader(open(csv_path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[hi], rows[hi + 1:]
ismp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
lines = open(dis_path).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
cur, insts = ("?", 0), []
for l in lines[start + 1:]:
    if (l.startswith(".text.") or l.startswith("\t.section")) and insts: break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): insts.append(cur)
agg, aggx, why = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for r, loc in zip(data, insts):
    s, x = int(r[ismp] or 0), int(r[iex] or 0)
    name = loc[0]
    if loc[0] == fname:
        name = f"{fname} l{loc[1]}"
        for nm, lo, hi_ in ranges:
            if lo <= loc[1] <= hi_: name = nm
    agg[name] += s; aggx[name] += x
    for i, h in stall: why[name][h] += int(r[i] or 0)
tot = sum(agg.values())
print("total samples", tot)
for k, v in agg.most_common(24):
    top = ", ".join(f"{h} {100*c/max(v,1):.0f}%" for h, c in why[k].most_common(3))
    print(f"{k:36s} {v:8d} {100*v/tot:5.1f}%  warp-instrs {aggx[k]:10d}   {top}")
