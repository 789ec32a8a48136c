"""Top SASS instructions of an ncu source-page CSV by stall samples, with their dominant stall reasons and
the CUDA source line (nvdisasm -g).  usage: top_sass.py <source_page.csv> <nvdisasm_g.txt> <kernel substring> [n]"""
import csv, re, sys
csv_path, dis_path, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(csv_path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[hi], rows[hi + 1:]
ismp = hdr.index("# Samples")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
lines = open(dis_path).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
cur, insts = ("?", 0), []
for l in lines[start + 1:]:
    if (l.startswith(".text.") or l.startswith("\t.section")) and insts:
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        insts.append(cur)
tot = sum(int(r[ismp] or 0) for r in data)
agg = {}
for i, h in stall:
    agg[h] = sum(int(r[i] or 0) for r in data)
print("total samples", tot, {k: f"{100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(data)), key=lambda k: -int(data[k][ismp] or 0))[:topn]
for k in order:
    r = data[k]
    reasons = sorted(((int(r[i] or 0), h[6:]) for i, h in stall), reverse=True)[:3]
    loc = insts[k] if k < len(insts) else ("?", 0)
    print(f"{int(r[ismp]):7d} {100*int(r[ismp])/tot:5.1f}%  {loc[0]}:{loc[1]:<4d} {r[1][:70]:70s} " + " ".join(f"{h}={v}" for v, h in reasons if v))
