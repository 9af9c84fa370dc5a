#!/usr/bin/env python
"""Opcode mix and hot spots from `ncu -i rep --page source --csv` (first kernel section matching argv[2])."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2] if len(sys.argv) > 2 else ""
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
sec = [s for s in secs if pat in s["name"]][int(sys.argv[3]) if len(sys.argv) > 3 else 0]
hdr, data = sec["rows"][0], [r for r in sec["rows"][1:] if len(r) == len(sec["rows"][0])]
print(sec["name"], "sections:", len(secs))
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[ismp]) for r in data)
print("static instrs", len(data), "executed", tot, "samples", tots)
op, ops = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc]); o = m.group(2).split(".")[0] if m else "?"
    op[o] += int(r[ia]); ops[o] += int(r[ismp])
for o, c in op.most_common(24):
    print(f"{o:10s} {c/tot*100:6.2f}% exec   {ops[o]/max(tots,1)*100:6.2f}% samples")
hot = [r for r in data if int(r[ia]) > 0]
print("static instrs executed at least once:", len(hot))
hs = sorted((int(r[ia]) for r in hot), reverse=True); c = 0
for i, v in enumerate(hs):
    c += v
    if c > 0.9 * tot:
        print("90% of dynamic instrs in", i + 1, "static instrs"); break
print("top sampled instructions:")
for r in sorted(data, key=lambda r: -int(r[ismp]))[:25]:
    print(f"  {int(r[ismp]):6d} smp {int(r[ia]):9d} exec  {r[isrc].strip()[:90]}")
