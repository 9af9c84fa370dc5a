import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2]; which=int(sys.argv[3])
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
sec = [s for s in secs if pat in s["name"]][which]
hdr = sec["rows"][0]; data=[r for r in sec["rows"][1:] if len(r)==len(hdr)]
cols=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot=collections.Counter()
for r in data:
    for i in cols: tot[hdr[i]]+=int(r[i] or 0)
T=sum(tot.values())
for k,v in tot.most_common(): print(f"{k:28s} {v:7d} {v/T*100:5.1f}%")
# which opcodes carry long_sb
il=hdr.index("stall_long_sb"); isrc=hdr.index("Source"); ia=hdr.index("Instructions Executed")
top=sorted(data,key=lambda r:-int(r[il] or 0))[:25]
print("top long_sb:")
for r in top: print(f"  {int(r[il]):5d} exec/warp {int(r[ia])/32768:5.1f}  {r[isrc].strip()[:80]}")
