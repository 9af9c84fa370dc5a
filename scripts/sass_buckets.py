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
ia, isrc, ismp, iaddr = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Address")
tot=sum(int(r[ia]) for r in data)
# contiguous regions by similar exec count
b=collections.Counter(); bs=collections.Counter(); bn=collections.Counter()
for r in data:
    e=int(r[ia]); k= 0 if e==0 else round(e/32768,1)
    b[k]+=e; bs[k]+=int(r[ismp]); bn[k]+=1
print("exec/warp  static  dyn%  samples%")
ts=sum(bs.values())
for k,v in sorted(b.items(), key=lambda kv:-kv[1])[:40]:
    print(f"{k:8.1f} {bn[k]:6d} {v/tot*100:6.2f} {bs[k]/ts*100:6.2f}")
if len(sys.argv)>4:
    out=open(sys.argv[4],'w')
    for i,r in enumerate(data):
        out.write(f"{i:5d} {int(r[ia])/32768:7.2f} {int(r[ismp]):5d}  {r[isrc].strip()}\n")
