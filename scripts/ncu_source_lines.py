import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
nw=float(sys.argv[2]) if len(sys.argv)>2 else 32768.0
secs=[];cur=None
i=0
while i < len(rows):
    r=rows[i]
    if r and r[0]=="File Path":
        cur={"file":r[1].split('/')[-1],"func":rows[i+1][1],"hdr":rows[i+2],"rows":[]}; secs.append(cur); i+=3; continue
    if cur is not None: cur["rows"].append(r)
    i+=1
# group sections into launches: a new launch starts when file repeats
launches=[];seen=set();cur=[]
for s in secs:
    if s["file"] in seen:
        launches.append(cur);cur=[];seen=set()
    seen.add(s["file"]);cur.append(s)
launches.append(cur)
which=[int(x) for x in sys.argv[3].split(',')] if len(sys.argv)>3 else range(len(launches))
top=int(sys.argv[4]) if len(sys.argv)>4 else 60
for li in which:
    L=launches[li]
    agg=[]
    for s in L:
        h=s["hdr"]; ie=h.index("Instructions Executed"); ism=h.index("# Samples")
        for r in s["rows"]:
            if r[0]!='' and len(r)>ie:
                try: agg.append((int(r[ie]), int(r[ism]), s["file"], int(r[0]), r[1].strip()))
                except ValueError: pass
    tot=sum(a[0] for a in agg); ts=sum(a[1] for a in agg)
    print(f"=== launch {li}: {tot/nw:.0f} instr/warp, samples {ts}")
    byfile=collections.Counter()
    for a in agg: byfile[a[2]]+=a[0]
    print({k:round(v/nw) for k,v in byfile.items()})
    for a in sorted(agg,key=lambda a:-a[0])[:top]:
        print(f"{a[0]/nw:7.1f} {a[1]/max(ts,1)*100:5.1f}%  {a[2]}:{a[3]:<4d} {a[4][:110]}")
