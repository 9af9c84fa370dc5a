import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
li=int(sys.argv[2])
secs=[];cur=None
i=0
while i < len(rows):
    r=rows[i]
    if r and r[0]=="File Path":
        cur={"file":r[1].split('/')[-1],"hdr":rows[i+2],"rows":[]}; secs.append(cur); i+=3; continue
    if cur is not None: cur["rows"].append(r)
    i+=1
launches=[];seen=set();cur=[]
for s in secs:
    if s["file"] in seen:
        launches.append(cur);cur=[];seen=set()
    seen.add(s["file"]);cur.append(s)
launches.append(cur)
L=launches[li]
agg=collections.Counter(); tot=0
for s in L:
    h=s["hdr"]; ie=h.index("Instructions Executed")
    line=None
    for r in s["rows"]:
        if r[0]!='': line=(s["file"],int(r[0]),r[1].strip()[:90]); continue
        if len(r)>ie and ('LDL' in r[3] or 'STL' in r[3]):
            try: e=int(r[ie])
            except: continue
            agg[line]+=e; tot+=e
print("total local ld/st per warp:", tot/32768)
for k,v in agg.most_common(int(sys.argv[3])): print(f"{v/32768:6.1f}  {k[0]}:{k[1]}  {k[2]}")
