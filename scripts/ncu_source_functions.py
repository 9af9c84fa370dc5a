import csv, collections, sys, re
src_dir='/root/repo/beamletoptics.jl_b200/csrc/'
def fn_ranges(fname):
    # function starts: lines at column 0 containing '(' and ending with '{', or template lines
    starts=[]
    for i,l in enumerate(open(src_dir+fname),1):
        if re.match(r'^(template|BMO_D|BMO_NI|BMO_HD|__global__|static|struct|int32_t|inline)',l) and ('(' in l or 'struct' in l):
            m=re.search(r'([A-Za-z_0-9]+)\s*\(',l)
            name=m.group(1) if m else l.split()[1]
            if l.startswith('struct'): name='struct '+l.split()[1]
            starts.append((i,name))
    return starts
ranges={f:fn_ranges(f) for f in ('bmo_geom.cuh','bmo_lean.cuh','bmo_math.cuh','bmo_trace.cu')}
def lookup(f,ln):
    if f not in ranges: return f
    name=f+':top'
    for s,n in ranges[f]:
        if s<=ln: name=n
        else: break
    return f.split('.')[0][4:]+'::'+name
rows=list(csv.reader(open(sys.argv[1])))
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
sel=[int(x) for x in sys.argv[2].split(',')]
tot_all=collections.Counter()
for li in sel:
    b=collections.Counter(); bs=collections.Counter()
    for s in launches[li]:
        h=s["hdr"]; ie=h.index("Instructions Executed"); ism=h.index("# Samples")
        for r in s["rows"]:
            if r[0]!='' and len(r)>ie:
                try: e,sm,ln=int(r[ie]),int(r[ism]),int(r[0])
                except ValueError: continue
                k=lookup(s["file"],ln); b[k]+=e; bs[k]+=sm
    tot=sum(b.values()); ts=sum(bs.values())
    print(f"=== wave {li}: {tot/32768:.0f} instr/warp (line-attributed)")
    for k,v in b.most_common(28):
        print(f"  {v/32768:7.0f} {v/tot*100:5.1f}%  smp {bs[k]/ts*100:5.1f}%  {k}")
