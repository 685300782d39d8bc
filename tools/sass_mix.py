import csv,sys,re,collections
want = sys.argv[2] if len(sys.argv) > 2 else ""
rows=list(csv.reader(open(sys.argv[1])))
sections=[]; cur=None
for r in rows:
    if r and r[0]=="Kernel Name": cur={"name":r[1],"rows":[]}; sections.append(cur); continue
    if cur is not None: cur["rows"].append(r)
for sec in sections:
    if want and want not in sec["name"]: continue
    hdr=sec["rows"][0]; data=sec["rows"][1:]
    iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iSt=hdr.index("Warp Stall Sampling (All Samples)")
    mix=collections.Counter(); stall=collections.Counter(); tot=0; tots=0
    for d in data:
        if len(d)<=iE: continue
        s=d[iS].strip(); s=re.sub(r'^@!?U?P\d+\s+','',s)
        op=s.split()[0].split('.')[0] if s else '?'
        if op in('LDG','STG','LDS','STS','LDL','STL','ATOMS','ATOMG','RED'): op=s.split()[0]
        try: n=int(d[iE] or 0); st=int(d[iSt] or 0)
        except: continue
        mix[op]+=n; stall[op]+=st; tot+=n; tots+=st
    print("==",sec["name"][:100]); print("total warp instrs",tot,"static",len(data))
    for op,n in mix.most_common(int(sys.argv[3]) if len(sys.argv)>3 else 30):
        print(f"{op:22s} {n:12d} {100*n/tot:6.2f}%   stall-samples {100*stall[op]/max(tots,1):6.2f}%")
