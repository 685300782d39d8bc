import csv,sys,re,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iSt=hdr.index("Warp Stall Sampling (All Samples)")
mix=collections.Counter(); stall=collections.Counter(); tot=0; tots=0
for d in data:
    if len(d)<=iE: continue
    s=d[iS].strip(); s=re.sub(r'^@!?U?P\d+\s+','',s)
    op=s.split()[0].split('.')[0] if s else '?'
    if op in('LDG','STG','LDS','STS','LDL','STL','ATOMS','ATOMG','RED'): op=s.split()[0]
    n=int(d[iE] or 0); st=int(d[iSt] or 0)
    mix[op]+=n; stall[op]+=st; tot+=n; tots+=st
print("total warp instrs",tot,"static",len(data))
for op,n in mix.most_common(40):
    print(f"{op:22s} {n:12d} {100*n/tot:6.2f}%   stall-samples {100*stall[op]/max(tots,1):6.2f}%")
