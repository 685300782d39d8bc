"""Top stalled SASS instructions of one kernel from `ncu --page source --csv`: python tools/top_stalls.py src.csv <kernel substring> [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]; N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
sections = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; sections.append(cur); continue
    if cur is not None: cur["rows"].append(r)
sec = [s for s in sections if want in s["name"]][0]
hdr = sec["rows"][0]; data = [d for d in sec["rows"][1:] if len(d) > 45]
iS, iA = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(d[iA] or 0) for d in data)
rt = {h: sum(int(d[i] or 0) for d in data) for i, h in reasons}
print("kernel:", sec["name"][:90]); print("total samples", tot)
print("by reason:", ", ".join(f"{h[6:]}={100*v/tot:.1f}%" for h, v in sorted(rt.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(data)), key=lambda k: -int(data[k][iA] or 0))[:N]
for k in sorted(order):
    d = data[k]; s = int(d[iA] or 0)
    top = sorted(((int(d[i] or 0), h[6:]) for i, h in reasons), reverse=True)[:2]
    print(f"{k:5d} {100*s/tot:5.2f}%  {d[iS].strip()[:70]:70s} " + " ".join(f"{h}:{v}" for v, h in top if v))
