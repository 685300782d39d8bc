#!/usr/bin/env python
"""Hot spots of one kernel from `ncu --page source --csv --print-source sass`: stall samples per SASS instruction, grouped
into runs between barriers, plus the top single instructions.  Usage: python tools/sass_hot.py file.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]; data = rows[2:]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(d[isamp]) for d in data)
print(f"total samples {tot}, instructions {len(data)}")
agg = collections.Counter()
for d in data:
    for i, h in stall_cols:
        agg[h] += int(d[i])
print("stall mix:", ", ".join(f"{h[6:]}={100*v/tot:.1f}%" for h, v in agg.most_common(9)))
print("--- top instructions")
for k in sorted(range(len(data)), key=lambda k: -int(data[k][isamp]))[:top]:
    d = data[k]
    why = sorted(((int(d[i]), h[6:]) for i, h in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {100*int(d[isamp])/tot:5.2f}%  x{d[iexec]:>9s}  {d[isrc].strip()[:70]:70s} {why}")
print("--- segments between BAR/SYNCS (share of samples, #instr)")
seg_start = 0; acc = 0; segs = []
for k, d in enumerate(data):
    acc += int(d[isamp])
    s = d[isrc]
    if "BAR.SYNC" in s or "SYNCS.PHASECHK" in s or "BRA" in s.split()[0:1] or "EXIT" in s:
        segs.append((seg_start, k, acc, s.strip()[:40])); seg_start = k + 1; acc = 0
for a, b, c, s in segs:
    if c * 100 / tot >= 0.8:
        print(f"[{a:5d}-{b:5d}] {100*c/tot:5.1f}%  ends with {s}")
