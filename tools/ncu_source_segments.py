#!/usr/bin/env python
"""Phase breakdown of a kernel from `ncu -i rep --page source --csv --print-source sass`: the SASS listing is cut at barriers,
branches and TMA instructions and every segment is printed with its share of the warp-state samples, its top stall reasons and
its opcode mix (the tool behind profiles/r02_pass1_source_profile.md).
Usage: python tools/ncu_source_segments.py <csv> [kernel index = 0] [min samples = 150]; --mix prints the executed-opcode mix."""
import collections, csv, re, sys


def kernels_of(fn):
    out, cur = [], None
    with open(fn) as f:
        for row in csv.reader(f):
            if not row:
                continue
            if row[0] == "Kernel Name":
                cur = {"name": row[1], "hdr": None, "rows": []}; out.append(cur); continue
            if cur is None:
                continue
            if cur["hdr"] is None:
                cur["hdr"] = row; continue
            cur["rows"].append(row)
    return out


def opcode(src):
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src.strip())
    return m.group(2) if m else src[:10]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ks = kernels_of(args[0])
    k = ks[int(args[1]) if len(args) > 1 else 0]
    floor = int(args[2]) if len(args) > 2 else 150
    h = k["hdr"]; ix = {n: i for i, n in enumerate(h)}
    num = lambda r, n: int(r[ix[n]] or 0)
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(num(r, "# Samples") for r in k["rows"]); texec = sum(num(r, "Instructions Executed") for r in k["rows"])
    print(f"# {k['name'][:110]}\n# {len(k['rows'])} SASS instructions, {texec} warp instructions executed, {tot} samples")
    if "--mix" in sys.argv:
        mix, smp = collections.Counter(), collections.Counter()
        for r in k["rows"]:
            op = opcode(r[ix["Source"]]).split(".")[0]
            mix[op] += num(r, "Instructions Executed"); smp[op] += num(r, "# Samples")
        for op, c in mix.most_common(24):
            print(f"  {op:8s} executed {100 * c / texec:5.1f} %   samples {100 * smp[op] / tot:5.1f} %")
        agg = {s[6:]: sum(num(r, s) for r in k["rows"]) for s in stalls}
        print("  stall totals:", ", ".join(f"{n} {100 * v / tot:.1f} %" for n, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 200 > tot))
        return
    seg = None
    def flush(i, why):
        nonlocal seg
        if seg and seg["n"] and seg["samp"] >= floor:
            top = ", ".join(f"{s[6:]} {v}" for s, v in seg["st"].most_common(3))
            ops = ", ".join(f"{o} {c}" for o, c in seg["ops"].most_common(5))
            print(f"[{seg['start']:5d}-{i:5d}] {seg['n']:4d} instr  {100 * seg['samp'] / tot:5.1f} % of samples  ends at {why[:30]:30s} | {top} | {ops}")
        seg = {"start": i + 1, "samp": 0, "n": 0, "st": collections.Counter(), "ops": collections.Counter()}
    flush(-1, "")
    for i, r in enumerate(k["rows"]):
        op = opcode(r[ix["Source"]])
        seg["samp"] += num(r, "# Samples"); seg["n"] += 1; seg["ops"][op.split(".")[0]] += 1
        for s in stalls:
            v = num(r, s)
            if v:
                seg["st"][s] += v
        if op.split(".")[0] in ("BAR", "BRA", "EXIT", "SYNCS", "UTMALDG", "UTMASTG", "WARPSYNC", "CALL", "RET", "DEPBAR"):
            flush(i, r[ix["Source"]].strip())
    flush(len(k["rows"]), "end")


if __name__ == "__main__":
    main()
