#!/usr/bin/env python
"""ncu raw-page CSV (+ optional launch list) -> a markdown summary for profiles/.
Usage: python tools/ncu_md.py raw.csv [launches.csv] title points_per_launch > out.md"""
import csv, re, sys, collections
raw = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 4 else None
title, pts = sys.argv[-2], float(sys.argv[-1])
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
print(f"# {title}\n")
if launches:
    lines = [l for l in open(launches) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    rows = rows[len(rows) // 2:]   # the second repetition (warm tables)
    tot = collections.OrderedDict(); cnt = collections.Counter()
    for r in rows:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("xrftb::", "")
        if "at::" in r["Kernel Name"]:
            n = "[torch: synthetic input] " + n[:40]
        tot[n] = tot.get(n, 0.0) + float(r["Metric Value"].replace(",", "")); cnt[n] += 1
    T = sum(v for k, v in tot.items() if not k.startswith("[torch"))
    print("Launch list of one repetition (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare SHARES):\n")
    print("| kernel | launches | total us | share of our kernels |\n|---|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if k.startswith("[torch"):
            continue
        print(f"| `{k[:100]}` | {cnt[k]} | {v / 1e3:.1f} | {100 * v / T:.1f}% |")
    print(f"\nSum of our kernels: {T / 1e3:.1f} us = {pts / (T / 1e9) / 1e9:.1f} GPoints/s for {pts / 1e6:.1f} M points (under ncu).\n")
rows = list(csv.reader(open(raw))); hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
print("`ncu --set full --clock-control none`, one launch per kernel:\n")
for d in data:
    name = re.sub(r"\(.*", "", d[col["Kernel Name"]]).replace("void ", "")
    print(f"## `{name[:110]}`\n")
    for k in KEYS:
        if k in col and d[col[k]] not in ("", "0", "0.0"):
            print(f"- {k}: {d[col[k]]} {units[col[k]]}")
    try:
        def val(k, scale):
            v = float(d[col[k]].replace(",", "")); return v * scale[units[col[k]].lower()]
        b = val("dram__bytes_read.sum", {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}) + val("dram__bytes_write.sum", {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9})
        t = val("gpu__time_duration.sum", {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1})
        print(f"- **DRAM traffic {b / 1e6:.1f} MB per launch = {b / pts:.2f} B/point; {b / t / 1e9:.0f} GB/s while running**")
    except Exception:
        pass
    print()
