#!/usr/bin/env python
"""Summarise ncu output into profiles/: launch list shares + per-kernel DRAM traffic / stall picture."""
import csv, json, re, sys, collections
launch_csv, raw_csv, out_md, traffic_json, pts_per_launch = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
lines = [l for l in open(launch_csv) if not l.startswith("==")]
tot = collections.OrderedDict(); cnt = collections.Counter()
for row in csv.DictReader(lines):
    n = row["Kernel Name"]
    short = re.sub(r"\(.*", "", n)
    short = re.sub(r"xrftb::", "", short)
    if "at::" in n or "at_cuda" in n or "elementwise" in n: short = "[torch] " + short[:60]
    tot[short] = tot.get(short, 0.0) + float(row["Metric Value"].replace(",", "")); cnt[short] += 1
total = sum(tot.values())
md = ["# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)", "",
      "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    md.append(f"| `{k[:110]}` | {cnt[k]} | {v/1e3:.1f} | {100*v/total:.1f}% |")
rows = list(csv.reader(open(raw_csv))); hdr, units, data = rows[0], rows[1], rows[2:]
def col(name): return hdr.index(name)
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
traffic = {}
md += ["", "# ncu --set full (one launch per kernel class)", ""]
for d in data:
    name = re.sub(r"\(.*", "", d[col("Kernel Name")]).replace("void ", "")
    md.append(f"## `{name[:120]}`")
    md.append("")
    vals = {}
    for k in keys:
        if k in hdr:
            vals[k] = d[col(k)]; md.append(f"- {k}: {d[col(k)]} {units[col(k)]}")
    try:
        def tobytes(k):
            v = float(vals[k].replace(",", "")); u = units[col(k)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        def tosec(k):
            v = float(vals[k].replace(",", "")); u = units[col(k)].lower()
            return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[u]
        b = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
        t = tosec("gpu__time_duration.sum")
        md.append(f"- **dram traffic per launch: {b/1e6:.1f} MB = {b/pts_per_launch:.2f} B/point; {b/t/1e9:.0f} GB/s while running**")
        cls = ("moments_kernel" if "moments" in name else
               "rowsz_power_kernel (pass 2: column separation + row FFT + |F|^2 + mirrored row)" if "rowsz" in name else
               "cols_async_kernel<ColsR2CPack> (pass 1: detrend + window + packed column FFT, TMA in / TMA out)" if "cols_async_kernel" in name and "ColsR2CPack" in name else
               "rows2_kernel<RowsR2CFused>" if "rows" in name else "cols_kernel<ColsFused POWER>" if "cols_kernel" in name else "mirror_fill_kernel")
        traffic[cls] = {"dram_bytes_per_point": b / pts_per_launch, "launch_us": t * 1e6}
    except Exception as e:
        md.append(f"- (traffic parse failed: {e})")
    md.append("")
open(out_md, "w").write("\n".join(md) + "\n")
# keyed by the kernel sources the capture was taken from: bench.py only quotes the traffic while the sources are unchanged
import datetime, hashlib, os
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_h = hashlib.sha256()
for _f in ("fft_core.cuh", "fft_kernels.cuh"):
    _h.update(open(os.path.join(_root, "xrft_b200", "csrc", _f), "rb").read())
traffic = {"kernel_source_sha16": _h.hexdigest()[:16], "captured": datetime.date.today().isoformat(),
           "how": "ncu --set full --clock-control none, one launch per kernel (dram__bytes_read.sum + dram__bytes_write.sum)", "kernels": traffic}
json.dump(traffic, open(traffic_json, "w"), indent=1)
print("\n".join(md[:16])); print(json.dumps(traffic))
