#!/usr/bin/env python
"""A/B of the kernel variants selected by environment knobs (each read once per process): runs tools/probe_c2.py in one
subprocess per variant and prints its per-kernel times.  Usage: python tools/ab_variants.py [ny] [T] [reps]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ny = sys.argv[1] if len(sys.argv) > 1 else "4096"
T = sys.argv[2] if len(sys.argv) > 2 else "64"
reps = sys.argv[3] if len(sys.argv) > 3 else "3"
VARIANTS = [
    ("columns-first, moments pass, LDG pass 1", {"XRFTB_COLS_FIRST": "1", "XRFTB_ROWLINE": "0", "XRFTB_COLS_ASYNC": "0"}),
    ("columns-first, column-line, LDG pass 1", {"XRFTB_COLS_FIRST": "1", "XRFTB_ROWLINE": "1", "XRFTB_COLS_ASYNC": "0"}),
    ("columns-first, moments pass, TMA pass 1", {"XRFTB_COLS_FIRST": "1", "XRFTB_ROWLINE": "0"}),
    ("columns-first, column-line, TMA pass 1", {"XRFTB_COLS_FIRST": "1", "XRFTB_ROWLINE": "1"}),
    ("register-prefetch cols + row-line detrend", {"XRFTB_COLS_ASYNC": "0", "XRFTB_ROWLINE": "1"}),
]
for name, env in VARIANTS:
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "probe_c2.py"), ny, T, reps, "f32", "8"], env=e, capture_output=True, text=True, cwd=ROOT)
    print(f"[{name}] " + (r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
