#!/usr/bin/env python
"""A/B of the kernel variants selected by environment knobs (each read once per process): runs tools/probe_c2.py in one
subprocess per variant and prints its per-kernel times.  Usage: python tools/ab_variants.py [ny] [T] [reps]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ny = sys.argv[1] if len(sys.argv) > 1 else "4096"
T = sys.argv[2] if len(sys.argv) > 2 else "64"
reps = sys.argv[3] if len(sys.argv) > 3 else "3"
VARIANTS = [
    ("default: columns first, z mode, TMA in / TMA out", {}),
    ("z mode, Z stored from registers", {"XRFTB_ZTMA": "0"}),
    ("columns first, separated half spectrum (no z mode)", {"XRFTB_ZPACK": "0"}),
    ("z mode, packed FP32x2 butterflies", {"XRFTB_F32X2": "1"}),
    ("columns first, LDG pass 1", {"XRFTB_ZPACK": "0", "XRFTB_COLS_ASYNC": "0"}),
    ("columns first, moments pass instead of column lines", {"XRFTB_ROWLINE": "0"}),
    ("rows first + mirror pass", {"XRFTB_COLS_FIRST": "0"}),
]
for name, env in VARIANTS:
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "probe_c2.py"), ny, T, reps, "f32", "32"], env=e, capture_output=True, text=True, cwd=ROOT)
    print(f"[{name}] " + (r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
