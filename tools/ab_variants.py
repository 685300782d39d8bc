#!/usr/bin/env python
"""A/B of the kernel chains selected by the options (XRFTB_<NAME> = the value an option starts with) and of alternative builds of the library (XRFTB_LIB): runs tools/probe_c2.py in one
subprocess per variant and prints its per-kernel times.  Usage: python tools/ab_variants.py [ny] [T] [reps]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ny = sys.argv[1] if len(sys.argv) > 1 else "4096"
T = sys.argv[2] if len(sys.argv) > 2 else "64"
reps = sys.argv[3] if len(sys.argv) > 3 else "3"
BASE = os.path.join(ROOT, "variants", "base", "xrft_b200", "libxrftb200.so")      # previous commit (tools/README: built by hand)
VARIANTS = [
    ("default", {}),
    ("Z stored from registers (ztma=0)", {"XRFTB_ZTMA": "0"}),
    ("separated half spectrum (no z mode)", {"XRFTB_ZPACK": "0"}),
    ("rows first + mirror pass", {"XRFTB_COLS_FIRST": "0"}),
]
if os.path.exists(BASE):
    VARIANTS += [("previous commit", {"XRFTB_LIB": BASE})]
for name, env in VARIANTS:
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "probe_c2.py"), ny, T, reps, "f32", "32"], env=e, capture_output=True, text=True, cwd=ROOT)
    print(f"[{name}] " + (r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
