#!/usr/bin/env python
"""Config 4 (isotropic_power_spectrum of 512^2 float32 planes) sweep: points in flight per kernel chain (L2-resident
intermediate vs HBM-resident) for the register-LUT radial-bin kernel and the generic epilogue.
Usage: python tools/probe_c4.py [planes]"""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = r'''
import sys, json, warnings, numpy as np, torch
sys.path.insert(0, %(root)r)
warnings.simplefilter("ignore")
import xrft_b200 as xrft
from xrft_b200 import backend as B
P, n = %(planes)d, 512
x = torch.randn((P, n, n), device="cuda") + 0.25
da = xrft.DataArray(x, dims=["t", "y", "x"], coords={"t": np.arange(P) * 1.0, "y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0})
ref = None
for pts in (None, 16 << 18, 32 << 18, 64 << 18, 128 << 18, 256 << 18, 1024 << 18):
    B._BINS_CHUNK_POINTS = pts
    f = lambda: xrft.isotropic_power_spectrum(da, dim=["y", "x"], detrend="constant", window="hann").data
    for _ in range(2): out = f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    if ref is None: ref = out.double()
    err = float((out.double() - ref).norm() / ref.norm())
    print(json.dumps({"planes_per_chain": (pts // (n * n)) if pts else "default", "ms": ms, "GPts_s": P * n * n / ms / 1e6, "vs_first": err}), flush=True)
'''
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for knob in ("1", "0"):
    e = dict(os.environ); e["XRFTB_BINS_STATIC"] = knob
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "planes": planes}], env=e, capture_output=True, text=True)
    print(f"== XRFTB_BINS_STATIC={knob}\n" + r.stdout.strip() + ("\n" + r.stderr.strip()[-1500:] if r.returncode else ""), flush=True)
