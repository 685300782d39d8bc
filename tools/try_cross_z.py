#!/usr/bin/env python
"""Round-2 opener for the EXPERIMENTAL two-field z-mode chain (XRFTB_CROSS_Z=1; written blind at the end of round 1): runs
cross spectrum and cross phase of two 2048^2 fields (config 3's shapes) with the knob off and on, each in a fresh process,
checks both against a float64 numpy evaluation of the reference formulas and prints the timings.
Usage: python tools/try_cross_z.py [n] [T]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = r'''
import sys, numpy as np, torch, scipy.signal as sps
sys.path.insert(0, %(root)r)
from xrft_b200 import backend as B, _lib as L
n, T = %(n)d, %(T)d
g = torch.Generator(device="cuda").manual_seed(3)
x1 = torch.randn((T, n, n), generator=g, device="cuda") + 2.0
x2 = torch.roll(x1, shifts=(3, 5), dims=(1, 2)) + 0.5 * torch.randn((T, n, n), generator=g, device="cuda") - 1.0
wy = torch.from_numpy(sps.windows.hann(n, sym=False)); wx = wy.clone()
lib = L.load()
for mode, name in ((L.EPI_CROSS, "cross"), (L.EPI_PHASE, "phase")):
    f = lambda: B.spectrum2d(x1, x2, mode, detrend=1, win_y=wy, win_x=wx, shift_y=True, shift_x=True, scale=1.0 / (n * n))
    for _ in range(4):
        out = None
        out = f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = None
        out = f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    w = wy.numpy()[:, None] * wx.numpy()[None, :]
    a = x1[0].double().cpu().numpy(); b = x2[0].double().cpu().numpy()
    fa = np.fft.fftshift(np.fft.fft2(((a - a.mean()).astype(np.float32)).astype(np.float64) * w))
    fb = np.fft.fftshift(np.fft.fft2(((b - b.mean()).astype(np.float32)).astype(np.float64) * w))
    ref = fa * np.conj(fb) / (n * n)
    got = out[0].cpu().numpy()
    if name == "cross":
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    else:
        big = np.abs(ref) > 1e-6 * np.abs(ref).max()
        err = np.abs(np.angle(np.exp(1j * (got - np.angle(ref)))))[big].max()
    print(f"{name}: path={lib.xrftb_spectrum2d_last_path()} err={err:.2e} {ms:.3f} ms  {T * n * n / ms / 1e6:.1f} GPts/s per field")
'''
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
for knob in ("0", "1"):
    e = dict(os.environ); e["XRFTB_CROSS_Z"] = knob
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "n": n, "T": T}], env=e, capture_output=True, text=True)
    print(f"== XRFTB_CROSS_Z={knob}\n" + (r.stdout.strip() or "") + ("\n" + r.stderr.strip()[-1500:] if r.returncode else ""), flush=True)
