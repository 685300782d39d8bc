#!/usr/bin/env python
"""Timing of complex transforms along non-power-of-two axes (mixed-radix kernel of smooth.cu; Bluestein for the rest) against
numpy's pocketfft on the host.  Usage: python tools/probe_smooth.py   (XRFTB_LIB selects another build of the library)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xrft_b200 import backend as B

CASES = [  # (shape, axes, dtype)
    ((64, 360, 720), (1, 2), np.complex64), ((64, 360, 720), (1, 2), np.complex128),
    ((8192, 3600), (1,), np.complex64), ((8192, 1000), (1,), np.complex64), ((4096, 1000), (1,), np.complex128),
    ((64, 1000, 512), (1,), np.complex64), ((16, 5000, 256), (1,), np.complex64),
    ((8192, 3650), (1,), np.complex64), ((8192, 1009), (1,), np.complex64),   # 3650 = 2 5^2 73 and a prime: Bluestein
]
for shape, axes, dt in CASES:
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)
    t = torch.from_numpy(x).cuda()
    for _ in range(2): y = B.fftn(t, axes=list(axes))
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps): y = B.fftn(t, axes=list(axes))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    xs = x[: max(1, shape[0] // 16)]
    t0 = time.perf_counter(); ref = np.fft.fftn(xs, axes=axes); cpu_ms = (time.perf_counter() - t0) * 1e3 * (shape[0] / xs.shape[0])
    err = np.linalg.norm(y[: xs.shape[0]].cpu().numpy() - ref) / np.linalg.norm(ref)
    npts = np.prod(shape)
    print(f"{shape} axes={axes} {np.dtype(dt).name}: {ms:8.3f} ms  {npts / ms / 1e6:7.2f} GPoints/s  {2 * npts * np.dtype(dt).itemsize / ms / 1e6:7.1f} GB/s (1R+1W)  relerr={err:.1e}  numpy (1 thread, extrapolated) {cpu_ms:8.1f} ms", flush=True)
