#!/usr/bin/env python
"""Short transform lengths (<= 64) with large batches: direct DFT (one thread per sequence) against the mixed-radix kernel
(XRFTB_LIB pointing at a build with -DXRFTB_SMOOTH_SMALL=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xrft_b200 import backend as B
for shape, axes, dt, real in [((400000, 40), (1,), np.complex64, False), ((200000, 60), (1,), np.complex64, False), ((100000, 48), (1,), np.complex128, False),
                              ((2000, 30, 512), (1,), np.complex64, False), ((400000, 40), (1,), np.float32, True), ((4096, 36, 72), (1, 2), np.float32, True),
                              ((65536, 12, 20), (1, 2), np.float32, True), ((200000, 15), (1,), np.complex64, False)]:
    rng = np.random.default_rng(1)
    x = rng.standard_normal(shape).astype(dt) if real else (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)
    t = torch.from_numpy(x).cuda()
    f = (lambda: B.rfftn(t, axes=list(axes))) if real else (lambda: B.fftn(t, axes=list(axes)))
    for _ in range(2): y = f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): y = f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    sub = x[:64]
    ref = np.fft.rfftn(sub.astype(np.float64), axes=axes) if real else np.fft.fftn(sub.astype(np.complex128), axes=axes)
    err = np.linalg.norm(y[:64].cpu().numpy() - ref) / np.linalg.norm(ref)
    print(f"{shape} axes={axes} {np.dtype(dt).name}: {ms:.3f} ms  {np.prod(shape) / ms / 1e6:.1f} GPoints/s  relerr {err:.1e}", flush=True)
