#!/usr/bin/env python
"""Public-API timing and parity on a non-power-of-two grid (a 0.5-degree globe, 360 x 720): power_spectrum and
isotropic_power_spectrum with linear detrend + hann through the composed chain (detrend_window -> mixed-radix rfftn ->
spectral_post), against the oracle on a sub-batch and numpy's pocketfft timing on the host."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xrft_b200 as xrft
from oracle import xrft_oracle as O
warnings.simplefilter("ignore")
T, ny, nx = 256, 360, 720
g = torch.Generator(device="cuda").manual_seed(3)
x = torch.randn((T, ny, nx), generator=g, device="cuda", dtype=torch.float32) + 0.01 * torch.arange(nx, device="cuda")
c = {"t": np.arange(T) * 1.0, "y": np.arange(ny) * 0.5, "x": np.arange(nx) * 0.5}
da = xrft.DataArray(x, dims=["t", "y", "x"], coords=c)
for name, f in (("power_spectrum", lambda d: xrft.power_spectrum(d, dim=["y", "x"], detrend="linear", window="hann")),
                ("isotropic_power_spectrum", lambda d: xrft.isotropic_power_spectrum(d, dim=["y", "x"], detrend="linear", window="hann"))):
    for _ in range(2): out = f(da); _ = out.data
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = f(da); _ = out.data
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    sub = O.Labelled(x[:2].cpu().numpy(), ("t", "y", "x"), {"t": c["t"][:2], "y": c["y"], "x": c["x"]})
    t0 = time.perf_counter()
    ref = getattr(O, name)(sub, dim=["y", "x"], detrend="linear", window="hann").data
    cpu_ms = (time.perf_counter() - t0) * 1e3 * T / 2
    got = np.asarray(out.values)[:2]
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print(f"{name} {T} x {ny} x {nx} float32: {ms:.3f} ms = {T * ny * nx / ms / 1e6:.1f} GPoints/s; relerr vs oracle {err:.1e}; oracle on one host thread (extrapolated) {cpu_ms:.0f} ms", flush=True)
