#!/usr/bin/env python
"""power_spectrum along the LEADING axis of a [time][y][x] array (dim="time": the transform axis is moved last by xrftb_permute,
the reference's da.transpose, xrft.py:386-396) -- timing through the public API."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xrft_b200 as xrft
warnings.simplefilter("ignore")
for T, ny, nx in ((1024, 256, 256), (4096, 128, 128), (1000, 180, 360)):
    x = torch.randn((T, ny, nx), device="cuda", dtype=torch.float32)
    da = xrft.DataArray(x, dims=["time", "y", "x"], coords={"time": np.arange(T) * 1.0, "y": np.arange(ny) * 1.0, "x": np.arange(nx) * 1.0})
    f = lambda: xrft.power_spectrum(da, dim="time", detrend="linear", window="hann")
    for _ in range(2): out = f(); _ = out.data
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = f(); _ = out.data
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ref = np.abs(np.fft.fftshift(np.fft.fft(__import__("scipy.signal").signal.detrend(x[:, 0, :4].double().cpu().numpy(), axis=0) * __import__("scipy.signal").signal.windows.hann(T, sym=False)[:, None], axis=0), axes=0)) ** 2 / T
    got = np.asarray(out.values)[:, 0, :4] if out.dims[0].startswith("freq") else np.moveaxis(np.asarray(out.values), out.dims.index("freq_time"), 0)[:, 0, :4]
    print(f"power_spectrum(dim='time') {T} x {ny} x {nx} float32: {ms:.3f} ms = {T * ny * nx / ms / 1e6:.1f} GPoints/s; relerr vs numpy {np.linalg.norm(got - ref) / np.linalg.norm(ref):.1e}", flush=True)
