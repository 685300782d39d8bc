#!/usr/bin/env python
"""Host<->device link probe beside the end-to-end number: pinned H2D alone, D2H alone and both directions concurrently
(two streams), CUDA-event timed; then the public power_spectrum call on pinned host buffers for several stream-chunk
sizes.  Usage: python tools/pcie_probe.py [slices]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

nbytes = 1 << 30
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


res = {"h2d_GBs": nbytes / timed(h2d) / 1e9, "d2h_GBs": nbytes / timed(d2h) / 1e9}
t = timed(both)
res["both_each_GBs"] = nbytes / t / 1e9
print(json.dumps(res), flush=True)
del h_in, h_out, d_a, d_b

import warnings
warnings.simplefilter("ignore")
import xrft_b200 as xrft
from xrft_b200 import api as A

Te = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ny = nx = 4096
hin = torch.randn((Te, ny, nx), dtype=torch.float32).pin_memory()
hout = torch.empty((Te, ny, nx), dtype=torch.float32).pin_memory()
hda = xrft.DataArray(hin.numpy(), dims=["time", "y", "x"], coords={"time": np.arange(Te) * 1.0, "y": np.arange(ny) * 1.0, "x": np.arange(nx) * 1.0})
for chunk_mb in (128, 256, 512, 1024):
    A._STREAM_CHUNK_BYTES = chunk_mb << 20

    def step():
        xrft.power_spectrum(hda, dim=["y", "x"], detrend="linear", window="hann", out=hout)
        torch.cuda.synchronize()

    dt = timed(step, reps=3)
    print(json.dumps({"stream_chunk_MiB": chunk_mb, "e2e_GPts_s": Te * ny * nx / dt / 1e9, "each_way_GBs": Te * ny * nx * 4 / dt / 1e9}), flush=True)
