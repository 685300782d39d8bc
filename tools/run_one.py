#!/usr/bin/env python
"""One small invocation of a BASELINE config's hot path (for ncu / compute-sanitizer; NOT a timing harness).
Usage: python tools/run_one.py {c2|c3|c3both|c4|c4power|c5} [batch] [reps]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xrft_b200 as xrft
warnings.simplefilter("ignore")
what = sys.argv[1] if len(sys.argv) > 1 else "c4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = "cuda"


def field(T, n, dt=torch.float32, seed=1):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn((T, n, n), generator=g, device=dev, dtype=dt) + 0.25
    c = {"t": np.arange(T) * 1.0, "y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0}
    return xrft.DataArray(x, dims=["t", "y", "x"], coords=c)


if what in ("c4", "c4power"):
    da = field(B or 1024, 512)
    f = (lambda: xrft.isotropic_power_spectrum(da, dim=["y", "x"], detrend="constant", window="hann")) if what == "c4" else \
        (lambda: xrft.power_spectrum(da, dim=["y", "x"], detrend="constant", window="hann"))
elif what in ("c3", "c3both"):
    a, b = field(B or 16, 2048, seed=1), field(B or 16, 2048, seed=2)
    kw = dict(dim=["y", "x"], detrend="constant", window="hann")
    if what == "c3":
        f = lambda: (xrft.cross_spectrum(a, b, **kw), xrft.cross_phase(a, b, **kw))
    else:
        f = lambda: xrft.cross_spectrum_and_phase(a, b, **kw)
elif what == "c2":
    da = field(B or 8, 4096)
    da.data.add_(0.3 * torch.arange(4096, device=dev) - 0.7 * torch.arange(4096, device=dev)[:, None] + 5)
    f = lambda: xrft.power_spectrum(da, dim=["y", "x"], detrend="linear", window="hann")
elif what == "c5":
    n, p = B or 8192, (B or 8192) // 2
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn((n, n), generator=g, device=dev, dtype=torch.float64)
    d0 = xrft.DataArray(x, dims=["y", "x"], coords={"y": np.arange(n) * 0.5, "x": np.arange(n) * 0.5})

    def f():
        padded = xrft.pad(d0, x=p, y=p)
        ft = xrft.fft(padded, real_dim="x")
        back = xrft.ifft(ft, real_dim="freq_x")
        un = xrft.unpad(back, {"x": p, "y": p})
        un.data
        return un
else:
    raise SystemExit("unknown workload " + what)
for _ in range(reps):
    out = f()
    del out
torch.cuda.synchronize()
print("done", what)
