"""Small-shape run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, warnings; sys.path.insert(0, ".")
import numpy as np, torch
import xrft_b200 as xrft
from xrft_b200 import backend as B
warnings.simplefilter("ignore")
rng = np.random.default_rng(0)
def da(shape, dims, dt=np.float32):
    return xrft.DataArray(rng.standard_normal(shape).astype(dt), dims=dims, coords={d: np.arange(n) * 1.0 for d, n in zip(dims, shape)})
for dt in (np.float32, np.float64):
    a, b = da((3, 64, 128), ["t", "y", "x"], dt), da((3, 64, 128), ["t", "y", "x"], dt)
    xrft.power_spectrum(a, dim=["y", "x"], detrend="linear", window="hann").values
    xrft.power_spectrum(a, dim=["y", "x"], real_dim="x", detrend="constant").values
    xrft.cross_spectrum(a, b, dim=["y", "x"], detrend="constant", window="hann").values
    xrft.cross_phase(a, b, dim=["y", "x"]).values
    xrft.isotropic_power_spectrum(a, dim=["y", "x"], window="hann").values
    xrft.isotropic_cross_spectrum(a, b, dim=["y", "x"]).values
    f = xrft.fft(a, dim=["y", "x"], real_dim="x"); xrft.ifft(f, dim=["freq_y", "freq_x"], real_dim="freq_x").values
    xrft.fft(da((4, 20, 30), ["t", "y", "x"], dt), dim=["y", "x"], detrend="linear").values
    xrft.fft(da((2, 300), ["t", "x"], dt), dim=["x"]).values
    xrft.fft(da((5, 16, 8, 32), ["t", "z", "y", "x"], dt), dim=["z", "y", "x"], detrend="linear", window="hann").values
    xrft.detrend(a, ["y", "x"], "linear").values
    x = torch.from_numpy((rng.standard_normal((2, 1 << 15)) + 1j * rng.standard_normal((2, 1 << 15))).astype(np.complex64 if dt == np.float32 else np.complex128)).cuda()
    B.fftn(x, axes=[1]); B.fftn(x.reshape(2, 128, 256), axes=[1])
    xrft.power_spectrum(da((2, 512, 1024), ["t", "y", "x"], dt), dim=["y", "x"], detrend="linear", window="hann").values
torch.cuda.synchronize(); print("sanitize run ok")
