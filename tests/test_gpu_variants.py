"""Every kernel-variant knob of the fused power-spectrum chain must give the reference's numbers: each variant runs the chain in
a fresh process (the knobs are read once per process) and is compared with a float64 numpy evaluation of the reference formulas
(xrft/detrend.py:100-113 plane fit, xrft/xrft.py:96-103 window, :439-447 fft2 + fftshift, :740-748 |F|^2)."""
import os, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, numpy as np, torch, scipy.signal as sps
sys.path.insert(0, %(root)r)
from xrft_b200 import backend as B, _lib as L
worst = 0.0
for (ny, nx, T, detrend) in [(1024, 1024, 3, 2), (512, 2048, 2, 2), (2048, 1024, 2, 1), (4096, 4096, 1, 2), (64, 128, 5, 2)]:
    g = torch.Generator(device="cuda").manual_seed(7 + ny)
    x = torch.randn((T, ny, nx), generator=g, device="cuda", dtype=torch.float32)
    x += 0.3 * torch.arange(nx, device="cuda") - 0.7 * torch.arange(ny, device="cuda")[:, None] + 5
    wy = torch.from_numpy(sps.windows.hann(ny, sym=False)); wx = torch.from_numpy(sps.windows.hann(nx, sym=False))
    out = B.spectrum2d(x, None, L.EPI_POWER, detrend=detrend, win_y=wy, win_x=wx, shift_y=True, shift_x=True, scale=1.0 / (ny * nx))
    for t in (0, T - 1):
        xd = x[t].double().cpu().numpy()
        ii = np.arange(ny)[:, None] - 0.5 * (ny - 1); jj = np.arange(nx)[None, :] - 0.5 * (nx - 1)
        pl = xd.mean() + (0 if detrend == 1 else ii * ((ii * xd).sum() / (nx * ny * (ny * ny - 1) / 12)) + jj * ((jj * xd).sum() / (ny * nx * (nx * nx - 1) / 12)))
        d = (xd - pl).astype(np.float32).astype(np.float64)
        ref = np.fft.fftshift(np.abs(np.fft.fft2(d * wy.numpy()[:, None] * wx.numpy()[None, :])) ** 2) / (ny * nx)
        got = out[t].double().cpu().numpy()
        worst = max(worst, np.linalg.norm(got - ref) / np.linalg.norm(ref))
        # Hermitian symmetry of the spectrum of real data: mirrored cells are copies (exact) except where a chain computes
        # both members of a pair independently (self-mirrored row / column of the rows-first chain): rounding only
        assert np.abs(got[1:, 1:] - got[1:, 1:][::-1, ::-1]).max() <= 2e-6 * ref.max(), "Hermitian symmetry of the power spectrum"
print("WORST", worst)
assert worst < 1e-3   # north_star: 1e-3 relative for float32
'''

VARIANTS = [
    {},
    {"XRFTB_ZPACK": "0"},
    {"XRFTB_ZTMA": "0"},
    {"XRFTB_F32X2": "1"},
    {"XRFTB_COLS_ASYNC": "0"},
    {"XRFTB_ROWLINE": "0"},
    {"XRFTB_COLS_FIRST": "0"},
    {"XRFTB_COLS_FIRST": "0", "XRFTB_ROWLINE": "0"},
]


@pytest.mark.gpu
@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join(f"{k[6:]}={v}" for k, v in e.items()) or "default")
def test_power_chain_variant(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    worst = float(r.stdout.strip().split("WORST")[-1])
    assert worst < 2e-6, worst   # measured: 3e-7 (fp32 FFT with fp64-derived tables)


# ---- experimental chains written at the end of round 1 without GPU time left: opt in with XRFTB_TEST_EXPERIMENTAL=1
EXPERIMENTAL = os.environ.get("XRFTB_TEST_EXPERIMENTAL", "0") == "1"

CROSS_SCRIPT = r'''
import sys, numpy as np, torch, scipy.signal as sps
sys.path.insert(0, %(root)r)
from xrft_b200 import backend as B, _lib as L
worst = 0.0
for (ny, nx, T, detrend) in [(1024, 1024, 3, 1), (512, 2048, 2, 2), (2048, 4096, 1, 0)]:
    g = torch.Generator(device="cuda").manual_seed(11 + ny)
    x1 = torch.randn((T, ny, nx), generator=g, device="cuda") + 0.01 * torch.arange(nx, device="cuda") + 2
    x2 = torch.roll(x1, shifts=(3, 5), dims=(1, 2)) + 0.5 * torch.randn((T, ny, nx), generator=g, device="cuda") - 1
    wy = torch.from_numpy(sps.windows.hann(ny, sym=False)); wx = torch.from_numpy(sps.windows.hann(nx, sym=False))
    w = wy.numpy()[:, None] * wx.numpy()[None, :]
    def prep(x):
        xd = x.double().cpu().numpy()
        ii = np.arange(ny)[:, None] - 0.5 * (ny - 1); jj = np.arange(nx)[None, :] - 0.5 * (nx - 1)
        pl = 0.0 if detrend == 0 else xd.mean() + (0 if detrend == 1 else ii * ((ii * xd).sum() / (nx * ny * (ny * ny - 1) / 12)) + jj * ((jj * xd).sum() / (ny * nx * (nx * nx - 1) / 12)))
        return np.fft.fftshift(np.fft.fft2((xd - pl).astype(np.float32).astype(np.float64) * w))
    cs = B.spectrum2d(x1, x2, L.EPI_CROSS, detrend=detrend, win_y=wy, win_x=wx, shift_y=True, shift_x=True, scale=1.0 / (ny * nx))
    ph = B.spectrum2d(x1, x2, L.EPI_PHASE, detrend=detrend, win_y=wy, win_x=wx, shift_y=True, shift_x=True, scale=1.0 / (ny * nx))
    assert L.load().xrftb_spectrum2d_last_path() == 3, "the experimental chain did not run"
    for t in (0, T - 1):
        ref = prep(x1[t]) * np.conj(prep(x2[t])) / (ny * nx)
        worst = max(worst, np.linalg.norm(cs[t].cpu().numpy() - ref) / np.linalg.norm(ref))
        big = np.abs(ref) > 1e-5 * np.abs(ref).max()
        assert np.abs(np.angle(np.exp(1j * (ph[t].cpu().numpy() - np.angle(ref)))))[big].max() < 1e-2
print("WORST", worst)
'''


@pytest.mark.gpu
@pytest.mark.skipif(not EXPERIMENTAL, reason="experimental two-field z-mode chain: not yet validated on hardware (XRFTB_TEST_EXPERIMENTAL=1)")
def test_experimental_cross_z_chain():
    e = dict(os.environ)
    e["XRFTB_CROSS_Z"] = "1"
    r = subprocess.run([sys.executable, "-c", CROSS_SCRIPT % {"root": ROOT}], env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    assert float(r.stdout.strip().split("WORST")[-1]) < 1e-5
