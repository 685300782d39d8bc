"""Every chain-selection option of the C-ABI (xrftb_set_option, include/xrft_b200.h) must give the reference's numbers: each
setting runs the fused chains and is compared with a float64 numpy evaluation of the reference formulas
(xrft/detrend.py:100-113 plane fit, xrft/xrft.py:96-103 window, :439-447 fft2 + fftshift, :740-748 |F|^2, :825-833 cross
spectrum, :895-906 radial bins)."""
import warnings

import numpy as np
import pytest
import scipy.signal as sps

pytestmark = pytest.mark.gpu

DEFAULTS = {"cols_first": 1, "zpack": 1, "ztma": 1, "cols_async": 2, "rowline": 1, "cross_z": 1, "bins_static": 1}
VARIANTS = [
    {},
    {"zpack": 0},
    {"ztma": 0},
    {"cols_async": 0},
    {"cols_async": 1},
    {"rowline": 0},
    {"cols_first": 0},
    {"cols_first": 0, "rowline": 0},
]


@pytest.fixture
def options():
    from xrft_b200 import _lib as L

    def apply(over):
        for k, v in {**DEFAULTS, **over}.items():
            L.set_option(k, v)

    yield apply
    apply({})


def plane_detrend(xd, detrend):
    ny, nx = xd.shape
    if detrend == 0:
        return xd
    ii = np.arange(ny)[:, None] - 0.5 * (ny - 1); jj = np.arange(nx)[None, :] - 0.5 * (nx - 1)
    pl = xd.mean() + (0 if detrend == 1 else ii * ((ii * xd).sum() / (nx * ny * (ny * ny - 1) / 12)) + jj * ((jj * xd).sum() / (ny * nx * (nx * nx - 1) / 12)))
    return xd - pl


def test_options_api():
    from xrft_b200 import _lib as L
    assert L.get_option("zpack") in (0, 1)
    L.set_option("zpack", 0)
    assert L.get_option("zpack") == 0
    L.set_option("zpack", 1)
    with pytest.raises(L.XrftbError):
        L.set_option("no_such_option", 1)


@pytest.mark.parametrize("over", VARIANTS, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()) or "default")
def test_power_chain_variant(over, options):
    import torch
    from xrft_b200 import backend as B, _lib as L
    options(over)
    worst = 0.0
    for (ny, nx, T, detrend) in [(1024, 1024, 3, 2), (512, 2048, 2, 2), (2048, 1024, 2, 1), (4096, 4096, 1, 2), (64, 128, 5, 2), (512, 512, 3, 1)]:
        g = torch.Generator(device="cuda").manual_seed(7 + ny)
        x = torch.randn((T, ny, nx), generator=g, device="cuda", dtype=torch.float32)
        x += 0.3 * torch.arange(nx, device="cuda") - 0.7 * torch.arange(ny, device="cuda")[:, None] + 5
        wy = torch.from_numpy(sps.windows.hann(ny, sym=False)); wx = torch.from_numpy(sps.windows.hann(nx, sym=False))
        out = B.spectrum2d(x, None, L.EPI_POWER, detrend=detrend, win_y=wy, win_x=wx, shift_y=True, shift_x=True, scale=1.0 / (ny * nx))
        for t in (0, T - 1):
            d = plane_detrend(x[t].double().cpu().numpy(), detrend).astype(np.float32).astype(np.float64)
            ref = np.fft.fftshift(np.abs(np.fft.fft2(d * wy.numpy()[:, None] * wx.numpy()[None, :])) ** 2) / (ny * nx)
            got = out[t].double().cpu().numpy()
            worst = max(worst, np.linalg.norm(got - ref) / np.linalg.norm(ref))
            # Hermitian symmetry of the spectrum of real data: mirrored cells are copies (exact) except where a chain computes
            # both members of a pair independently (self-mirrored row / column of the rows-first chain): rounding only
            assert np.abs(got[1:, 1:] - got[1:, 1:][::-1, ::-1]).max() <= 2e-6 * ref.max(), "Hermitian symmetry of the power spectrum"
    assert worst < 2e-6, worst   # north_star: 1e-3 for float32; measured 3e-7 (fp32 FFT with fp64-derived tables)


@pytest.mark.parametrize("over", [{}, {"cross_z": 0}], ids=["cross_z=1", "cross_z=0"])
def test_cross_chain_variant(over, options):
    """cross spectrum, cross phase and both-in-one-pass on the two-field z-mode chain and on the rows-first two-field chain"""
    import torch
    from xrft_b200 import backend as B, _lib as L
    options(over)
    lib = L.load()
    worst = 0.0
    for (ny, nx, T, detrend) in [(1024, 1024, 3, 1), (512, 2048, 2, 2), (2048, 4096, 1, 0)]:
        g = torch.Generator(device="cuda").manual_seed(11 + ny)
        x1 = torch.randn((T, ny, nx), generator=g, device="cuda") + 0.01 * torch.arange(nx, device="cuda") + 2
        x2 = torch.roll(x1, shifts=(3, 5), dims=(1, 2)) + 0.5 * torch.randn((T, ny, nx), generator=g, device="cuda") - 1
        wy = torch.from_numpy(sps.windows.hann(ny, sym=False)); wx = torch.from_numpy(sps.windows.hann(nx, sym=False))
        w = wy.numpy()[:, None] * wx.numpy()[None, :]
        prep = lambda x: np.fft.fftshift(np.fft.fft2(plane_detrend(x.double().cpu().numpy(), detrend).astype(np.float32).astype(np.float64) * w))
        kw = dict(detrend=detrend, win_y=wy, win_x=wx, shift_y=True, shift_x=True, scale=1.0 / (ny * nx))
        cs = B.spectrum2d(x1, x2, L.EPI_CROSS, **kw)
        assert lib.xrftb_spectrum2d_last_path() == (3 if over.get("cross_z", 1) else 0)
        ph = B.spectrum2d(x1, x2, L.EPI_PHASE, **kw)
        cs2, ph2 = B.spectrum2d(x1, x2, L.EPI_CROSS, with_phase=True, **kw)
        assert torch.equal(cs2, cs)
        dphi = (ph2 - ph).abs()
        assert float(torch.minimum(dphi, 2 * np.pi - dphi).max()) < 1e-5
        for t in (0, T - 1):
            ref = prep(x1[t]) * np.conj(prep(x2[t])) / (ny * nx)
            worst = max(worst, np.linalg.norm(cs[t].cpu().numpy() - ref) / np.linalg.norm(ref))
            big = np.abs(ref) > 1e-5 * np.abs(ref).max()
            assert np.abs(np.angle(np.exp(1j * (ph[t].cpu().numpy() - np.angle(ref)))))[big].max() < 1e-2
    assert worst < 1e-5, worst


@pytest.mark.parametrize("over", [{}, {"bins_static": 0}, {"cols_first": 0}, {"cols_first": 0, "bins_static": 0},
                                  {"cols_first": 0, "bins_static": 0, "cols_async": 1}],
                         ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()) or "default")
def test_bins_chain_variant(over, options):
    """isotropic power spectrum: radial bins in pass 2 of the columns-first chain (default), the register-LUT column kernel of
    the rows-first chain and the generic LUT epilogue agree with numpy's bincount of the reference's codes"""
    import torch
    import xrft_b200 as xrft
    from xrft_b200 import api as A
    options(over)
    warnings.simplefilter("ignore")
    rng = np.random.default_rng(5)
    for (T, ny, nx, detrend) in [(5, 512, 512, "constant"), (3, 256, 1024, "linear"), (2, 1024, 256, None), (3, 64, 64, "constant"),
                                 (2, 256, 4096, "constant"), (1, 128, 8192, "linear"), (3, 256, 64, "constant"), (2, 512, 128, None)]:
        x = (rng.standard_normal((T, ny, nx)) + 0.5).astype(np.float32)
        c = {"t": np.arange(T) * 1.0, "y": np.arange(ny) * 1.0, "x": np.arange(nx) * 1.0}
        iso = xrft.isotropic_power_spectrum(xrft.DataArray(torch.from_numpy(x).cuda(), dims=["t", "y", "x"], coords=c), dim=["y", "x"],
                                            detrend=detrend, window="hann").values.astype(np.float64)
        det = {None: 0, "constant": 1, "linear": 2}[detrend]
        w = sps.windows.hann(ny, sym=False)[:, None] * sps.windows.hann(nx, sym=False)[None, :]
        k = np.fft.fftshift(np.fft.fftfreq(nx, 1.0)); l = np.fft.fftshift(np.fft.fftfreq(ny, 1.0))
        codes, nbins, _ = A._radial_bins(k, l, 4, False)      # codes[i_k, i_l]
        for t in range(T):
            d = plane_detrend(x[t].astype(np.float64), det).astype(np.float32).astype(np.float64)
            ps = np.fft.fftshift(np.abs(np.fft.fft2(d * w)) ** 2) / (ny * nx)
            ref = np.bincount(codes.T.ravel(), weights=ps.ravel(), minlength=nbins)
            assert np.linalg.norm(iso[t] - ref) / np.linalg.norm(ref) < 1e-5


@pytest.mark.parametrize("over", [{}, {"bins_static": 0}, {"cross_z": 0}], ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()) or "default")
def test_bins_cross_chain_variant(over, options):
    """isotropic cross spectrum: radial bins in the two-field pass 2 of the z-mode chain (default) and the generic two-field
    LUT epilogue agree with numpy's bincount of the reference's codes (real part; the imaginary part cancels)"""
    import torch
    import xrft_b200 as xrft
    from xrft_b200 import api as A
    options(over)
    warnings.simplefilter("ignore")
    rng = np.random.default_rng(6)
    for (T, ny, nx, detrend) in [(3, 1024, 1024, "constant"), (2, 512, 2048, "linear"), (2, 2048, 1024, None), (3, 64, 128, "constant")]:
        x = (rng.standard_normal((T, ny, nx)) + 0.5).astype(np.float32)
        y = (np.roll(x, (2, 3), axis=(1, 2)) + 0.3 * rng.standard_normal((T, ny, nx))).astype(np.float32)
        c = {"t": np.arange(T) * 1.0, "y": np.arange(ny) * 1.0, "x": np.arange(nx) * 1.0}
        mk_ = lambda v: xrft.DataArray(torch.from_numpy(v).cuda(), dims=["t", "y", "x"], coords=c)
        iso = xrft.isotropic_cross_spectrum(mk_(x), mk_(y), dim=["y", "x"], detrend=detrend, window="hann").values
        det = {None: 0, "constant": 1, "linear": 2}[detrend]
        w = sps.windows.hann(ny, sym=False)[:, None] * sps.windows.hann(nx, sym=False)[None, :]
        k = np.fft.fftshift(np.fft.fftfreq(nx, 1.0)); l = np.fft.fftshift(np.fft.fftfreq(ny, 1.0))
        codes, nbins, _ = A._radial_bins(k, l, 4, False)
        for t in range(T):
            fa = np.fft.fftshift(np.fft.fft2(plane_detrend(x[t].astype(np.float64), det).astype(np.float32).astype(np.float64) * w))
            fb = np.fft.fftshift(np.fft.fft2(plane_detrend(y[t].astype(np.float64), det).astype(np.float32).astype(np.float64) * w))
            cs = fa * np.conj(fb) / (ny * nx)
            ref = np.bincount(codes.T.ravel(), weights=cs.real.ravel(), minlength=nbins)
            assert np.linalg.norm(iso[t].real - ref) / np.linalg.norm(ref) < 1e-5
            assert np.abs(iso[t].imag).max() < 1e-5 * np.abs(ref).max()
