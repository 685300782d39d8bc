"""GPU parity of the raw C-ABI kernels against numpy.fft (the reference's own backend, xrft.py:32-36)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {np.float32: 2e-5, np.float64: 1e-12}


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def cplx(rng, shape, dt):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64 if dt == np.float32 else np.complex128)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_c2c_rows(dt, n):
    from xrft_b200 import backend as B
    rng = np.random.default_rng(n)
    x = cplx(rng, (37, n), dt)
    y = B.fftn(torch.from_numpy(x).cuda(), axes=[1]).cpu().numpy()
    assert relerr(y, np.fft.fft(x.astype(np.complex128), axis=1)) < TOL[dt]
    yi = B.ifftn(torch.from_numpy(x).cuda(), axes=[1]).cpu().numpy()
    assert relerr(yi, np.fft.ifft(x.astype(np.complex128), axis=1)) < TOL[dt]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [2, 4, 8, 16, 64, 256, 512, 1024, 4096, 8192])
def test_c2c_cols(dt, n):
    from xrft_b200 import backend as B
    rng = np.random.default_rng(n + 1)
    x = cplx(rng, (3, n, 21), dt)
    y = B.fftn(torch.from_numpy(x).cuda(), axes=[1]).cpu().numpy()
    assert relerr(y, np.fft.fft(x.astype(np.complex128), axis=1)) < TOL[dt]
    yi = B.ifftn(torch.from_numpy(x).cuda(), axes=[1]).cpu().numpy()
    assert relerr(yi, np.fft.ifft(x.astype(np.complex128), axis=1)) < TOL[dt]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 256, 1024, 4096, 16384])
def test_r2c_c2r(dt, n):
    from xrft_b200 import backend as B
    rng = np.random.default_rng(n + 2)
    x = rng.standard_normal((5, 16, n)).astype(dt)
    y = B.rfftn(torch.from_numpy(x).cuda(), axes=[1, 2]).cpu().numpy()
    ref = np.fft.rfftn(x.astype(np.float64), axes=[1, 2])
    assert relerr(y, ref) < TOL[dt]
    back = B.irfftn(torch.from_numpy(ref.astype(y.dtype)).cuda(), axes=[1, 2]).cpu().numpy()
    assert relerr(back, x) < TOL[dt]
    y1 = B.rfftn(torch.from_numpy(x).cuda(), axes=[2]).cpu().numpy()
    assert relerr(y1, np.fft.rfft(x.astype(np.float64), axis=2)) < TOL[dt]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_fftn_3d(dt):
    from xrft_b200 import backend as B
    rng = np.random.default_rng(7)
    x = cplx(rng, (2, 32, 16, 64), dt)
    y = B.fftn(torch.from_numpy(x).cuda(), axes=[1, 2, 3]).cpu().numpy()
    assert relerr(y, np.fft.fftn(x.astype(np.complex128), axes=[1, 2, 3])) < TOL[dt]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(3, 16, 32), (2, 64, 64), (2, 512, 256), (1, 1024, 2048), (1, 8, 4), (5, 2, 16)])
@pytest.mark.parametrize("detrend", [0, 1, 2])
def test_spectrum2d_power(dt, shape, detrend):
    """fused detrend+hann+rfft2+|F|^2 (full, shifted) vs numpy."""
    import scipy.signal as sps
    from xrft_b200 import backend as B, _lib as L
    from oracle import xrft_oracle as O
    rng = np.random.default_rng(sum(shape) + detrend)
    b, ny, nx = shape
    x = (rng.standard_normal(shape) + 0.3 * np.arange(nx) - 0.7 * np.arange(ny)[:, None] + 5).astype(dt)
    la = O.Labelled(x, ("t", "y", "x"))
    d = {0: None, 1: "constant", 2: "linear"}[detrend]
    ref = O.power_spectrum(O.Labelled(x.astype(np.float64), ("t", "y", "x")), dim=["y", "x"], detrend=d, window="hann").data
    wy = torch.from_numpy(sps.windows.hann(ny, sym=False))
    wx = torch.from_numpy(sps.windows.hann(nx, sym=False))
    scale = 1.0 / (ny * nx)  # dx = dy = 1: (dx dy)^2 / (N dx N dy)
    out = B.spectrum2d(torch.from_numpy(x).cuda(), None, L.EPI_POWER, detrend=detrend, win_y=wy, win_x=wx,
                       shift_y=True, shift_x=True, scale=scale).cpu().numpy()
    tol = 5e-4 if dt == np.float32 else 1e-10
    assert relerr(out, ref) < tol
