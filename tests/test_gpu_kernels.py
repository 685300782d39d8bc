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


# ---------------------------------------------------------------- arbitrary lengths
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 3, 5, 10, 15, 19, 20, 30, 60, 100, 258, 366, 1000, 2601, 3650])
def test_c2c_any_length(dt, n):
    from xrft_b200 import backend as B
    rng = np.random.default_rng(n)
    x = cplx(rng, (3, n), dt)
    tol = TOL[dt] * (10 if n > 64 else 1)
    y = B.fftn(torch.from_numpy(x).cuda(), axes=[1]).cpu().numpy()
    assert relerr(y, np.fft.fft(x.astype(np.complex128), axis=1)) < tol
    yi = B.ifftn(torch.from_numpy(x).cuda(), axes=[1]).cpu().numpy()
    assert relerr(yi, np.fft.ifft(x.astype(np.complex128), axis=1)) < tol
    if n <= 1000:  # strided axis
        xs = cplx(rng, (2, n, 7), dt)
        ys = B.fftn(torch.from_numpy(xs).cuda(), axes=[1]).cpu().numpy()
        assert relerr(ys, np.fft.fft(xs.astype(np.complex128), axis=1)) < tol


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(4, 10), (3, 15, 20), (2, 30, 40), (2, 6, 100), (1, 12, 1000), (5, 2), (3, 3), (2, 8, 2)])
def test_r2c_c2r_any_length(dt, shape):
    from xrft_b200 import backend as B
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(dt)
    axes = list(range(1, len(shape)))
    y = B.rfftn(torch.from_numpy(x).cuda(), axes=axes).cpu().numpy()
    ref = np.fft.rfftn(x.astype(np.float64), axes=axes)
    tol = TOL[dt] * 10
    assert relerr(y, ref) < tol
    if shape[-1] % 2 == 0:
        back = B.irfftn(torch.from_numpy(ref.astype(y.dtype)).cuda(), axes=axes).cpu().numpy()
        assert relerr(back, x) < tol


def test_binned_sum_and_moments():
    from xrft_b200 import backend as B
    rng = np.random.default_rng(3)
    a = rng.standard_normal((4, 50, 60))
    lut = rng.integers(-1, 17, size=(50, 60)).astype(np.int32)
    out = B.binned_sum(torch.from_numpy(a).cuda(), torch.from_numpy(lut), 17, 2).cpu().numpy()
    ref = np.stack([np.bincount(lut.ravel()[lut.ravel() >= 0], a[i].ravel()[lut.ravel() >= 0], 17) for i in range(4)])
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)
    ac = (a + 1j * rng.standard_normal(a.shape)).astype(np.complex64)
    outc = B.binned_sum(torch.from_numpy(ac).cuda(), torch.from_numpy(lut), 17, 2).cpu().numpy()
    refc = np.stack([np.bincount(lut.ravel()[lut.ravel() >= 0], ac[i].real.ravel()[lut.ravel() >= 0], 17)
                     + 1j * np.bincount(lut.ravel()[lut.ravel() >= 0], ac[i].imag.ravel()[lut.ravel() >= 0], 17) for i in range(4)])
    np.testing.assert_allclose(outc, refc, rtol=1e-5, atol=1e-5)
    m = B.moments(torch.from_numpy(a).cuda(), 2).cpu().numpy()
    iy = np.arange(50) - 24.5; ix = np.arange(60) - 29.5
    np.testing.assert_allclose(m[:, 0], a.sum(axis=(1, 2)), rtol=1e-12)
    np.testing.assert_allclose(m[:, 2], (a * iy[None, :, None]).sum(axis=(1, 2)), rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(m[:, 3], (a * ix[None, None, :]).sum(axis=(1, 2)), rtol=1e-10, atol=1e-9)


# ------------------------------------------------- lengths beyond one CTA: four-step (+ Bluestein on top of it)
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1 << 15, 1 << 16, 1 << 18, 20000, 440000])
def test_long_contiguous(dt, n):
    from xrft_b200 import backend as B
    rng = np.random.default_rng(n % 1000)
    x = cplx(rng, (2, n), dt)
    tol = 5e-5 if dt == np.float32 else 1e-11
    y = B.fftn(torch.from_numpy(x).cuda(), axes=[1]).cpu().numpy()
    assert relerr(y, np.fft.fft(x.astype(np.complex128), axis=1)) < tol
    yi = B.ifftn(torch.from_numpy(x).cuda(), axes=[1]).cpu().numpy()
    assert relerr(yi, np.fft.ifft(x.astype(np.complex128), axis=1)) < tol
    xr = rng.standard_normal((2, n)).astype(dt)
    yr = B.rfftn(torch.from_numpy(xr).cuda(), axes=[1]).cpu().numpy()
    ref = np.fft.rfft(xr.astype(np.float64), axis=1)
    assert relerr(yr, ref) < tol
    if n % 2 == 0:
        back = B.irfftn(torch.from_numpy(ref.astype(yr.dtype)).cuda(), axes=[1]).cpu().numpy()
        assert relerr(back, xr) < tol


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_long_strided_16384(dt):
    """the strided axis of BASELINE config 5 at full size (8192^2 padded to 16384^2): four-step 128 x 128"""
    from xrft_b200 import backend as B
    rng = np.random.default_rng(5)
    x = cplx(rng, (16384, 9), dt)
    y = B.fftn(torch.from_numpy(x).cuda(), axes=[0]).cpu().numpy()
    assert relerr(y, np.fft.fft(x.astype(np.complex128), axis=0)) < (5e-5 if dt == np.float32 else 1e-11)
    xr = rng.standard_normal((16384, 64)).astype(dt)
    yr = B.rfftn(torch.from_numpy(xr).cuda(), axes=[0, 1]).cpu().numpy()
    assert relerr(yr, np.fft.rfftn(xr.astype(np.float64))) < (5e-5 if dt == np.float32 else 1e-11)


@pytest.mark.gpu
def test_fft_module_matches_numpy_fft_at_the_reference_call_sites():
    """backend.fft_module() driven exactly like xrft/xrft.py:398-404, 439-447 (forward) and :586-591, 612-621 (inverse)
    drives the object `_fft_module` returns: numpy in, numpy out, numpy.fft semantics."""
    import numpy as np
    from xrft_b200.backend import fft_module
    fftm = fft_module()
    rng = np.random.default_rng(3)
    for shape, axis_num in (((6, 32, 64), [1, 2]), ((20, 30), [0, 1]), ((5, 16, 8), [0]), ((4, 15, 19), [2, 1])):
        a = rng.standard_normal(shape)
        # forward: fftn(ifftshift(a, axes), axes) then fftshift   (xrft.py:439-447)
        f = fftm.fftshift(fftm.fftn(fftm.ifftshift(a, axes=axis_num), axes=axis_num), axes=axis_num)
        ref = np.fft.fftshift(np.fft.fftn(np.fft.ifftshift(a, axes=axis_num), axes=axis_num), axes=axis_num)
        assert isinstance(f, np.ndarray) and f.dtype == np.complex128
        np.testing.assert_allclose(f, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
        # real transform over the last listed axis, no shift on it   (xrft.py:398-404)
        fr = fftm.rfftn(a, axes=axis_num)
        np.testing.assert_allclose(fr, np.fft.rfftn(a, axes=axis_num), rtol=1e-9, atol=1e-9 * np.abs(ref).max())
        # inverse   (xrft.py:586-591, 612-621)
        b = fftm.ifftshift(fftm.ifftn(fftm.ifftshift(f, axes=axis_num), axes=axis_num), axes=axis_num)
        refb = np.fft.ifftshift(np.fft.ifftn(np.fft.ifftshift(ref, axes=axis_num), axes=axis_num), axes=axis_num)
        np.testing.assert_allclose(b, refb, rtol=1e-9, atol=1e-9)
        if shape[axis_num[-1]] % 2 == 0:
            br = fftm.irfftn(fr, axes=axis_num)
            np.testing.assert_allclose(br, np.fft.irfftn(np.fft.rfftn(a, axes=axis_num), axes=axis_num), rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(br, a, rtol=1e-9, atol=1e-9)
    a32 = rng.standard_normal((3, 64, 128)).astype(np.float32)
    f32 = fftm.fftn(a32, axes=[1, 2])
    assert f32.dtype == np.complex64
    np.testing.assert_allclose(f32, np.fft.fftn(a32.astype(np.float64), axes=[1, 2]), rtol=1e-4, atol=1e-4 * 128)
    with pytest.raises(NotImplementedError):
        fftm.fftn(a32, s=(8, 8))
    np.testing.assert_array_equal(fftm.fftfreq(8, 0.5), np.fft.fftfreq(8, 0.5))


@pytest.mark.gpu
@pytest.mark.parametrize("dt,ny,nx", [(np.float64, 64, 128), (np.float32, 256, 64), (np.float64, 16384, 32), (np.float32, 32768, 16), (np.float64, 8, 4096)])
def test_fft2r_fused_forward_and_inverse(dt, ny, nx):
    """xrftb_fft2r (pad predicate / ramps / rolls / crop folded into the passes of a 2-D real transform, single-pass and
    four-step strided axis) against numpy: pad -> rfft2 -> x ramps x scale, and x ramps -> roll -> irfft2 -> roll -> crop"""
    from xrft_b200 import backend as B
    rng = np.random.default_rng(ny + nx)
    tol = 1e-10 if dt == np.float64 else 2e-4
    cdt = np.complex128 if dt == np.float64 else np.complex64
    # ---- forward, padded input
    py, px = (ny // 4, ny // 8), (nx // 4, nx // 2)
    iny, inx = ny - sum(py), nx - sum(px)
    x = rng.standard_normal((3, iny, inx)).astype(dt)
    ry = np.exp(1j * rng.uniform(0, 6.28, ny)).astype(cdt); rx = np.exp(1j * rng.uniform(0, 6.28, nx // 2 + 1)).astype(cdt)
    got = B.fft2r_forward(torch.from_numpy(x).cuda(), (py, px), torch.from_numpy(ry), torch.from_numpy(rx), 0.37).cpu().numpy()
    xp = np.pad(x.astype(np.float64), ((0, 0), py, px))
    ref = np.fft.rfft2(xp) * ry[None, :, None] * rx[None, None, :] * 0.37
    assert got.shape == ref.shape and relerr(got, ref) < tol
    # ---- forward, no padding, no ramps
    got0 = B.fft2r_forward(torch.from_numpy(xp.astype(dt)).cuda(), None, None, None, 1.0).cpu().numpy()
    assert relerr(got0, np.fft.rfft2(xp.astype(dt).astype(np.float64))) < tol
    # ---- inverse with rolled rows, ramps, output rolls, scale and crop
    f = (rng.standard_normal((2, ny, nx // 2 + 1)) + 1j * rng.standard_normal((2, ny, nx // 2 + 1))).astype(cdt)
    roll_in, (sy, sx) = ny // 2 + 3, (5 % ny, (nx // 2 + 2) % nx)
    got = B.fft2r_inverse(torch.from_numpy(f).cuda(), roll_in, torch.from_numpy(ry), torch.from_numpy(rx), (sy, sx), 1.7).cpu().numpy()
    g = f.astype(np.complex128) * ry[None, :, None] * rx[None, None, :]
    g = np.roll(g, -roll_in, axis=1)                       # transform row r reads input row (r + roll_in) % ny
    full = np.roll(np.fft.irfft2(g, s=(ny, nx)), (sy, sx), axis=(1, 2)) * 1.7
    assert relerr(got, full) < tol
    crop = ((ny // 4, ny // 2), (2, nx - 6))
    gotc = B.fft2r_inverse(torch.from_numpy(f).cuda(), roll_in, torch.from_numpy(ry), torch.from_numpy(rx), (sy, sx), 1.7, crop=crop).cpu().numpy()
    assert relerr(gotc, full[:, crop[0][0]:crop[0][0] + crop[0][1], crop[1][0]:crop[1][0] + crop[1][1]]) < tol
    # odd roll / odd crop offsets take the scalar store path
    goto = B.fft2r_inverse(torch.from_numpy(f).cuda(), 0, None, None, (0, 3), 1.0, crop=((1, ny - 2), (1, nx - 3))).cpu().numpy()
    fullo = np.roll(np.fft.irfft2(f.astype(np.complex128), s=(ny, nx)), 3, axis=2)
    assert relerr(goto, fullo[:, 1:ny - 1, 1:nx - 2]) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_chirpz_strided_long_and_inverse(dt):
    """Bluestein on a strided axis whose padded length needs the four-step decomposition (hooks travel through both
    decompositions), forward and inverse, plus in-place"""
    from xrft_b200 import backend as B
    rng = np.random.default_rng(11)
    for n, b in ((5003, 3), (4099, 2), (777, 5)):
        x = cplx(rng, (2, n, b), dt)
        t = torch.from_numpy(x).cuda()
        y = B.fftn(t, axes=[1]).cpu().numpy()
        assert relerr(y, np.fft.fft(x.astype(np.complex128), axis=1)) < TOL[dt] * 20
        yi = B.ifftn(t, axes=[1]).cpu().numpy()
        assert relerr(yi, np.fft.ifft(x.astype(np.complex128), axis=1)) < TOL[dt] * 20
    x2 = cplx(rng, (3, 101, 67), dt)     # two prime axes: strided then contiguous chirp-z
    y2 = B.fftn(torch.from_numpy(x2).cuda(), axes=[1, 2]).cpu().numpy()
    assert relerr(y2, np.fft.fftn(x2.astype(np.complex128), axes=[1, 2])) < TOL[dt] * 20


SMOOTH = [72, 96, 100, 120, 360, 720, 1000, 1440, 2187, 2401, 3125, 3000, 3600, 5000, 6144, 6300]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", SMOOTH + [10000, 12288, 12600])
def test_c2c_smooth_lengths(dt, n):
    """2^a 3^b 5^c 7^d lengths run the mixed-radix shared-memory kernel (smooth.cu) instead of Bluestein: contiguous and strided
    axes, forward and inverse, partial tiles, against numpy's pocketfft in double (the reference's backend, xrft.py:32-36)"""
    from xrft_b200 import backend as B
    if dt == np.float64 and n > 6399:
        pytest.skip("beyond the float64 shared-memory capacity: Bluestein (covered by test_c2c_any_length)")
    rng = np.random.default_rng(n)
    tol = TOL[dt] * 10
    rows = 70 if n < 200 else 5
    x = cplx(rng, (rows, n), dt)
    t = torch.from_numpy(x).cuda()
    assert relerr(B.fftn(t, axes=[1]).cpu().numpy(), np.fft.fft(x.astype(np.complex128), axis=1)) < tol
    assert relerr(B.ifftn(t, axes=[1]).cpu().numpy(), np.fft.ifft(x.astype(np.complex128), axis=1)) < tol
    for b in (7, 37):
        if n * b > 400000:
            continue
        xs = cplx(rng, (2, n, b), dt)
        ts = torch.from_numpy(xs).cuda()
        assert relerr(B.fftn(ts, axes=[1]).cpu().numpy(), np.fft.fft(xs.astype(np.complex128), axis=1)) < tol
        assert relerr(B.ifftn(ts, axes=[1]).cpu().numpy(), np.fft.ifft(xs.astype(np.complex128), axis=1)) < tol


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_smooth_real_transforms_2d(dt):
    """real 2-D transforms of a 360 x 720 grid (a 0.5-degree globe): rfftn / irfftn compose the mixed-radix passes"""
    from xrft_b200 import backend as B
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 360, 720)).astype(dt)
    y = B.rfftn(torch.from_numpy(x).cuda(), axes=[1, 2]).cpu().numpy()
    ref = np.fft.rfftn(x.astype(np.float64), axes=[1, 2])
    assert relerr(y, ref) < TOL[dt] * 10
    back = B.irfftn(torch.from_numpy(ref.astype(y.dtype)).cuda(), axes=[1, 2]).cpu().numpy()
    assert relerr(back, x) < TOL[dt] * 10


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [96, 100, 120, 360, 1000, 1440, 3600, 10000, 12600, 25200, 135, 1125])
def test_real_transforms_smooth_lengths(dt, n):
    """rfft / irfft along a smooth non-power-of-two axis: even lengths run the packed half-length mixed-radix transform with the
    split / merge in its stores / loads (one pass); odd lengths and lengths beyond its capacity take the promoted complex path.
    irfft ignores the imaginary parts of the DC and Nyquist terms like numpy."""
    from xrft_b200 import backend as B
    rng = np.random.default_rng(n)
    rows = 37 if n < 2000 else 5
    x = rng.standard_normal((rows, n)).astype(dt)
    y = B.rfftn(torch.from_numpy(x).cuda(), axes=[1]).cpu().numpy()
    ref = np.fft.rfft(x.astype(np.float64), axis=1)
    assert y.shape == ref.shape
    assert relerr(y, ref) < TOL[dt] * 10
    if n % 2 == 0:
        spec = ref.copy()
        spec[:, 0] += 0.25j
        spec[:, -1] -= 0.5j
        back = B.irfftn(torch.from_numpy(spec.astype(y.dtype)).cuda(), axes=[1]).cpu().numpy()
        assert relerr(back, np.fft.irfft(spec, n=n, axis=1)) < TOL[dt] * 10
    x3 = rng.standard_normal((3, 90, n if n <= 1440 else 96)).astype(dt)     # two axes: strided smooth pass behind the real one
    y3 = B.rfftn(torch.from_numpy(x3).cuda(), axes=[1, 2]).cpu().numpy()
    assert relerr(y3, np.fft.rfftn(x3.astype(np.float64), axes=[1, 2])) < TOL[dt] * 10


@pytest.mark.gpu
def test_permute_flip_kernel():
    """xrftb_permute (axis permutation + reversal; the reference's da.transpose / flipped coordinates) equals numpy"""
    from xrft_b200 import backend as B
    rng = np.random.default_rng(12)
    for dt in (np.float32, np.float64, np.complex64, np.complex128):
        x = rng.standard_normal((3, 4, 5, 6, 7)).astype(dt) if np.dtype(dt).kind != "c" else cplx(rng, (3, 4, 5, 6, 7), np.float32 if dt == np.complex64 else np.float64)
        for perm, flips in (((0, 2, 3, 4, 1), ()), ((4, 3, 2, 1, 0), (1, 3)), ((0, 1, 2, 3, 4), (4,)), ((1, 0, 2, 3, 4), (0,)), ((0, 3, 4, 1, 2), (2,))):
            got = B.permute_flip(torch.from_numpy(x).cuda(), perm, flips).cpu().numpy()
            ref = np.transpose(np.flip(x, axis=flips) if flips else x, perm)
            np.testing.assert_array_equal(got, ref)


def test_permute_transpose_fast_path():
    """permutations that collapse to a batched 2-D transpose (one block of axes moved behind the others: dim="time" of a
    [time][y][x] array) take the tiled shared-memory kernel; sizes off the 32-multiples, all element sizes, with a batch axis"""
    from xrft_b200 import backend as B
    rng = np.random.default_rng(13)
    for dt in (np.float32, np.float64, np.complex64, np.complex128):
        for shape, perm in (((100, 37, 65), (1, 2, 0)), ((100, 37, 65), (2, 0, 1)), ((3, 70, 33, 45), (0, 2, 3, 1)), ((5, 129, 257), (0, 2, 1)),
                            ((1024, 2048), (1, 0)), ((9, 16), (1, 0)), ((4, 3, 50), (0, 2, 1))):
            x = rng.standard_normal(shape).astype(dt) if np.dtype(dt).kind != "c" else cplx(rng, shape, np.float32 if dt == np.complex64 else np.float64)
            got = B.permute_flip(torch.from_numpy(x).cuda(), perm, ()).cpu().numpy()
            np.testing.assert_array_equal(got, np.transpose(x, perm))
