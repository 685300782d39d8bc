"""GPU parity of the xrft-facing API against the oracle (numpy restatement of the reference).

Ported from the reference's own tests (xrft/tests/test_xrft.py, test_detrend.py, test_padding.py):
same compositions, same known answers, compared on the same seeded inputs.  Tolerances:
float64 1e-6 relative (north_star; in practice ~1e-12), float32 1e-3 relative (normwise).
"""
import warnings

import numpy as np
import pytest
import scipy.signal as sps

pytestmark = pytest.mark.gpu

import xrft_b200 as xrft  # noqa: E402
from xrft_b200 import DataArray  # noqa: E402
from oracle import xrft_oracle as O  # noqa: E402

warnings.simplefilter("ignore", FutureWarning)


def lab(da: DataArray) -> O.Labelled:
    coords = {d: da[d].values for d in da.dims if d in da.coords}
    cattrs = {d: dict(da[d].attrs) for d in da.dims if d in da.coords}
    chunks = None
    if da.chunks:
        chunks = {d: c[0] for d, c in zip(da.dims, da.chunks)}
    return O.Labelled(da.values, da.dims, coords, cattrs, chunks)


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0)


def same(out: DataArray, ref: O.Labelled, tol=1e-9, check_attrs=True):
    assert tuple(out.dims) == tuple(ref.dims), (out.dims, ref.dims)
    assert out.shape == ref.data.shape
    for d in ref.dims:
        if d in ref.coords:
            np.testing.assert_allclose(out[d].values.astype(float) if out[d].values.dtype.kind != "M" else out[d].values.astype("i8"),
                                       ref.coords[d].astype(float) if ref.coords[d].dtype.kind != "M" else ref.coords[d].astype("i8"),
                                       rtol=1e-12, atol=1e-12, equal_nan=True)
        if check_attrs and d in ref.coord_attrs:
            for k, v in ref.coord_attrs[d].items():
                assert k in out[d].attrs, (d, k)
                np.testing.assert_allclose(out[d].attrs[k], v, rtol=1e-12)
    assert relerr(out.values, ref.data) < tol, relerr(out.values, ref.data)


def mk(shape, dims, rng, dt=np.float64, spacing=None, offset=None, cplx=False):
    data = rng.standard_normal(shape)
    if cplx:
        data = data + 1j * rng.standard_normal(shape)
    data = data.astype(np.complex128 if cplx and dt == np.float64 else np.complex64 if cplx else dt)
    coords = {}
    for i, (d, n) in enumerate(zip(dims, shape)):
        dx = (spacing or {}).get(d, 0.5 + 0.25 * i)
        x0 = (offset or {}).get(d, 1.0 * i)
        coords[d] = x0 + dx * np.arange(n)
    return DataArray(data, dims=dims, coords=coords)


# ------------------------------------------------------------------------------------------- fft
@pytest.mark.parametrize("kw", [
    dict(), dict(detrend="constant"), dict(detrend="linear"), dict(window="hann"), dict(shift=False),
    dict(true_phase=False, true_amplitude=False), dict(detrend="linear", window="tukey", true_phase=False),
])
@pytest.mark.parametrize("n", [16, 64, 4096])
def test_fft_1d(kw, n):
    rng = np.random.default_rng(n)
    da = mk((n,), ("x",), rng)
    same(xrft.fft(da, **kw), O.fft(lab(da), **kw))


def test_fft_1d_nocoords_and_dft_warning():
    rng = np.random.default_rng(0)
    da = DataArray(rng.random(32), dims=["x"])
    with pytest.warns(FutureWarning):
        ft = xrft.dft(da, detrend="constant")
    assert ft.dims == ("freq_x",)
    np.testing.assert_allclose(ft["freq_x"].values, np.fft.fftshift(np.fft.fftfreq(32, 1)))
    data = da.values - da.values.mean()
    np.testing.assert_allclose(ft.values, np.fft.fftshift(np.fft.fft(data)), atol=1e-12)
    ft = xrft.fft(da, detrend="linear", true_phase=False, true_amplitude=False)
    np.testing.assert_allclose(ft.values, np.fft.fftshift(np.fft.fft(sps.detrend(da.values))), atol=1e-12)


@pytest.mark.parametrize("kw", [
    dict(), dict(dim=["y"]), dict(dim=["x"]), dict(dim=["y", "x"], detrend="linear", window="hann"),
    dict(dim=["y", "x"], detrend="constant", window="hamming", shift=False), dict(dim=["x", "y"]),
    dict(dim=["time"]), dict(real_dim="x", dim=["y", "x"]), dict(real_dim="y", dim=["y", "x"]),
    dict(real_dim="x", dim=["x"], detrend="linear"), dict(dim=["time", "y", "x"], detrend="linear", window="hann"),
    dict(dim=["y", "x"], true_phase=False, true_amplitude=False),
])
def test_fft_nd(kw):
    rng = np.random.default_rng(5)
    da = mk((4, 16, 32), ("time", "y", "x"), rng)
    same(xrft.fft(da, **kw), O.fft(lab(da), **kw))


def test_fft_complex_input_and_decreasing_coords():
    rng = np.random.default_rng(6)
    da = mk((8, 16), ("y", "x"), rng, cplx=True)
    same(xrft.fft(da), O.fft(lab(da)))
    same(xrft.fft(da, dim=["x"], detrend="constant"), O.fft(lab(da), dim=["x"], detrend="constant"))
    dd = DataArray(rng.standard_normal((8, 16)), dims=["y", "x"], coords={"y": np.arange(8, 0, -1) * 1.0, "x": np.arange(16, 0, -1) * 0.5})
    same(xrft.fft(dd), O.fft(lab(dd)))
    ps = xrft.power_spectrum(dd, shift=False, density=True)
    assert (ps.values >= 0.0).all()


def test_fft_keeps_other_coords_and_name_order():
    rng = np.random.default_rng(7)
    da = mk((4, 16, 8), ("time", "y", "x"), rng)
    ft = xrft.fft(da, dim=["y"])
    assert ft.dims == ("time", "freq_y", "x")
    assert "time" in ft.coords and "x" in ft.coords and "y" not in ft.coords
    np.testing.assert_array_equal(xrft.fft(da, dim="y", shift=False).values, xrft.fft(da, dim=["y"], shift=False).values)


def test_fft_float32_mixed_precision_tolerance():
    rng = np.random.default_rng(8)
    da = mk((3, 64, 128), ("t", "y", "x"), rng, dt=np.float32)
    out = xrft.fft(da, dim=["y", "x"], detrend="linear", window="hann")
    ref = O.fft(lab(da), dim=["y", "x"], detrend="linear", window="hann")
    same(out, ref, tol=1e-3, check_attrs=False)


def test_chunks_to_segments():
    rng = np.random.default_rng(9)
    da = mk((64,), ("x",), rng).chunk({"x": 16})
    out = xrft.fft(da, dim=["x"], chunks_to_segments=True)
    same(out, O.fft(lab(da), dim=["x"], chunks_to_segments=True))
    assert out.dims == ("x_segment", "freq_x")
    da2 = mk((32, 64), ("y", "x"), rng).chunk({"y": 16, "x": 32})
    same(xrft.power_spectrum(da2, chunks_to_segments=True, window="hann", window_correction=True),
         O.power_spectrum(lab(da2), chunks_to_segments=True, window="hann", window_correction=True))


# ------------------------------------------------------------------------------------- spectra
@pytest.mark.parametrize("kw", [
    dict(dim=["y", "x"]), dict(dim=["y", "x"], detrend="constant", window="hann"),
    dict(dim=["y", "x"], detrend="linear", window="hann", window_correction=True),
    dict(dim=["y", "x"], scaling="spectrum", window="flattop", window_correction=True),
    dict(dim=["y", "x"], density=False, window="hann", detrend="constant"),
    dict(dim=["y"], real_dim="x", window="hann", density=False, detrend="constant"),
    dict(dim=["x"], real_dim="x", detrend="constant"), dict(dim=["time"], shift=False),
    dict(dim=["y", "x"], real_dim="x", detrend="linear", window="bartlett"),
])
def test_power_spectrum(kw):
    rng = np.random.default_rng(10)
    da = mk((2, 16, 32), ("time", "y", "x"), rng)
    same(xrft.power_spectrum(da, **kw), O.power_spectrum(lab(da), **kw))


def test_power_spectrum_periodogram_known_answer():
    """xrft/tests/test_xrft.py:389-404"""
    rng = np.random.default_rng(11)
    da = DataArray(rng.random(16), dims=["x"], coords={"x": np.arange(16)})
    f, p = sps.periodogram(da.values, window="rectangular", return_onesided=True)
    ps = xrft.power_spectrum(da, dim="x", real_dim="x", detrend="constant")
    np.testing.assert_almost_equal(ps.values, p)


def test_power_spectrum_sine_amplitude_known_answer():
    """xrft/tests/test_xrft.py:406-442 (segment length 1024 instead of 1000: power-of-two path)"""
    A, fs, fsig, nseg = 20, 16384.0, 512.0, 1024
    tt = np.arange(int(fs)) / fs
    x = A * np.sin(2 * np.pi * fsig * tt)
    for window_type in ["hann", "bartlett", "tukey", "flattop"]:
        x_da = DataArray(x, coords=[tt], dims=["t"]).chunk({"t": nseg})
        ps = xrft.power_spectrum(x_da, dim="t", window=window_type, chunks_to_segments=True, window_correction=True).mean("t_segment")
        np.testing.assert_allclose(np.sqrt(np.trapezoid(ps.values, ps["freq_t"].values)), A * np.sqrt(2) / 2, rtol=1e-3)
        ps = xrft.power_spectrum(x_da, dim="t", window=window_type, chunks_to_segments=True, scaling="spectrum",
                                 window_correction=True).mean("t_segment")
        np.testing.assert_allclose(ps.sel(freq_t=fsig).values, 0.5 * A ** 2 / 2.0)


@pytest.mark.parametrize("kw", [
    dict(dim=["y", "x"]), dict(dim=["y", "x"], detrend="constant", window="hann", window_correction=True),
    dict(dim=["y", "x"], true_phase=False, scaling="spectrum"), dict(dim=["x"], real_dim="x", detrend="linear"),
    dict(dim=["y", "x"], density=False, window="hann"),
])
def test_cross_spectrum_and_phase(kw):
    rng = np.random.default_rng(12)
    a = mk((2, 16, 32), ("time", "y", "x"), rng)
    b = mk((2, 16, 32), ("time", "y", "x"), rng)
    cs_ref = O.cross_spectrum(lab(a), lab(b), **kw)
    same(xrft.cross_spectrum(a, b, **kw), cs_ref)
    ref = O.cross_phase(lab(a), lab(b), **kw)
    out = xrft.cross_phase(a, b, **kw)
    assert out.dims == ref.dims
    d = np.angle(np.exp(1j * (out.values - ref.data)))  # compare modulo 2 pi (branch cut at +-pi)
    sig = np.abs(cs_ref.data) > 1e-9 * np.abs(cs_ref.data).max()  # the angle of a ~0 bin (DC after detrend) is noise
    assert np.abs(d[sig]).max() < 1e-7


def test_cross_phase_true_phase_lagged_coords():
    """xrft/tests/test_xrft.py:661-690: different coordinate origins leave a phase ramp in the cross spectrum."""
    rng = np.random.default_rng(13)
    a = mk((16, 32), ("y", "x"), rng, offset={"y": 0.0, "x": 0.0})
    b = mk((16, 32), ("y", "x"), rng, offset={"y": 0.5, "x": 2.0})
    same(xrft.cross_spectrum(a, b), O.cross_spectrum(lab(a), lab(b)))
    out, ref = xrft.cross_phase(a, b), O.cross_phase(lab(a), lab(b))
    d = np.angle(np.exp(1j * (out.values - ref.data)))
    assert np.abs(d).max() < 1e-7


def test_cross_phase_1d_known_answer():
    """xrft/tests/test_xrft.py:608-633: two sines in quadrature have cross phase pi/2 at the signal frequency."""
    N = 32
    x = np.linspace(0, 1, num=N, endpoint=False)
    f = 6
    p1, p2 = 0, np.pi / 2
    da1 = DataArray(np.cos(2 * np.pi * f * x + p1), dims=["x"], coords={"x": x}, name="a")
    da2 = DataArray(np.cos(2 * np.pi * f * x + p2), dims=["x"], coords={"x": x}, name="b")
    cp = xrft.cross_phase(da1, da2)
    assert cp.name == "a_b_phase"
    actual = cp.sel(freq_x=f).values
    np.testing.assert_almost_equal(actual, p1 - p2)
    with pytest.raises(ValueError):
        xrft.cross_phase(da1, DataArray(da2.values[:, None].repeat(4, 1), dims=["x", "z"], coords={"x": x, "z": np.arange(4.)}))


def test_parseval():
    """xrft/tests/test_xrft.py:693-842 (power-of-two sizes)"""
    rng = np.random.default_rng(14)
    N = 16
    dx, dy = 1.0, 0.5
    da = DataArray(rng.random((N, N)), dims=["x", "y"], coords={"x": dx * np.arange(N), "y": dy * np.arange(N)})
    # (1/dxdy) * mean(ps) == mean(da^2)   (xrft/tests/test_xrft.py:727-731)
    ps = xrft.power_spectrum(da, window=None)
    np.testing.assert_almost_equal(ps.values.mean() / (dx * dy), (np.asarray(da.values) ** 2).mean(), decimal=5)
    ps = xrft.power_spectrum(da, window="hann", detrend="constant")
    w = sps.windows.hann(N, sym=False)
    win = w[:, None] * w[None, :]
    dprime = da.values - da.values.mean()
    np.testing.assert_almost_equal(ps.values.mean() / (dx * dy), ((dprime * win) ** 2).mean(), decimal=5)
    # one-sided (real_dim) spectrum integrates to the same variance: sum(ps) dk dl == mean(da^2)
    ps = xrft.power_spectrum(da, real_dim="y")
    np.testing.assert_almost_equal(ps.values.sum() * np.prod([ps[d].attrs["spacing"] for d in ps.dims]),
                                   (da.values ** 2).mean(), decimal=5)


# ------------------------------------------------------------------------------------- inverse
@pytest.mark.parametrize("real_dim", [None, "x"])
def test_ifft_fft_round_trip(real_dim):
    rng = np.random.default_rng(15)
    da = mk((8, 16, 32), ("t", "y", "x"), rng, offset={"y": -3.0, "x": 7.0})
    ft = xrft.fft(da, dim=["y", "x"], real_dim=real_dim)
    ref_ft = O.fft(lab(da), dim=["y", "x"], real_dim=real_dim)
    kw = dict(dim=["freq_y", "freq_x"], real_dim="freq_x" if real_dim else None)
    back = xrft.ifft(ft, **kw)
    ref = O.ifft(ref_ft, **kw)
    same(back, ref, check_attrs=True)
    np.testing.assert_allclose(back.values.real, da.values, atol=1e-10)
    np.testing.assert_allclose(back["x"].values, da["x"].values, atol=1e-12)
    # explicit lag, shift False, true_phase False
    for kw2 in [dict(lag=[0.0, 0.0]), dict(shift=False), dict(true_phase=False, lag=[0.0, 0.0])]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            same(xrft.ifft(ft, **kw, **kw2), O.ifft(ref_ft, **kw, **kw2))


def test_idft_dft_1d():
    rng = np.random.default_rng(16)
    da = mk((64,), ("x",), rng)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        back = xrft.idft(xrft.dft(da, true_phase=True, true_amplitude=True), true_phase=True, true_amplitude=True)
    np.testing.assert_allclose(back.values.real, da.values, atol=1e-12)


# ------------------------------------------------------------------------------------ isotropic
@pytest.mark.parametrize("truncate", [False, True])
@pytest.mark.parametrize("n", [32, 128])
def test_isotropize_and_isotropic_spectra(truncate, n):
    rng = np.random.default_rng(17)
    da = mk((3, n, n), ("t", "y", "x"), rng, spacing={"y": 1.0, "x": 1.0})
    kw = dict(dim=["y", "x"], detrend="constant", window="hann", truncate=truncate)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = xrft.isotropic_power_spectrum(da, **kw)
        ref = O.isotropic_power_spectrum(lab(da), **kw)
    assert out.dims == ref.dims == ("t", "freq_r")
    np.testing.assert_allclose(out["freq_r"].values, ref.coords["freq_r"], rtol=1e-12, equal_nan=True)
    assert relerr(out.values, ref.data) < 1e-10
    # sum conservation (xrft/tests/test_xrft.py:963)
    ps = xrft.power_spectrum(da, dim=["y", "x"], detrend="constant", window="hann")
    np.testing.assert_allclose(out.values.sum(axis=-1), ps.values.sum(axis=(-2, -1)), rtol=1e-10)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        iso2 = xrft.isotropize(ps, ["freq_y", "freq_x"], truncate=truncate)
    assert relerr(iso2.values, ref.data) < 1e-10
    b = mk((3, n, n), ("t", "y", "x"), rng, spacing={"y": 1.0, "x": 1.0})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        outc = xrft.isotropic_cross_spectrum(da, b, **kw)
        refc = O.isotropic_cross_spectrum(lab(da), lab(b), **kw)
    assert relerr(outc.values, refc.data) < 1e-10
    with pytest.raises(ValueError):
        xrft.isotropic_power_spectrum(da, dim=["t", "y", "x"])


def test_isotropic_slope():
    """xrft/tests/test_xrft.py:995-1031: a k^-3 synthetic field isotropizes to slope -3 (N = 512)."""
    N, dL, amp, s = 512, 1.0, 1e1, -3.0
    rng = np.random.default_rng(18)
    k = np.fft.fftshift(np.fft.fftfreq(N, dL))
    kk, ll = np.meshgrid(k, k)
    K = np.sqrt(kk ** 2 + ll ** 2)
    with np.errstate(divide="ignore"):
        spec = np.where(K > 0, amp * K ** (s - 1.0), 0.0)
    phase = rng.uniform(-np.pi, np.pi, (N, N))
    F = np.sqrt(spec) * np.exp(1j * phase)
    field = np.real(np.fft.ifft2(np.fft.ifftshift(F))) * N
    da = DataArray(field, dims=["y", "x"], coords={"y": np.arange(N) * dL, "x": np.arange(N) * dL})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        iso = xrft.isotropic_power_spectrum(da, dim=["y", "x"], window="hann", detrend="constant", truncate=True)
    kr = iso["freq_r"].values
    good = np.isfinite(kr) & (kr > 0)
    sel = good & (np.arange(kr.size) >= 4) & (np.arange(kr.size) < kr.size // 2)
    _, slope, _ = xrft.fit_loglog(kr[sel], iso.values[sel])
    assert abs(slope - s) < 0.25


# ---------------------------------------------------------------------------------- detrend/pad
@pytest.mark.parametrize("dims,detrend_dims", [(("x",), ["x"]), (("y", "x"), ["y", "x"]), (("t", "y", "x"), ["y", "x"]),
                                               (("z", "y", "x"), ["z", "y", "x"]), (("t", "y", "x"), ["t"]), (("t", "y", "x"), ["y"])])
@pytest.mark.parametrize("kind", ["constant", "linear"])
def test_detrend(dims, detrend_dims, kind):
    rng = np.random.default_rng(19)
    shape = (8, 16, 32)[3 - len(dims):]
    da = mk(shape, dims, rng)
    trend = sum((i + 1) * 0.37 * np.arange(n).reshape([-1 if j == i else 1 for j in range(len(shape))]) for i, n in enumerate(shape))
    da = da + trend
    same(xrft.detrend(da, detrend_dims, kind), O.detrend(lab(da), detrend_dims, kind), tol=1e-10)


def test_pad_fft_ifft_unpad_round_trip():
    """xrft/tests/test_padding.py:207-234"""
    rng = np.random.default_rng(20)
    da = mk((24, 48), ("y", "x"), rng, spacing={"y": 0.5, "x": 0.25}, offset={"y": -1.0, "x": 3.0})
    padded = xrft.pad(da, x=8, y=4)
    assert padded.shape == (32, 64)
    ref_p = O.pad(lab(da), {"x": 8, "y": 4})
    np.testing.assert_allclose(padded["x"].values, ref_p.coords["x"])
    assert padded["x"].attrs["pad_width"] == 8
    ft = xrft.fft(padded, real_dim="x", true_phase=True)
    back = xrft.ifft(ft, real_dim="freq_x", true_phase=True)
    back = back.assign_coords({"x": DataArray(back["x"].values, dims=("x",), attrs={"pad_width": 8}),
                               "y": DataArray(back["y"].values, dims=("y",), attrs={"pad_width": 4})})
    un = xrft.unpad(back)
    np.testing.assert_allclose(un.values, da.values, atol=1e-10)
    np.testing.assert_allclose(un["x"].values, da["x"].values, atol=1e-10)


# ------------------------------------------------------------------ large sizes, property checks
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_large_power_spectrum_properties(dt):
    """BASELINE config-2 slice size (4096^2): Parseval + oracle parity on one slice."""
    import torch
    n = 4096 if dt == np.float32 else 2048
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn((2, n, n), generator=g, device="cuda", dtype=torch.float32 if dt == np.float32 else torch.float64)
    x += 0.3 * torch.arange(n, device="cuda") - 0.7 * torch.arange(n, device="cuda")[:, None] + 5
    da = DataArray(x, dims=["time", "y", "x"], coords={"time": np.arange(2.), "y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0})
    ps = xrft.power_spectrum(da, dim=["y", "x"], detrend="linear", window="hann")
    assert ps.dims == ("time", "freq_y", "freq_x")
    # Parseval: sum(ps) * dk*dl == sum((w * detrended)^2) * dx*dy
    xd = xrft.detrend(da, ["y", "x"], "linear").values.astype(np.float64)
    w = sps.windows.hann(n, sym=False)
    e_space = ((xd * (w[:, None] * w[None, :])) ** 2).mean(axis=(1, 2))
    e_spec = ps.values.astype(np.float64).mean(axis=(1, 2))  # dx = dy = 1
    np.testing.assert_allclose(e_spec, e_space, rtol=2e-4 if dt == np.float32 else 1e-10)
    # Hermitian symmetry of the full spectrum of a real field
    p = ps.values[0]
    np.testing.assert_allclose(p[1:, 1:], p[1:, 1:][::-1, ::-1], rtol=1e-5)
    # oracle parity on the first slice (the oracle's dense least-squares detrend takes a few seconds)
    ref = O.power_spectrum(O.Labelled(da.values[:1].astype(np.float64), da.dims, {"y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0}),
                           dim=["y", "x"], detrend="linear", window="hann").data
    assert relerr(ps.values[:1], ref) < (1e-3 if dt == np.float32 else 1e-6)


def test_rfft_irfft_round_trip_padded_f64():
    """BASELINE config 5 at a reduced size: pad -> rfft -> irfft -> unpad + Parseval."""
    import torch
    n, p = 1024, 512
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64)
    da = DataArray(x, dims=["y", "x"], coords={"y": np.arange(n) * 0.5, "x": np.arange(n) * 0.5})
    padded = xrft.pad(da, x=p, y=p)
    ft = xrft.fft(padded, real_dim="x")
    back = xrft.ifft(ft, real_dim="freq_x")
    un = xrft.unpad(back, {"x": p, "y": p})
    err = np.abs(un.values - da.values).max() / np.abs(da.values).max()
    assert err < 1e-6
    ps = xrft.power_spectrum(padded, real_dim="x")
    np.testing.assert_allclose(ps.values.sum() * ps["freq_x"].attrs["spacing"] * ps["freq_y"].attrs["spacing"],
                               (padded.values ** 2).mean(), rtol=1e-10)


# ------------------------------------------------------ the reference's own (non power-of-two) sizes
@pytest.mark.parametrize("shape,dims,kw", [
    ((16,), ("x",), dict(detrend="linear")), ((10,), ("x",), dict(detrend="constant", window="hann")),
    ((15, 19), ("y", "x"), dict(detrend="linear", window="hann")), ((20, 30), ("y", "x"), dict(real_dim="x")),
    ((2, 10, 21), ("t", "y", "x"), dict(dim=["y", "x"], detrend="constant")), ((5, 40, 60), ("t", "y", "x"), dict(dim=["y", "x"], window="tukey")),
    ((6, 8, 10), ("z", "y", "x"), dict(detrend="linear", window="hann")), ((100,), ("x",), dict()),
    ((1000,), ("x",), dict(detrend="linear")), ((3, 258), ("t", "x"), dict(dim=["x"], real_dim="x")), ((366, 4), ("t", "x"), dict(dim=["t"])),
])
def test_reference_sizes_fft_and_power(shape, dims, kw):
    rng = np.random.default_rng(sum(shape))
    da = mk(shape, dims, rng)
    same(xrft.fft(da, **kw), O.fft(lab(da), **kw), tol=1e-8)
    same(xrft.power_spectrum(da, **kw), O.power_spectrum(lab(da), **kw), tol=1e-8)


def test_reference_sizes_round_trip_and_iso():
    rng = np.random.default_rng(77)
    da = mk((20, 30), ("y", "x"), rng)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        back = xrft.ifft(xrft.fft(da))
        np.testing.assert_allclose(back.values.real, da.values, atol=1e-10)
        back = xrft.ifft(xrft.fft(da, real_dim="x"), real_dim="freq_x")
        np.testing.assert_allclose(back.values, da.values, atol=1e-10)
        da2 = mk((3, 40, 60), ("t", "y", "x"), rng, spacing={"y": 1.0, "x": 1.0})
        out = xrft.isotropic_power_spectrum(da2, dim=["y", "x"], detrend="linear", window="hann")
        ref = O.isotropic_power_spectrum(lab(da2), dim=["y", "x"], detrend="linear", window="hann")
    assert relerr(out.values, ref.data) < 1e-8
    np.testing.assert_allclose(out["freq_r"].values, ref.coords["freq_r"], rtol=1e-12)


def test_host_streaming_matches_device_path():
    """numpy in -> numpy out through the chunk-streamed (H2D | kernels | D2H) path equals the device-resident path."""
    import torch
    from xrft_b200 import api as A
    rng = np.random.default_rng(21)
    x = (rng.standard_normal((24, 128, 256)) + 0.1 * np.arange(256)).astype(np.float32)
    c = {"t": np.arange(24.0), "y": np.arange(128) * 1.0, "x": np.arange(256) * 1.0}
    old = (A._STREAM_MIN_BYTES, A._STREAM_CHUNK_BYTES)
    A._STREAM_MIN_BYTES, A._STREAM_CHUNK_BYTES = 1, 5 * 128 * 256 * 4  # 5 items per chunk -> ragged last chunk
    try:
        host = xrft.power_spectrum(DataArray(x, dims=["t", "y", "x"], coords=c), dim=["y", "x"], detrend="linear", window="hann")
        assert isinstance(host.data, np.ndarray)
        outbuf = torch.empty((24, 128, 256), dtype=torch.float32).pin_memory()
        host2 = xrft.power_spectrum(DataArray(x, dims=["t", "y", "x"], coords=c), dim=["y", "x"], detrend="linear", window="hann", out=outbuf)
        cs = xrft.cross_spectrum(DataArray(x, dims=["t", "y", "x"], coords=c), DataArray(x[::-1].copy(), dims=["t", "y", "x"], coords=c), dim=["y", "x"])
        iso_h = xrft.isotropic_power_spectrum(DataArray(x, dims=["t", "y", "x"], coords=c), dim=["y", "x"], detrend="constant", window="hann")
        assert isinstance(iso_h.data, np.ndarray)
    finally:
        A._STREAM_MIN_BYTES, A._STREAM_CHUNK_BYTES = old
    devres = xrft.power_spectrum(DataArray(torch.from_numpy(x).cuda(), dims=["t", "y", "x"], coords=c), dim=["y", "x"], detrend="linear", window="hann")
    np.testing.assert_array_equal(host.values, devres.values)
    np.testing.assert_array_equal(host2.values, devres.values)
    np.testing.assert_array_equal(outbuf.numpy(), devres.values)
    # radial-bin modes stream too (config 4 from host memory): chunk-wise result == device-resident result == oracle
    iso_d = xrft.isotropic_power_spectrum(DataArray(torch.from_numpy(x).cuda(), dims=["t", "y", "x"], coords=c), dim=["y", "x"], detrend="constant", window="hann")
    np.testing.assert_allclose(iso_h.values, iso_d.values, rtol=1e-6)
    iso_r = O.isotropic_power_spectrum(lab(DataArray(x, dims=["t", "y", "x"], coords=c)), dim=["y", "x"], detrend="constant", window="hann")
    assert relerr(iso_h.values, iso_r.data) < 1e-3
    ref = O.cross_spectrum(lab(DataArray(x, dims=["t", "y", "x"], coords=c)), lab(DataArray(x[::-1].copy(), dims=["t", "y", "x"], coords=c)), dim=["y", "x"])
    assert relerr(cs.values, ref.data) < 1e-3


def test_theoretical_matching_440000_points():
    """xrft/tests/test_xrft.py:1210-1228: FT of a gate function is a sinc (440 000-point transform, atol 1e-3)."""
    dx = 0.0001
    x = np.arange(-22.0, 22.0, dx)[:440000]
    T = 1.0
    y = np.where(np.abs(x) <= T / 2, 1.0, 0.0)
    da = DataArray(y, dims=["x"], coords={"x": x})
    ft = xrft.fft(da, true_phase=True, true_amplitude=True)
    k = ft["freq_x"].values
    sel = np.abs(k) < 30
    np.testing.assert_allclose(ft.values.real[sel], T * np.sinc(k[sel] * T), atol=1e-3)
    np.testing.assert_allclose(ft.values.imag[sel], 0, atol=1e-3)
    ref = O.fft(lab(da), true_phase=True, true_amplitude=True)
    assert relerr(ft.values, ref.data) < 1e-9


@pytest.mark.parametrize("dt,shape", [(np.float32, (2, 512, 1024)), (np.float64, (1, 256, 512)), (np.float32, (1, 2048, 2048))])
def test_cross_spectrum_phase_large(dt, shape):
    """two-field fused kernels at multi-stage FFT sizes (both fields interleaved per thread; mirror pass)"""
    rng = np.random.default_rng(31)
    a = mk(shape, ("t", "y", "x"), rng, dt=dt, spacing={"t": 1.0, "y": 1.0, "x": 1.0})
    b = mk(shape, ("t", "y", "x"), rng, dt=dt, spacing={"t": 1.0, "y": 1.0, "x": 1.0})
    kw = dict(dim=["y", "x"], detrend="constant", window="hann")
    tol = 1e-3 if dt == np.float32 else 1e-9
    ref = O.cross_spectrum(lab(a), lab(b), **kw)
    same(xrft.cross_spectrum(a, b, **kw), ref, tol=tol, check_attrs=False)
    out = xrft.cross_phase(a, b, **kw)
    d = np.abs(np.angle(np.exp(1j * (out.values - np.angle(ref.data)))))
    sig = np.abs(ref.data) > 1e-3 * np.abs(ref.data).max()
    # float32: a cell's angle is as accurate as its cross spectrum relative to its own magnitude: cells above 1e-3 of the peak
    # with a normwise error of ~3e-7 of the peak are within ~3e-4 rad; the bound leaves a factor of a few
    assert d[sig].max() < (2e-3 if dt == np.float32 else 1e-6)
    # Hermitian structure of the cross spectrum of real fields: C(-k) = conj(C(k))
    c = xrft.cross_spectrum(a, b, **kw).values[0]
    np.testing.assert_allclose(c[1:, 1:], np.conj(c[1:, 1:][::-1, ::-1]), rtol=1e-4, atol=1e-6 * np.abs(c).max())


# ------------------------------------------------------------------ advisor findings of round 1
def test_batch_over_65535_composed_paths():
    """more than 65535 batch items through the composed path (moments / detrend / transform / radial bins stride over
    the items beyond one grid dimension); the reference has no such limit"""
    rng = np.random.default_rng(65)
    T = 66000
    da = mk((T, 16), ("t", "x"), rng, dt=np.float32)
    da = da + 0.3 * np.arange(16, dtype=np.float32)
    out = xrft.power_spectrum(da, dim=["x"], detrend="linear")
    ref = O.power_spectrum(lab(da), dim=["x"], detrend="linear")
    assert relerr(out.values, ref.data) < 1e-3
    det = xrft.detrend(da, ["x"], "linear")
    assert relerr(det.values, O.detrend(lab(da), ["x"], "linear").data) < 1e-4
    # fused 2-D chain with more than 65535 items in one call
    small = mk((66000, 8, 16), ("t", "y", "x"), rng, dt=np.float32, spacing={"t": 1.0, "y": 1.0, "x": 1.0})
    out2 = xrft.power_spectrum(small, dim=["y", "x"], detrend="constant", window="hann")
    ref2 = O.power_spectrum(lab(small), dim=["y", "x"], detrend="constant", window="hann")
    assert relerr(out2.values[-3:], ref2.data[-3:]) < 1e-3 and relerr(out2.values, ref2.data) < 1e-3


def test_isotropize_more_than_2048_bins():
    rng = np.random.default_rng(66)
    n = 2100
    ps = DataArray(rng.random((2, n, n)), dims=["t", "freq_y", "freq_x"],
                   coords={"t": np.arange(2.0), "freq_y": np.fft.fftshift(np.fft.fftfreq(n)), "freq_x": np.fft.fftshift(np.fft.fftfreq(n))})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        iso = xrft.isotropize(ps, ["freq_y", "freq_x"], nfactor=1, truncate=False)
        ref = O.isotropize(lab(ps), ["freq_y", "freq_x"], nfactor=1, truncate=False)
    assert iso.shape == (2, n) and relerr(iso.values, ref.data) < 1e-10


def test_cross_spectrum_opposite_coordinate_orientation():
    """true_phase flips each array by ITS OWN decreasing coordinates (xrft.py:436-441)"""
    rng = np.random.default_rng(67)
    a = mk((3, 16, 32), ("t", "y", "x"), rng)
    b = mk((3, 16, 32), ("t", "y", "x"), rng)
    b = DataArray(b.values, dims=b.dims, coords={"t": b["t"].values, "y": b["y"].values, "x": b["x"].values[::-1].copy()})
    kw = dict(dim=["y", "x"], true_phase=True)
    same(xrft.cross_spectrum(a, b, **kw), O.cross_spectrum(lab(a), lab(b), **kw), tol=1e-9, check_attrs=False)


def test_container_convention_and_out_validation():
    """numpy in -> numpy out, torch (CUDA) in -> torch (CUDA) out; out= is validated up front"""
    import torch
    rng = np.random.default_rng(68)
    da = mk((4, 16, 32), ("t", "y", "x"), rng)
    assert isinstance(xrft.power_spectrum(da, dim=["y", "x"]).data, np.ndarray)
    assert isinstance(xrft.fft(da, dim=["x"]).data, np.ndarray)
    assert isinstance(xrft.detrend(da, ["x"], "linear").data, np.ndarray)
    dd = DataArray(torch.from_numpy(da.values).cuda(), dims=da.dims, coords={d: da[d].values for d in da.dims})
    r = xrft.power_spectrum(dd, dim=["y", "x"])
    assert isinstance(r.data, torch.Tensor) and r.data.is_cuda
    with pytest.raises(ValueError):
        xrft.power_spectrum(da, dim=["y", "x"], out=np.empty((4, 16, 31)))
    with pytest.raises(ValueError):
        xrft.power_spectrum(da, dim=["y", "x"], out=np.empty((4, 16, 32), dtype=np.float32))
    buf = np.empty((4, 16, 32))
    r2 = xrft.power_spectrum(da, dim=["y", "x"], out=buf)
    np.testing.assert_array_equal(buf, r2.values)


def test_cross_spectrum_and_phase_all_paths():
    """cross_spectrum_and_phase == (cross_spectrum, cross_phase) on the fused z-mode chain (float32 1024^2), the rows-first
    two-field chain (float64; lagged coordinates -> ramps) and the composed path (non power-of-two)"""
    rng = np.random.default_rng(69)
    for shape, dt, off in (((2, 1024, 1024), np.float32, None), ((2, 64, 128), np.float64, {"x": 2.5}), ((2, 20, 30), np.float64, None)):
        a = mk(shape, ("t", "y", "x"), rng, dt=dt, spacing={"t": 1.0, "y": 1.0, "x": 1.0})
        b = mk(shape, ("t", "y", "x"), rng, dt=dt, spacing={"t": 1.0, "y": 1.0, "x": 1.0}, offset=off)
        kw = dict(dim=["y", "x"], detrend="linear", window="hann")
        cs, ph = xrft.cross_spectrum_and_phase(a, b, **kw)
        cs1, ph1 = xrft.cross_spectrum(a, b, **kw), xrft.cross_phase(a, b, **kw)
        assert cs.dims == cs1.dims and ph.dims == ph1.dims
        np.testing.assert_allclose(cs.values, cs1.values, rtol=1e-6, atol=1e-6 * np.abs(cs1.values).max())
        d = np.abs(np.angle(np.exp(1j * (ph.values - ph1.values))))
        assert d[np.abs(cs1.values) > 1e-3 * np.abs(cs1.values).max()].max() < 1e-4
        ref = O.cross_spectrum(lab(a), lab(b), **kw)
        assert relerr(cs.values, ref.data) < (1e-3 if dt == np.float32 else 1e-8)


@pytest.mark.parametrize("mode", ["constant", "edge", "reflect", "symmetric", "wrap"])
def test_pad_modes_on_device(mode):
    """xrft.pad of device-resident data runs the CUDA pad kernel and equals numpy.pad (padding.py:157-181), incl. pads wider
    than the array; the statistic modes fail loudly instead of leaving the device"""
    import torch
    rng = np.random.default_rng(70)
    for dt in (np.float32, np.float64, np.complex128):
        x = rng.standard_normal((3, 5, 7)).astype(dt)
        da = DataArray(torch.from_numpy(x).cuda(), dims=["t", "y", "x"], coords={"t": np.arange(3.0), "y": np.arange(5) * 0.5, "x": np.arange(7) * 0.25})
        kw = dict(constant_values=1.5) if mode == "constant" else {}
        out = xrft.pad(da, {"x": (2, 9), "y": 4}, mode=mode, **kw)
        assert isinstance(out.data, torch.Tensor) and out.data.is_cuda
        np.testing.assert_array_equal(out.values, np.pad(x, ((0, 0), (4, 4), (2, 9)), mode=mode, **kw))
        ref = O.pad(lab(DataArray(x, dims=["t", "y", "x"], coords={"t": np.arange(3.0), "y": np.arange(5) * 0.5, "x": np.arange(7) * 0.25})), {"x": (2, 9), "y": 4})
        np.testing.assert_allclose(out["x"].values, ref.coords["x"])
        np.testing.assert_array_equal(xrft.unpad(out).values, x)
    with pytest.raises(NotImplementedError):
        xrft.pad(da, {"x": 2}, mode="mean")


def test_lazy_pad_fft_ifft_matches_materialised_path():
    """xrft.pad of a CUDA tensor defers the zero padding; xrft.fft(real_dim) reads the unpadded array through load predicates
    (xrftb_fft2r) and must equal the oracle on the materialised padding; xrft.ifft(real_dim) uses the fused inverse chain"""
    import torch
    rng = np.random.default_rng(71)
    for dt, tol in ((np.float64, 1e-9), (np.float32, 1e-3)):
        x = rng.standard_normal((2, 48, 96)).astype(dt)
        c = {"t": np.arange(2.0), "y": np.arange(48) * 0.5 - 3.0, "x": np.arange(96) * 0.25 + 1.0}
        da = DataArray(torch.from_numpy(x).cuda(), dims=["t", "y", "x"], coords=c)
        padded = xrft.pad(da, y=8, x=16)
        assert padded.lazy_pad is not None and padded.shape == (2, 64, 128)
        ft = xrft.fft(padded, dim=["y", "x"], real_dim="x")
        assert padded.lazy_pad is not None            # the transform did not materialise the padding
        ref_p = O.pad(lab(DataArray(x, dims=["t", "y", "x"], coords=c)), {"y": 8, "x": 16})
        ref = O.fft(ref_p, dim=["y", "x"], real_dim="x")
        same(ft, ref, tol=tol)
        np.testing.assert_array_equal(padded.values, ref_p.data)      # materialised on demand by the CUDA pad kernel
        back = xrft.ifft(ft, dim=["freq_y", "freq_x"], real_dim="freq_x")
        refb = O.ifft(ref, dim=["freq_y", "freq_x"], real_dim="freq_x")
        same(back, refb, tol=tol, check_attrs=False)
        un = xrft.unpad(back, {"y": 8, "x": 16})
        assert relerr(un.values, x) < tol
        # unpad of a deferred inverse transform narrows the stored box (crop in the transform's stores): nothing is computed
        # until the cropped data is asked for, and the full-size result is never produced
        from xrft_b200.dataarray import Deferred
        back2 = xrft.ifft(ft, dim=["freq_y", "freq_x"], real_dim="freq_x")
        un2 = xrft.unpad(back2, {"y": 8, "x": 16})
        assert isinstance(back2._store, Deferred) and isinstance(un2._store, Deferred) and un2.shape == (2, 48, 96)
        assert relerr(un2.values, x) < tol and isinstance(back2._store, Deferred)
        np.testing.assert_allclose(un2["x"].values, c["x"], atol=1e-9)
        odd = back2.isel(y=slice(3, 40), x=slice(5, 100))          # odd offsets: scalar store path
        assert relerr(odd.values, refb.data[:, 3:40, 5:100]) < tol
        # variants of the inverse: unshifted output, true_phase off, explicit lag
        for kw in (dict(shift=False), dict(true_phase=False), dict(lag=[0.0, 0.0])):
            b2 = xrft.ifft(ft, dim=["freq_y", "freq_x"], real_dim="freq_x", **kw)
            r2 = O.ifft(ref, dim=["freq_y", "freq_x"], real_dim="freq_x", **kw)
            same(b2, r2, tol=tol, check_attrs=False)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_welch_segment_mean_is_an_epilogue_reduction(dt):
    """chunks_to_segments=True followed by .mean('<dim>_segment') (Welch; xrft/tests/test_xrft.py:273-337, 406-442): on device
    data the per-segment spectra stay deferred and the mean is reduced inside the spectral epilogue
    (xrftb_spectral_post_segmean); any other access computes the per-segment spectra as before"""
    import torch
    from xrft_b200.dataarray import Deferred
    rng = np.random.default_rng(72)
    tol = 1e-9 if dt == np.float64 else 1e-3
    x = (rng.standard_normal((3, 256, 5)) + 0.2 * np.arange(256)[None, :, None]).astype(dt)
    c = {"a": np.arange(3.0), "t": np.arange(256) * 0.5, "b": np.arange(5.0)}
    da_h = DataArray(x, dims=["a", "t", "b"], coords=c).chunk({"t": 64})
    da_d = DataArray(torch.from_numpy(x).cuda(), dims=["a", "t", "b"], coords=c).chunk({"t": 64})
    for kw in (dict(dim="t", chunks_to_segments=True, window="hann", detrend="linear"),
               dict(dim="t", real_dim="t", chunks_to_segments=True, window="hann", detrend="constant", window_correction=True)):
        ref = O.power_spectrum(lab(da_h), **kw)
        ps = xrft.power_spectrum(da_d, **kw)
        assert isinstance(ps._store, Deferred) and ps.dims == ref.dims and ps.shape == ref.data.shape
        welch = ps.mean("t_segment")
        assert isinstance(ps._store, Deferred)                       # the per-segment spectra were never materialised
        assert welch.dims == tuple(d for d in ref.dims if d != "t_segment")
        ax = ref.dims.index("t_segment")
        assert relerr(welch.values, ref.data.mean(axis=ax)) < tol
        same(ps, ref, tol=tol, check_attrs=False)                   # other accesses compute them
        np.testing.assert_allclose(welch["freq_t"].values, ref.coords["freq_t"], rtol=1e-12)
    # cross spectrum of two chunked series, segments on the trailing axis
    y = (rng.standard_normal((4, 512))).astype(dt); z = (y + 0.3 * rng.standard_normal((4, 512))).astype(dt)
    cc = {"a": np.arange(4.0), "t": np.arange(512) * 1.0}
    kw = dict(dim="t", chunks_to_segments=True, window="hann", detrend="constant")
    mk2 = lambda v, dev: DataArray(torch.from_numpy(v).cuda() if dev else v, dims=["a", "t"], coords=cc).chunk({"t": 128})
    cs = xrft.cross_spectrum(mk2(y, True), mk2(z, True), **kw)
    refc = O.cross_spectrum(lab(mk2(y, False)), lab(mk2(z, False)), **kw)
    assert relerr(cs.mean("t_segment").values, refc.data.mean(axis=refc.dims.index("t_segment"))) < tol


def test_lazy_chunk_iterator_streams_through_the_gpu():
    """xrft_b200.stream: a generator of host chunks in, a generator of host results out (the role dask's chunk iteration
    plays in the reference, xrft.py:925-943); equals the whole-array call chunk by chunk"""
    import types
    rng = np.random.default_rng(73)
    x = (rng.standard_normal((10, 64, 128)) + 0.1 * np.arange(128)).astype(np.float32)
    y = (x[::-1] + 0.2 * rng.standard_normal(x.shape)).astype(np.float32)
    c = lambda lo, hi: {"t": np.arange(lo, hi) * 1.0, "y": np.arange(64) * 1.0, "x": np.arange(128) * 1.0}
    consumed = []

    def chunks(a):
        for lo in range(0, 10, 3):       # ragged last chunk
            consumed.append(lo)
            yield DataArray(a[lo:lo + 3], dims=["t", "y", "x"], coords=c(lo, min(lo + 3, 10)))

    kw = dict(dim=["y", "x"], detrend="linear", window="hann")
    it = xrft.stream(xrft.power_spectrum, chunks(x), **kw)
    assert isinstance(it, types.GeneratorType) and consumed == []          # nothing happens before the first request
    parts = []
    for r in it:
        assert isinstance(r.data, np.ndarray) and r.dims == ("t", "freq_y", "freq_x")
        parts.append(r)
    whole = xrft.power_spectrum(DataArray(x, dims=["t", "y", "x"], coords=c(0, 10)), **kw)
    np.testing.assert_array_equal(np.concatenate([p.values for p in parts]), whole.values)
    np.testing.assert_array_equal(np.concatenate([p["t"].values for p in parts]), np.arange(10.0))
    # two zipped iterables (cross spectrum) and the radial-bin mode
    cs = np.concatenate([r.values for r in xrft.stream(xrft.cross_spectrum, chunks(x), chunks(y), **kw)])
    ref = O.cross_spectrum(lab(DataArray(x, dims=["t", "y", "x"], coords=c(0, 10))), lab(DataArray(y, dims=["t", "y", "x"], coords=c(0, 10))), **kw)
    assert relerr(cs, ref.data) < 1e-3
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        iso = np.concatenate([r.values for r in xrft.stream(xrft.isotropic_power_spectrum, chunks(x), **kw)])
        riso = O.isotropic_power_spectrum(lab(DataArray(x, dims=["t", "y", "x"], coords=c(0, 10))), **kw)
    assert relerr(iso, riso.data) < 1e-3
