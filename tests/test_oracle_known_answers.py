"""Pins the oracle (numpy restatement of the reference) with the reference's own known-answer tests
(SURVEY.md section 8c): the reference ships no golden vectors, so its arithmetic is pinned by analytic
identities; each test cites the reference test it restates."""
import warnings

import numpy as np
import pandas as pd
import pytest
import scipy.signal as sps

from oracle import xrft_oracle as O

warnings.simplefilter("ignore")


def L(data, dims, coords=None, chunks=None):
    return O.Labelled(np.asarray(data), dims, coords or {d: np.arange(n) for d, n in zip(dims, np.shape(data))}, chunks=chunks)


def test_fft_1d_compositions():
    """test_xrft.py:58-97"""
    rng = np.random.default_rng(0)
    Nx = 16
    x = np.linspace(0, 1.0, Nx)
    dx = x[1] - x[0]
    da = L(rng.random(Nx), ("x",), {"x": x})
    ft = O.fft(da, detrend="constant", true_phase=False, true_amplitude=False)
    assert ft.dims == ("freq_x",)
    fx = np.fft.fftshift(np.fft.fftfreq(Nx, dx))
    np.testing.assert_allclose(ft.coords["freq_x"], fx)
    assert ft.coord_attrs["freq_x"]["spacing"] == fx[1] - fx[0]
    np.testing.assert_allclose(ft.data, np.fft.fftshift(np.fft.fft(da.data - da.data.mean())), atol=1e-14)
    ft = O.fft(da, detrend="linear", true_phase=False, true_amplitude=False)
    np.testing.assert_allclose(ft.data, np.fft.fftshift(np.fft.fft(sps.detrend(da.data))), atol=1e-14)
    bad = L(da.data, ("x",), {"x": np.r_[x[:-1], x[-1] * 2]})
    with pytest.raises(ValueError):
        O.fft(bad)


def test_fft_2d_window_detrend():
    """test_xrft.py:115-138"""
    rng = np.random.default_rng(1)
    N = 16
    da = L(rng.random((N, N)), ("x", "y"))
    ft = O.fft(da, shift=False, true_phase=False, true_amplitude=False)
    np.testing.assert_almost_equal(ft.data, np.fft.fftn(da.data))
    ft = O.fft(da, shift=False, window="hann", detrend="constant", true_phase=False, true_amplitude=False)
    w = sps.windows.hann(N, sym=False)
    np.testing.assert_almost_equal(ft.data, np.fft.fftn((da.data - da.data.mean()) * (w * w[:, None])))


def test_rfft_real_dim_not_last():
    """test_xrft.py:243-270: real_dim is moved last for rfftn and the result transposed back"""
    rng = np.random.default_rng(2)
    da = L(rng.random((8, 10, 12)), ("t", "y", "x"))
    ft = O.fft(da, dim=["y", "x"], real_dim="y", true_phase=False, true_amplitude=False)
    assert ft.dims == ("t", "freq_y", "freq_x")
    ref = np.fft.rfftn(da.data.transpose(0, 2, 1), axes=(1, 2)).transpose(0, 2, 1)
    np.testing.assert_allclose(ft.data, ref)
    np.testing.assert_allclose(ft.coords["freq_y"], np.fft.rfftfreq(10, 1))


def test_power_spectrum_periodogram_and_scalings():
    """test_xrft.py:389-404, 444-495"""
    rng = np.random.default_rng(3)
    N = 16
    da = L(rng.random(N), ("x",))
    f, p = sps.periodogram(da.data, window="rectangular", return_onesided=True)
    np.testing.assert_almost_equal(O.power_spectrum(da, dim="x", real_dim="x", detrend="constant").data, p)
    da3 = L(rng.random((2, N, N)), ("time", "y", "x"))
    ps = O.power_spectrum(da3, dim=["y", "x"], window="hann", detrend="constant")
    daft = O.fft(da3, dim=["y", "x"], window="hann", detrend="constant")
    test = np.real(daft.data * np.conj(daft.data)) / N ** 4
    dk = np.diff(np.fft.fftfreq(N, 1.0))[0]
    np.testing.assert_almost_equal(ps.data, test / dk ** 2)
    ps = O.power_spectrum(da3, dim=["y"], real_dim="x", window="hann", density=False, detrend="constant")
    daft = O.fft(da3, dim=["y"], real_dim="x", detrend="constant", window="hann")
    f2 = np.full(daft.data.shape[-1], 2.0)
    f2[0], f2[-1] = 1.0, 1.0
    np.testing.assert_almost_equal(ps.data, np.real(daft.data * np.conj(daft.data)) * f2)
    with pytest.raises(ValueError):
        O.power_spectrum(da3, dim=["y", "x"], window=None, window_correction=True)


@pytest.mark.parametrize("window_type", ["hann", "bartlett", "tukey", "flattop"])
def test_sine_amplitude_and_energy(window_type):
    """test_xrft.py:406-442"""
    A, fs, fsig = 20, 1e4, 300
    nseg = int(fs // 10)
    tt = np.arange(fs) / fs
    x = A * np.sin(2 * np.pi * fsig * tt)
    la = O.Labelled(x, ("t",), {"t": tt}, chunks={"t": nseg})
    ps = O.power_spectrum(la, dim="t", window=window_type, chunks_to_segments=True, window_correction=True)
    m = ps.data.mean(axis=0)
    np.testing.assert_allclose(np.sqrt(np.trapezoid(m, ps.coords["freq_t"])), A * np.sqrt(2) / 2, rtol=1e-3)
    ps = O.power_spectrum(la, dim="t", window=window_type, chunks_to_segments=True, scaling="spectrum", window_correction=True)
    m = ps.data.mean(axis=0)
    np.testing.assert_allclose(m[np.argmin(np.abs(ps.coords["freq_t"] - fsig))], 0.5 * A ** 2 / 2.0)


def test_parseval():
    """test_xrft.py:693-842 (power and cross spectra, with and without segments / windows)"""
    rng = np.random.default_rng(4)
    N = 16
    for c2s in (False, True):
        ch = {"x": N // 2, "y": N // 2} if c2s else None
        da = O.Labelled(rng.random((N, N)), ("x", "y"), {"x": np.arange(N), "y": np.arange(N)}, chunks=ch)
        da2 = O.Labelled(rng.random((N, N)), ("x", "y"), {"x": np.arange(N), "y": np.arange(N)}, chunks=ch)
        seg = (lambda a: O._stack_chunks(a, ["x", "y"]).data) if c2s else (lambda a: a.data)
        ax = (1, 3) if c2s else (0, 1)
        ps = O.power_spectrum(da, chunks_to_segments=c2s)
        np.testing.assert_almost_equal(ps.data.mean(axis=ax), (seg(da) ** 2).mean(axis=ax), decimal=5)
        cs = O.cross_spectrum(da, da2, chunks_to_segments=c2s)
        np.testing.assert_almost_equal(cs.data.real.mean(axis=ax), (seg(da) * seg(da2)).mean(axis=ax), decimal=5)


def test_cross_phase_known_phase():
    """test_xrft.py:608-633"""
    N = 32
    x = np.linspace(0, 1, num=N, endpoint=False)
    f, p1, p2 = 6, 0, np.pi / 2
    a = L(np.cos(2 * np.pi * f * x + p1), ("x",), {"x": x})
    b = L(np.cos(2 * np.pi * f * x + p2), ("x",), {"x": x})
    cp = O.cross_phase(a, b)
    np.testing.assert_almost_equal(cp.data[np.argmin(np.abs(cp.coords["freq_x"] - f))], p1 - p2)


def test_true_phase_theoretical_sinc():
    """test_xrft.py:1210-1228 (theoretical matching): FT of a gate is a sinc, at reduced length 2^15"""
    N = 2 ** 15
    dx = 2e-4
    x = dx * (np.arange(-N // 2, -N // 2 + N))
    T = 1.0
    y = np.where(np.abs(x) <= T / 2, 1.0, 0.0)
    ft = O.fft(L(y, ("x",), {"x": x}), true_phase=True, true_amplitude=True)
    k = ft.coords["freq_x"]
    sel = np.abs(k) < 20
    np.testing.assert_allclose(ft.data.real[sel], T * np.sinc(k[sel] * T), atol=2e-3)
    np.testing.assert_allclose(ft.data.imag[sel], 0, atol=2e-3)


def test_ifft_fft_round_trip_and_lag():
    """test_xrft.py:1253-1300"""
    rng = np.random.default_rng(5)
    N = 32
    x = 0.25 * np.arange(N) + 3.0
    da = L(rng.random((N, N)), ("y", "x"), {"y": x, "x": x})
    back = O.ifft(O.fft(da))
    np.testing.assert_allclose(back.data.real, da.data, atol=1e-12)
    np.testing.assert_allclose(back.coords["x"], x)
    back0 = O.ifft(O.fft(da), lag=[0.0, 0.0])
    assert abs(back0.coords["x"][N // 2]) < 1e-12


def test_isotropize_sum_conservation_and_errors():
    """test_xrft.py:942-992, 1048-1049"""
    rng = np.random.default_rng(6)
    N = 64
    da = L(rng.random((3, N, N)), ("t", "y", "x"))
    ps = O.power_spectrum(da, dim=["y", "x"])
    for truncate in (True, False):
        iso = O.isotropize(ps, ["freq_y", "freq_x"], truncate=truncate)
        np.testing.assert_allclose(iso.data.sum(axis=-1), ps.data.sum(axis=(1, 2)))
        assert iso.dims == ("t", "freq_r")
    with pytest.raises(ValueError):
        O.isotropic_power_spectrum(da, dim=["t", "y", "x"])


@pytest.mark.parametrize("n,nbins", [(64, 16), (512, 128), (30, 7)])
def test_cut_codes_match_pandas(n, nbins):
    """the pandas.cut restatement used for the radial-bin LUT (xrft.py:921) equals pandas itself"""
    k = np.fft.fftshift(np.fft.fftfreq(n, 0.7))
    fr = np.sqrt(k[:, None] ** 2 + k[None, :] ** 2)
    codes, _ = O.cut_codes(fr, nbins)
    np.testing.assert_array_equal(codes, pd.cut(fr.ravel(), nbins).codes)
    from xrft_b200.api import _cut_codes
    np.testing.assert_array_equal(_cut_codes(fr, nbins), codes)


def test_detrend_recovers_noise_and_closed_form():
    """test_detrend.py:17-24, 84-85, 118-119 + SURVEY F7: plane fit == centred first moments"""
    rng = np.random.default_rng(7)
    for shape, dims in (((40,), ("x",)), ((20, 30), ("y", "x")), ((6, 8, 10), ("z", "y", "x"))):
        noise = rng.standard_normal(shape)
        noise = O.detrend(L(noise, dims), list(dims), "linear").data  # pre-detrended noise
        trend = sum((i + 1) * 3.7 * np.arange(n).reshape([-1 if j == i else 1 for j in range(len(shape))]) for i, n in enumerate(shape))
        out = O.detrend(L(noise + trend + 11.0, dims), list(dims), "linear").data
        np.testing.assert_allclose(out, noise, atol=1e-9)
        # closed form
        a = noise + trend + 11.0
        cf = a - a.mean()
        for ax, n in enumerate(shape):
            idx = (np.arange(n) - (n - 1) / 2).reshape([-1 if j == ax else 1 for j in range(len(shape))])
            cf = cf - (a * idx).sum() / ((idx ** 2).sum() * a.size / n) * idx
        np.testing.assert_allclose(cf, out, atol=1e-9)


def test_pad_coordinates_known_answers():
    """test_padding.py:35-71 and the docstring examples of xrft/padding.py:236-261"""
    x = np.linspace(-4, -1, 4)
    la = L(np.ones((4,)), ("x",), {"x": x})
    p = O.pad(la, {"x": 2})
    np.testing.assert_allclose(p.coords["x"], [-6, -5, -4, -3, -2, -1, 0, 1])
    assert p.coord_attrs["x"]["pad_width"] == 2
    p = O.pad(la, {"x": (1, 4)})
    np.testing.assert_allclose(p.coords["x"], [-5, -4, -3, -2, -1, 0, 1, 2, 3])
    np.testing.assert_allclose(p.data, [0, 1, 1, 1, 1, 0, 0, 0, 0])
    u = O.unpad(p)
    np.testing.assert_allclose(u.coords["x"], x)
    np.testing.assert_allclose(u.data, 1)
