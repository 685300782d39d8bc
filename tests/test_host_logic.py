"""Host-side bookkeeping (no GPU): the DataArray shim, coordinate/metadata plans against the oracle, and the
reference's error behaviour (all raised before any numerics)."""
import warnings

import numpy as np
import pytest

import xrft_b200 as xrft
from xrft_b200 import DataArray
from xrft_b200 import api as A
from oracle import xrft_oracle as O

warnings.simplefilter("ignore")


def da3(rng=None):
    rng = rng or np.random.default_rng(0)
    return DataArray(rng.random((4, 6, 8)), dims=["t", "y", "x"], coords={"t": np.arange(4.), "y": 0.5 * np.arange(6), "x": 2.0 + 0.25 * np.arange(8)}, name="v")


# ------------------------------------------------------------------ shim
def test_shim_basics_and_broadcast():
    a = da3()
    assert a.dims == ("t", "y", "x") and a.shape == (4, 6, 8) and a.sizes["y"] == 6
    assert a.get_axis_num("x") == 2 and a.get_axis_num(["t", "x"]) == (0, 2)
    np.testing.assert_allclose(a["x"].values, 2.0 + 0.25 * np.arange(8))
    np.testing.assert_allclose(a.x.values, a["x"].values)
    b = a.transpose("x", "t", "y")
    assert b.dims == ("x", "t", "y") and b.shape == (8, 4, 6)
    m = a.mean("t")
    assert m.dims == ("y", "x")
    np.testing.assert_allclose((a - m).values, a.values - a.values.mean(axis=0))
    w = DataArray(np.arange(8.), dims=["x"])
    np.testing.assert_allclose((a * w).values, a.values * np.arange(8.))
    np.testing.assert_allclose((w * a).transpose("t", "y", "x").values, a.values * np.arange(8.))
    np.testing.assert_allclose(np.abs(a - 0.5).values, np.abs(a.values - 0.5))
    np.testing.assert_allclose((a ** 2).values, a.values ** 2)
    assert (a * 2).name == "v" and (a * a.rename("w")).name is None
    s = a.swap_dims({"x": "k"})
    assert s.dims == ("t", "y", "k") and s["x"].dims == ("k",)
    d = a.isel(x=slice(2, 5), t=0)
    assert d.dims == ("y", "x") and d.shape == (6, 3)
    np.testing.assert_allclose(d["x"].values, a["x"].values[2:5])
    assert float(a[0, 0, 0]) == a.values[0, 0, 0]
    assert a.sel(y=1.0).dims == ("t", "x")
    c = a.chunk({"t": 2})
    assert c.chunks == ((2, 2), (6,), (8,)) and a.chunks is None
    sb = DataArray([3., 1., 2.], dims=["q"], coords={"q": [30., 10., 20.]}).sortby("q")
    np.testing.assert_allclose(sb.values, [1., 2., 3.])
    p = DataArray(np.ones((2, 3)), dims=["y", "x"], coords={"y": [0., 1.], "x": [0., 1., 2.]}).pad({"x": (1, 2)})
    assert p.shape == (2, 6) and np.isnan(p["x"].values[0])
    with xrft.set_options(keep_attrs=True):
        k = DataArray(np.arange(3.), dims=["k"], coords={"k": np.arange(3.)}, attrs={"spacing": 1.0})
        assert (k + 1.0).attrs["spacing"] == 1.0
    with pytest.raises(ValueError):
        DataArray(np.ones((2, 3)), dims=["y", "x"], coords={"x": [0., 1.]})


# ------------------------------------------------------------------ plans vs oracle metadata
@pytest.mark.parametrize("kw", [dict(), dict(dim=["y", "x"]), dict(dim=["x"], real_dim="x"), dict(dim=["y", "x"], real_dim="y"),
                                dict(shift=False), dict(true_phase=False)])
def test_fft_plan_matches_oracle_metadata(kw):
    a = da3()
    P = A._fft_prepare(a, 1e-3, kw.get("dim"), kw.get("real_dim"), kw.get("shift", True), kw.get("true_phase", True), False, "freq_", None)
    ref = O.fft(O.Labelled(a.values, a.dims, {d: a[d].values for d in a.dims}), **kw)
    out = A._label_output(P, np.zeros([ref.data.shape[ref.dims.index(A._new_name(d, "freq_") if d in P["dim"] else d)] for d in P["da"].dims]))
    assert out.dims == ref.dims
    for d in ref.dims:
        np.testing.assert_allclose(out[d].values, ref.coords[d])
        for k, v in ref.coord_attrs.get(d, {}).items():
            np.testing.assert_allclose(out[d].attrs[k], v)
    if kw.get("true_phase", True):
        for r, N, ku, lag in zip(A._phase_ramps(P), P["N"], P["k_unshifted"], P["lag_x"]):
            x = np.random.default_rng(1).random(N)
            got = np.fft.fft(x) * r if len(ku) == N else np.fft.rfft(x) * r
            want = (np.fft.fft if len(ku) == N else np.fft.rfft)(np.fft.ifftshift(x)) * np.exp(-2j * np.pi * ku * lag)
            np.testing.assert_allclose(got, want, atol=1e-12)


def test_spectrum_scale_factors_match_oracle():
    a = da3()
    la = O.Labelled(a.values, a.dims, {d: a[d].values for d in a.dims})
    for scaling in ("density", "spectrum"):
        for wc in (False, True):
            P = A._fft_prepare(a, 1e-3, ["y", "x"], None, True, False, False, "freq_", None)
            s = A._spectrum_scale(P, a, ["y", "x"], None, scaling, wc, "hann")
            ref = O.power_spectrum(la, dim=["y", "x"], scaling=scaling, window="hann", window_correction=wc).data
            raw = np.abs(np.fft.fftshift(np.fft.fftn(la.data * O.apply_window(la, ["y", "x"], "hann")[0], axes=(1, 2)), axes=(1, 2))) ** 2
            np.testing.assert_allclose(raw * s, ref, rtol=1e-12)


def test_radial_bins_match_oracle():
    k = np.fft.fftshift(np.fft.fftfreq(32, 0.5))
    l = np.fft.fftshift(np.fft.fftfreq(64, 1.0))
    codes, nbins, kr = A._radial_bins(k, l, 4, True)
    fr = np.sqrt(k[:, None] ** 2 + l[None, :] ** 2)
    c2, _ = O.cut_codes(fr, nbins)
    np.testing.assert_array_equal(codes.ravel(), c2)
    assert nbins == 8 and kr.shape == (8,)


# ------------------------------------------------------------------ error behaviour (reference tests cited)
def test_errors_raised_before_numerics():
    a = da3()
    bad = DataArray(a.values, dims=a.dims, coords={"t": np.arange(4.), "y": 0.5 * np.arange(6), "x": np.r_[np.arange(7) * 0.25, 9.0]})
    with pytest.raises(ValueError):  # test_xrft.py:94-97 uneven spacing
        xrft.fft(bad, dim=["x"])
    with pytest.raises(ValueError):  # zero spacing
        xrft.fft(DataArray(np.ones(4), dims=["x"], coords={"x": np.zeros(4)}))
    with pytest.raises(ValueError):  # test_xrft.py:240-241 real_dim not a dim
        xrft.fft(a, real_dim="w")
    with pytest.raises(TypeError):  # test_xrft.py:1132-1135 spacing_tol must be a number
        xrft.fft(a, dim=["x"], spacing_tol="string")
    with pytest.raises(ValueError):  # test_xrft.py:1364-1379 string coordinate
        xrft.fft(DataArray(np.ones(3), dims=["x"], coords={"x": np.array(["a", "b", "c"])}))
    extra = a.assign_coords(lon=DataArray(np.ones((6, 8)), dims=["y", "x"]))
    with pytest.raises(ValueError):  # test_xrft.py:1344-1362 coordinates sharing the transform dim
        xrft.fft(extra, dim=["x"])
    with pytest.raises(NotImplementedError):  # xrft.py:73-75
        xrft.fft(a, dim=["x"], window="nope")
    with pytest.raises(NotImplementedError):  # detrend.py:46-50
        xrft.detrend(a, ["x"], detrend_type="quadratic")
    with pytest.raises(ValueError):  # xrft.py:116-117
        A._stack_chunks(DataArray(np.ones(10), dims=["x"]).chunk({"x": 4}), ["x"])
    with pytest.raises(ValueError):  # detrend.py:62-63
        xrft.detrend(DataArray(np.ones((8, 8)), dims=["y", "x"]).chunk({"x": 4}), ["x"], "linear")
    with pytest.raises(ValueError):  # xrft.py:1078-1079
        xrft.isotropic_power_spectrum(a, dim=["t", "y", "x"])
    with pytest.raises(ValueError):  # xrft.py:650-653
        xrft.power_spectrum(a, dim=["y", "x"], window=None, window_correction=True)
    ft_like = DataArray(np.ones((8,), complex), dims=["freq_x"], coords={"freq_x": np.arange(8.)})
    with pytest.raises(ValueError):  # xrft.py:602-606 not centred on zero
        xrft.ifft(ft_like)
    with pytest.raises(ValueError):  # xrft.py:564-565
        xrft.ifft(ft_like, lag=[0.0, 1.0])


def test_pad_unpad_host_path():
    """test_padding.py:35-205 (coords are extrapolated linearly; attrs; slices; bad coords)"""
    da = DataArray(np.arange(9.).reshape(3, 3) + 1, dims=("y", "x"), coords={"x": [0, 1, 2], "y": [-5, -4, -3]})
    p = xrft.pad(da, x=2, y=1)
    assert p.shape == (5, 7)
    np.testing.assert_allclose(p["x"].values, [-2, -1, 0, 1, 2, 3, 4])
    np.testing.assert_allclose(p["y"].values, [-6, -5, -4, -3, -2])
    assert p["x"].attrs["pad_width"] == 2 and p["y"].attrs["pad_width"] == 1
    np.testing.assert_allclose(p.values[1:4, 2:5], da.values)
    assert p.values.sum() == da.values.sum()
    u = xrft.unpad(p)
    np.testing.assert_allclose(u.values, da.values)
    assert "pad_width" not in u["x"].attrs
    u2 = xrft.unpad(p, x=1, y=1)
    assert u2.shape == (3, 5)
    p2 = xrft.pad(da, x=(1, 4))
    np.testing.assert_allclose(p2["x"].values, [-1, 0, 1, 2, 3, 4, 5, 6])
    assert p2["x"].attrs["pad_width"] == (1, 4)
    np.testing.assert_allclose(xrft.pad(da, x=2, mode="edge").values[:, :2], da.values[:, :1].repeat(2, 1))
    with pytest.raises(ValueError):
        xrft.unpad(da)
    with pytest.raises(ValueError):
        xrft.pad(da.assign_coords(lon=DataArray(np.ones((3, 3)), dims=["y", "x"])), x=1)


def test_fit_loglog():
    x = np.arange(1, 50.)
    y = 3.0 * x ** -2.5
    _, a, b = xrft.fit_loglog(x, y)
    np.testing.assert_allclose([a, 2 ** b], [-2.5, 3.0])


def test_calendar_coordinates_through_a_cftime_stand_in(monkeypatch):
    """cftime is not installable in this image: a stand-in module with cftime's date2num signature drives the
    calendar branch of the coordinate helpers (xrft/xrft.py:195-234, 269-274) -- spacing, lag and validity of a time axis
    whose values carry a `calendar` attribute."""
    import datetime as dt
    import sys
    import types
    from xrft_b200 import api as A
    from xrft_b200 import DataArray

    class NoLeap(dt.datetime):
        calendar = "noleap"

    def date2num(dates, units, calendar):
        assert units == "seconds since 1800-01-01 00:00:00" and calendar == "noleap"
        ref = dt.datetime(1800, 1, 1)
        f = lambda d: (d - ref).total_seconds()
        return f(dates) if isinstance(dates, dt.datetime) else np.array([f(d) for d in np.asarray(dates).ravel()]).reshape(np.shape(dates))

    monkeypatch.setitem(sys.modules, "cftime", types.SimpleNamespace(date2num=date2num))
    t = np.array([NoLeap(2001, 1, 1) + dt.timedelta(hours=6 * i) for i in range(16)], dtype=object)
    coord = DataArray(np.zeros(16), dims=["time"], coords={"time": t})["time"]
    assert A._is_valid_fft_coord(coord)
    assert np.allclose(A._diff_coord(coord), 6 * 3600.0)
    assert A._get_coordinate_spacing(coord, 1e-3) == pytest.approx(6 * 3600.0)
    assert A._lag_coord(coord) == pytest.approx(date2num(t[8], "seconds since 1800-01-01 00:00:00", "noleap"))
    rev = DataArray(np.zeros(16), dims=["time"], coords={"time": t[::-1].copy()})["time"]
    assert A._lag_coord(rev) == pytest.approx(date2num(t[8], "seconds since 1800-01-01 00:00:00", "noleap"))
    bad = DataArray(np.zeros(3), dims=["time"], coords={"time": np.array(["a", "b", "c"], dtype=object)})["time"]
    assert not A._is_valid_fft_coord(bad)
