"""Golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py ran /root/reference/xrft
under stand-in xarray/dask/numpy_groupies modules).  CPU: they pin the oracle.  GPU: the CUDA path must match
them (float64 1e-6, float32 1e-3, relative, normwise)."""
import ast
import os
import warnings

import numpy as np
import pytest

from oracle import xrft_oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cases.npz"), allow_pickle=False)
CASES = [ast.literal_eval(s) for s in G["__cases__"]]
# second batch of reference cases (1-D one-sided tukey PSD, Welch segments, unshifted true-phase fft, 1-D cross phase, one-sided
# cross spectrum with window correction, 3-D PSD with 3-D linear detrend, isotropic nfactor / truncate): oracle AND CUDA
CPU_CASES = [ast.literal_eval(s) for s in G["__cases_cpu__"]] if "__cases_cpu__" in G.files else []
CHUNKS = {"ps1d_segments": {"x": 16}}
warnings.simplefilter("ignore")

COORDS = {
    ("x",): lambda: {"x": 0.1 * np.arange(64) - 1.0},
    ("y", "x"): lambda: {"y": 0.5 * np.arange(16) - 3.0, "x": 0.25 * np.arange(32) + 7.0},
}


def coords_for(name, dims, shape):
    if name in ("fft_nonpow2", "ps_nonpow2"):
        return {"t": np.arange(4.0), "y": np.arange(20) * 2.0, "x": np.arange(30) * 3.0}
    if name == "ps2d_f32_config2_like":
        return {"t": np.arange(2.0), "y": np.arange(32) * 1.0, "x": np.arange(64) * 1.0}
    if name.startswith("iso_"):
        return {"t": np.arange(2.0), "y": np.arange(64) * 1.0, "x": np.arange(64) * 1.0}
    if dims == ("x",):
        return {"x": 0.1 * np.arange(64) - 1.0}
    c = {"y": 0.5 * np.arange(16) - 3.0, "x": 0.25 * np.arange(32) + 7.0}
    if dims == ("t", "y", "x"):
        c = {"t": np.arange(3.0), **c}
    return c


def relerr(a, b):
    d = np.linalg.norm(np.asarray(b).ravel())
    return np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / (d if d > 0 else 1.0)


def inputs(name):
    out = []
    i = 0
    while f"{name}__in{i}" in G:
        out.append(G[f"{name}__in{i}"])
        i += 1
    return out


def check_labels(name, out_dims, get_coord, get_attr):
    for key in G.files:
        if key.startswith(name + "__coord__"):
            d = key.split("__coord__")[1]
            np.testing.assert_allclose(get_coord(d), G[key], rtol=1e-12, atol=1e-12, equal_nan=True)
        if key.startswith(name + "__attr__"):
            d, k = key[len(name + "__attr__"):].rsplit("__", 1)
            np.testing.assert_allclose(get_attr(d, k), float(G[key]), rtol=1e-12)


def phase_equal(a, b, weight=None, tol=1e-6):
    d = np.abs(np.angle(np.exp(1j * (np.asarray(a) - np.asarray(b)))))
    if weight is not None:
        d = d[weight]
    return d.max() < tol


ORACLE_FN = {"fft": O.fft, "power_spectrum": O.power_spectrum, "cross_spectrum": O.cross_spectrum, "cross_phase": O.cross_phase,
             "isotropic_power_spectrum": O.isotropic_power_spectrum, "isotropic_cross_spectrum": O.isotropic_cross_spectrum}


@pytest.mark.parametrize("case", CASES + CPU_CASES, ids=[c[0] for c in CASES + CPU_CASES])
def test_oracle_matches_reference_golden(case):
    name, fn, kwrepr, dims, out_dims = case
    kw = ast.literal_eval(kwrepr)
    ins = inputs(name)
    coords = coords_for(name, tuple(dims), ins[0].shape)
    labs = [O.Labelled(a, tuple(dims), coords, chunks=CHUNKS.get(name)) for a in ins]
    if fn == "<lambda>":
        dd = {"detrend_2d": ["y", "x"], "detrend_1d": ["x"], "detrend_3d": ["t", "y", "x"]}[name]
        out = O.detrend(labs[0], dd, kw["detrend_type"])
    else:
        out = ORACLE_FN[fn](*labs, **kw)
    assert list(out.dims) == list(out_dims)
    ref = G[f"{name}__out"]
    if fn == "cross_phase":
        cs = O.cross_spectrum(*labs, **kw).data
        assert phase_equal(out.data, ref, np.abs(cs) > 1e-9 * np.abs(cs).max())
    else:
        assert relerr(out.data, ref) < 1e-12, relerr(out.data, ref)
    check_labels(name, out_dims, lambda d: out.coords[d], lambda d, k: out.coord_attrs[d][k])


def test_oracle_ifft_pad_golden():
    c2 = {"y": 0.5 * np.arange(16) - 3.0, "x": 0.25 * np.arange(32) + 7.0}
    for nm, rd in (("ifft2d", None), ("irfft2d", "freq_x")):
        la = O.Labelled(G[f"{nm}__in0"], ("freq_y", "freq_x"), {"freq_y": G[f"{nm}__freq_y"], "freq_x": G[f"{nm}__freq_x"]},
                        {"freq_y": {"direct_lag": float(G[f"{nm}__lag_y"])}, "freq_x": {"direct_lag": float(G[f"{nm}__lag_x"])}})
        back = O.ifft(la, real_dim=rd)
        assert relerr(back.data, G[f"{nm}__out"]) < 1e-12
    la = O.Labelled(G["ifft2d__in0"], ("freq_y", "freq_x"), {"freq_y": G["ifft2d__freq_y"], "freq_x": G["ifft2d__freq_x"]},
                    {"freq_y": {"direct_lag": float(G["ifft2d__lag_y"])}, "freq_x": {"direct_lag": float(G["ifft2d__lag_x"])}})
    np.testing.assert_allclose(O.ifft(la).coords["x"], G["ifft2d__coord__x"])
    p = O.pad(O.Labelled(G["pad__in0"], ("y", "x"), c2), {"x": (3, 5), "y": 2})
    np.testing.assert_array_equal(p.data, G["pad__out"])
    np.testing.assert_allclose(p.coords["x"], G["pad__coord__x"])
    np.testing.assert_allclose(p.coords["y"], G["pad__coord__y"])
    up = O.unpad(p, {"x": (3, 5), "y": 2})
    np.testing.assert_array_equal(up.data, G["unpad__out"])
    np.testing.assert_allclose(up.coords["x"], G["unpad__coord__x"])
    np.testing.assert_allclose(up.coords["y"], G["unpad__coord__y"])


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + CPU_CASES, ids=[c[0] for c in CASES + CPU_CASES])
def test_cuda_matches_reference_golden(case):
    import xrft_b200 as xrft

    name, fn, kwrepr, dims, out_dims = case
    kw = ast.literal_eval(kwrepr)
    ins = inputs(name)
    coords = coords_for(name, tuple(dims), ins[0].shape)
    das = [xrft.DataArray(a, dims=list(dims), coords=coords) for a in ins]
    if name in CHUNKS:
        das = [d.chunk(CHUNKS[name]) for d in das]
    if fn == "<lambda>":
        dd = {"detrend_2d": ["y", "x"], "detrend_1d": ["x"], "detrend_3d": ["t", "y", "x"]}[name]
        out = xrft.detrend(das[0], dd, kw["detrend_type"])
    else:
        out = getattr(xrft, fn)(*das, **kw)
    assert list(out.dims) == list(out_dims)
    ref = G[f"{name}__out"]
    tol = 1e-3 if ins[0].dtype == np.float32 else 1e-6
    if fn == "cross_phase":
        cs = xrft.cross_spectrum(*das, **kw).values
        assert phase_equal(out.values, ref, np.abs(cs) > 1e-9 * np.abs(cs).max())
    else:
        assert relerr(out.values, ref) < tol, relerr(out.values, ref)
    check_labels(name, out_dims, lambda d: out[d].values, lambda d, k: out[d].attrs[k])


@pytest.mark.gpu
def test_cuda_ifft_pad_golden():
    import xrft_b200 as xrft

    for nm, rd in (("ifft2d", None), ("irfft2d", "freq_x")):
        da = xrft.DataArray(G[f"{nm}__in0"], dims=["freq_y", "freq_x"], coords={
            "freq_y": xrft.DataArray(G[f"{nm}__freq_y"], dims=["freq_y"], attrs={"direct_lag": float(G[f"{nm}__lag_y"])}),
            "freq_x": xrft.DataArray(G[f"{nm}__freq_x"], dims=["freq_x"], attrs={"direct_lag": float(G[f"{nm}__lag_x"])})})
        back = xrft.ifft(da, real_dim=rd)
        assert relerr(back.values, G[f"{nm}__out"]) < 1e-6
        if nm == "ifft2d":
            np.testing.assert_allclose(back["x"].values, G["ifft2d__coord__x"])
    c2 = {"y": 0.5 * np.arange(16) - 3.0, "x": 0.25 * np.arange(32) + 7.0}
    p = xrft.pad(xrft.DataArray(G["pad__in0"], dims=["y", "x"], coords=c2), x=(3, 5), y=2)
    np.testing.assert_array_equal(p.values, G["pad__out"])
    np.testing.assert_allclose(p["x"].values, G["pad__coord__x"])
    np.testing.assert_allclose(p["y"].values, G["pad__coord__y"])
    up = xrft.unpad(p, {"x": (3, 5), "y": 2})
    np.testing.assert_array_equal(up.values, G["unpad__out"])
    np.testing.assert_allclose(up["x"].values, G["unpad__coord__x"])
    np.testing.assert_allclose(up["y"].values, G["unpad__coord__y"])
    # device-resident data takes the CUDA pad / crop kernels
    import torch
    pd = xrft.pad(xrft.DataArray(torch.from_numpy(G["pad__in0"]).cuda(), dims=["y", "x"], coords=c2), x=(3, 5), y=2)
    np.testing.assert_array_equal(pd.values, G["pad__out"])
    np.testing.assert_array_equal(xrft.unpad(pd).values, G["unpad__out"])
