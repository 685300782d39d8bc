#!/usr/bin/env python
"""Generate golden vectors by executing the UNMODIFIED reference source (/root/reference/xrft).

The reference needs xarray, dask and numpy_groupies, none of which exist in this image.
This script installs stand-ins into sys.modules *for this process only*:
  xarray          -> xrft_b200.dataarray.DataArray (+ apply_ufunc / set_options restated below)
  dask.array      -> only looked up when `da.chunks` is truthy; never reached here (inputs are unchunked)
  numpy_groupies  -> aggregate(sum|mean) == np.bincount (what SURVEY.md section 8c states)
and then imports /root/reference/xrft unchanged, runs seeded cases through its public functions and stores
inputs + outputs in tests/golden/reference_cases.npz.  Every arithmetic line that runs is the reference's.

Run from the repo root in the build container:  python tests/golden/make_golden.py
(/root/reference does not exist on the GPU box; the committed .npz is what the tests read.)
"""
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from xrft_b200 import dataarray as shim  # noqa: E402
from xrft_b200.dataarray import DataArray  # noqa: E402


# ------------------------------------------------------------------ xarray stand-in
def apply_ufunc(func, *args, input_core_dims=None, output_core_dims=None, vectorize=False, kwargs=None, **_):
    kwargs = kwargs or {}
    das = [a for a in args if isinstance(a, DataArray)]
    first = das[0]
    if not input_core_dims:  # elementwise / whole-array call (sps.detrend(da, axis), np.angle(da))
        res = func(*[a.values if isinstance(a, DataArray) else a for a in args], **kwargs)
        return first._replace(data=np.asarray(res))
    # move core dims last for every DataArray argument
    moved = []
    for a, core in zip(args, input_core_dims):
        lead = [d for d in a.dims if d not in core]
        moved.append(a.transpose(*(lead + list(core))))
    lead_dims = [d for d in moved[0].dims if d not in input_core_dims[0]]
    lead_shape = [moved[0].sizes[d] for d in lead_dims]
    out_core = list(output_core_dims[0])
    if vectorize:
        out = np.empty(moved[0].shape, dtype=moved[0].values.dtype)
        for idx in np.ndindex(*lead_shape):
            out[idx] = func(moved[0].values[idx], **kwargs)
        res = DataArray(out, dims=lead_dims + out_core)
    else:
        vals = [m.values for m in moved]
        r = func(*vals, **kwargs)
        res = DataArray(np.asarray(r), dims=lead_dims + out_core)
    for name, c in first.coords.items():
        if all(d in res.dims for d in c.dims) and name not in res.coords and all(res.sizes[d] == c.sizes[d] for d in c.dims):
            res = res.assign_coords({name: c})
    return res


xr = types.ModuleType("xarray")
xr.DataArray = DataArray
xr.apply_ufunc = apply_ufunc
xr.set_options = shim.set_options
core = types.ModuleType("xarray.core")
utils = types.ModuleType("xarray.core.utils")
utils.either_dict_or_kwargs = shim.either_dict_or_kwargs
core.utils = utils
xr.core = core
sys.modules.update({"xarray": xr, "xarray.core": core, "xarray.core.utils": utils})

dask = types.ModuleType("dask")
dask.delayed = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("dask is not available"))
dsar = types.ModuleType("dask.array")
dsar.fft = None
dask.array = dsar
sys.modules.update({"dask": dask, "dask.array": dsar})

ng = types.ModuleType("numpy_groupies")


def aggregate(group_idx, a, func="sum", size=None, fill_value=0, dtype=None, axis=-1):
    a = np.asarray(a)
    lead = a.shape[:-1]
    out = np.full(lead + (size,), fill_value, dtype=dtype or (a.dtype if func == "sum" else float))
    cnt = np.bincount(group_idx, minlength=size)
    for i in np.ndindex(*lead):
        if np.iscomplexobj(a):
            s = np.bincount(group_idx, a[i].real, size) + 1j * np.bincount(group_idx, a[i].imag, size)
        else:
            s = np.bincount(group_idx, a[i], size)
        if func == "mean":
            s = np.where(cnt > 0, s / np.where(cnt == 0, 1, cnt), fill_value)
        out[i] = s
    return out


ng.aggregate = aggregate
sys.modules["numpy_groupies"] = ng

sys.path.insert(0, "/root/reference")
import xrft as ref  # noqa: E402  -- the unmodified reference

warnings.simplefilter("ignore")


def da_of(arr, dims, coords):
    return DataArray(arr, dims=dims, coords=coords)


def main():
    rng = np.random.default_rng(20260925)
    store = {}
    cases = []

    cpu_cases = []   # cases added after the last GPU run of the round: they pin the oracle only (promote to `cases` next round)

    def add(name, fn, inputs, kwargs, dims, coords, gpu=True, chunks=None):
        das = [da_of(a, dims, coords) for a in inputs]
        if chunks:
            das = [d.chunk(chunks) for d in das]
        out = fn(*das, **kwargs)
        store[f"{name}__out"] = np.asarray(out.values)
        for i, a in enumerate(inputs):
            store[f"{name}__in{i}"] = a
        for d in out.dims:
            if d in out.coords and out[d].values.dtype.kind in "fiu":
                store[f"{name}__coord__{d}"] = np.asarray(out[d].values, dtype=float)
                for k, v in out[d].attrs.items():
                    if isinstance(v, (int, float, np.floating, np.integer)):
                        store[f"{name}__attr__{d}__{k}"] = np.asarray(float(v))
        (cases if gpu else cpu_cases).append((name, fn.__name__, repr(kwargs), dims, list(out.dims)))
        print(name, fn.__name__, kwargs, "->", out.dims, out.shape)

    c2 = {"y": 0.5 * np.arange(16) - 3.0, "x": 0.25 * np.arange(32) + 7.0}
    c3 = {"t": np.arange(3.0), **c2}
    a2 = rng.standard_normal((16, 32)) + 0.3 * np.arange(32) - 0.7 * np.arange(16)[:, None] + 5
    b2 = rng.standard_normal((16, 32))
    a3 = rng.standard_normal((3, 16, 32)) + 0.2 * np.arange(32)
    b3 = rng.standard_normal((3, 16, 32))
    a1 = rng.standard_normal(64) + 0.1 * np.arange(64)
    c1 = {"x": 0.1 * np.arange(64) - 1.0}
    a20 = rng.standard_normal((4, 20, 30))
    c20 = {"t": np.arange(4.0), "y": np.arange(20) * 2.0, "x": np.arange(30) * 3.0}
    f32 = (rng.standard_normal((2, 32, 64)) + 0.3 * np.arange(64) - 0.7 * np.arange(32)[:, None] + 5).astype(np.float32)
    c32 = {"t": np.arange(2.0), "y": np.arange(32) * 1.0, "x": np.arange(64) * 1.0}
    iso_in = rng.standard_normal((2, 64, 64))
    ciso = {"t": np.arange(2.0), "y": np.arange(64) * 1.0, "x": np.arange(64) * 1.0}

    add("fft1d_default", ref.fft, [a1], {}, ("x",), c1)
    add("fft1d_linear_hann", ref.fft, [a1], dict(detrend="linear", window="hann"), ("x",), c1)
    add("fft1d_numpy_like", ref.fft, [a1], dict(true_phase=False, true_amplitude=False, shift=False), ("x",), c1)
    add("fft2d_default", ref.fft, [a2], {}, ("y", "x"), c2)
    add("fft2d_linear_hann", ref.fft, [a2], dict(detrend="linear", window="hann"), ("y", "x"), c2)
    add("fft3d_dim_y", ref.fft, [a3], dict(dim=["y"], detrend="constant"), ("t", "y", "x"), c3)
    add("fft3d_realdim_x", ref.fft, [a3], dict(dim=["y", "x"], real_dim="x", window="tukey"), ("t", "y", "x"), c3)
    add("fft3d_realdim_y", ref.fft, [a3], dict(dim=["y", "x"], real_dim="y"), ("t", "y", "x"), c3)
    add("fft3d_all_linear", ref.fft, [a3], dict(detrend="linear", window="hamming"), ("t", "y", "x"), c3)
    add("fft_nonpow2", ref.fft, [a20], dict(dim=["y", "x"], detrend="linear", window="hann"), ("t", "y", "x"), c20)
    add("ps2d_linear_hann", ref.power_spectrum, [a3], dict(dim=["y", "x"], detrend="linear", window="hann"), ("t", "y", "x"), c3)
    add("ps2d_spectrum_corr", ref.power_spectrum, [a3], dict(dim=["y", "x"], scaling="spectrum", window="flattop", window_correction=True), ("t", "y", "x"), c3)
    add("ps2d_realdim", ref.power_spectrum, [a3], dict(dim=["y", "x"], real_dim="x", detrend="constant", window="hann", window_correction=True), ("t", "y", "x"), c3)
    add("ps2d_f32_config2_like", ref.power_spectrum, [f32], dict(dim=["y", "x"], detrend="linear", window="hann"), ("t", "y", "x"), c32)
    add("ps_nonpow2", ref.power_spectrum, [a20], dict(dim=["y", "x"], detrend="constant", window="bartlett"), ("t", "y", "x"), c20)
    add("cs2d", ref.cross_spectrum, [a3, b3], dict(dim=["y", "x"], detrend="constant", window="hann"), ("t", "y", "x"), c3)
    add("cs2d_numpy_like", ref.cross_spectrum, [a3, b3], dict(dim=["y", "x"], true_phase=False, scaling="spectrum"), ("t", "y", "x"), c3)
    add("cphase2d", ref.cross_phase, [a3, b3], dict(dim=["y", "x"], detrend="constant", window="hann"), ("t", "y", "x"), c3)
    add("iso_ps", ref.isotropic_power_spectrum, [iso_in], dict(dim=["y", "x"], detrend="constant", window="hann"), ("t", "y", "x"), ciso)
    add("iso_ps_truncate", ref.isotropic_power_spectrum, [iso_in], dict(dim=["y", "x"], detrend="linear", window="hann", truncate=True), ("t", "y", "x"), ciso)
    add("iso_cs", ref.isotropic_cross_spectrum, [iso_in, iso_in[::-1].copy()], dict(dim=["y", "x"], window="hann"), ("t", "y", "x"), ciso)
    add("detrend_2d", lambda d, **k: ref.detrend(d, ["y", "x"], **k), [a3], dict(detrend_type="linear"), ("t", "y", "x"), c3)
    add("detrend_1d", lambda d, **k: ref.detrend(d, ["x"], **k), [a3], dict(detrend_type="linear"), ("t", "y", "x"), c3)
    add("detrend_3d", lambda d, **k: ref.detrend(d, ["t", "y", "x"], **k), [a3], dict(detrend_type="linear"), ("t", "y", "x"), c3)
    # ---- oracle-only additions (gpu=False)
    add("ps1d_tukey_realdim", ref.power_spectrum, [a1], dict(real_dim="x", detrend="linear", window="tukey"), ("x",), c1, gpu=False)
    add("ps1d_segments", ref.power_spectrum, [a1], dict(dim="x", chunks_to_segments=True, window="hann", detrend="constant"), ("x",), c1,
        gpu=False, chunks={"x": 16})
    add("fft2d_noshift", ref.fft, [a2], dict(shift=False, true_phase=True, detrend="constant"), ("y", "x"), c2, gpu=False)
    add("cphase1d", ref.cross_phase, [a1, a1[::-1].copy()], dict(), ("x",), c1, gpu=False)
    add("cs2d_realdim_corr", ref.cross_spectrum, [a3, b3], dict(dim=["y", "x"], real_dim="x", window="hann", window_correction=True,
                                                                 detrend="linear"), ("t", "y", "x"), c3, gpu=False)
    add("ps3d_all_linear", ref.power_spectrum, [a3], dict(detrend="linear", window="hann", scaling="spectrum"), ("t", "y", "x"), c3, gpu=False)
    add("iso_ps_nfactor2", ref.isotropic_power_spectrum, [iso_in], dict(dim=["y", "x"], nfactor=2, truncate=False, window="hann"),
        ("t", "y", "x"), ciso, gpu=False)
    add("iso_cs_truncate", ref.isotropic_cross_spectrum, [iso_in, iso_in[:, ::-1].copy()], dict(dim=["y", "x"], detrend="linear", window="hann",
                                                                                               truncate=True), ("t", "y", "x"), ciso, gpu=False)

    # round trips through the reference's own ifft
    ft = ref.fft(da_of(a2, ("y", "x"), c2))
    back = ref.ifft(ft)
    store["ifft2d__in0"] = np.asarray(ft.values)
    store["ifft2d__out"] = np.asarray(back.values)
    store["ifft2d__coord__x"] = np.asarray(back["x"].values, dtype=float)
    store["ifft2d__coord__y"] = np.asarray(back["y"].values, dtype=float)
    store["ifft2d__freq_x"] = np.asarray(ft["freq_x"].values)
    store["ifft2d__freq_y"] = np.asarray(ft["freq_y"].values)
    store["ifft2d__lag_x"] = np.asarray(float(ft["freq_x"].attrs["direct_lag"]))
    store["ifft2d__lag_y"] = np.asarray(float(ft["freq_y"].attrs["direct_lag"]))
    ftr = ref.fft(da_of(a2, ("y", "x"), c2), real_dim="x")
    backr = ref.ifft(ftr, real_dim="freq_x")
    store["irfft2d__in0"] = np.asarray(ftr.values)
    store["irfft2d__out"] = np.asarray(backr.values)
    store["irfft2d__freq_x"] = np.asarray(ftr["freq_x"].values)
    store["irfft2d__freq_y"] = np.asarray(ftr["freq_y"].values)
    store["irfft2d__lag_x"] = np.asarray(float(ftr["freq_x"].attrs["direct_lag"]))
    store["irfft2d__lag_y"] = np.asarray(float(ftr["freq_y"].attrs["direct_lag"]))
    # padding
    p = ref.pad(da_of(a2, ("y", "x"), c2), x=(3, 5), y=2)
    store["pad__in0"] = a2
    store["pad__out"] = np.asarray(p.values)
    store["pad__coord__x"] = np.asarray(p["x"].values, dtype=float)
    store["pad__coord__y"] = np.asarray(p["y"].values, dtype=float)
    # unpad of the padded array (xrft/padding.py:321-446): data and coordinates come back
    up = ref.unpad(p, dict(x=(3, 5), y=2))
    store["unpad__out"] = np.asarray(up.values)
    store["unpad__coord__x"] = np.asarray(up["x"].values, dtype=float)
    store["unpad__coord__y"] = np.asarray(up["y"].values, dtype=float)
    store["__cases__"] = np.array([repr(c) for c in cases])
    store["__cases_cpu__"] = np.array([repr(c) for c in cpu_cases])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cases.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
