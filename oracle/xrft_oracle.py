"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A plain numpy/scipy/pandas restatement of the arithmetic of xgcm/xrft
(reference @ efc1c30, /root/reference) on *labelled raw arrays*.  It is the
checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product package ``xrft_b200`` never imports anything from ``oracle/``.

Pinning: the reference ships no golden vectors on disk;
this restatement is pinned (a) by the reference's own known-answer tests
ported in ``tests/test_oracle_known_answers.py`` (periodogram, sine amplitude,
Parseval, sinc, cross-phase, isotropic sum/slope, detrend recovery, padding
linspace answers) and (b) by fixtures produced by executing the UNMODIFIED
reference source under a stand-in ``xarray`` module
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).

A "labelled array" here is the triple ``(data, dims, coords)``:
``data``  numpy ndarray, ``dims`` tuple of str, ``coords`` dict name -> 1-D
numpy array (only dimension coordinates; a missing entry means the default
integer index, like xarray).  Functions return ``Labelled`` objects.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd
import scipy.linalg as spl
import scipy.signal as sps

__all__ = [
    "Labelled",
    "fft",
    "ifft",
    "power_spectrum",
    "cross_spectrum",
    "cross_phase",
    "isotropize",
    "isotropic_power_spectrum",
    "isotropic_cross_spectrum",
    "detrend",
    "apply_window",
    "pad",
    "unpad",
    "fit_loglog",
    "cut_codes",
]


@dataclass
class Labelled:
    data: np.ndarray
    dims: Tuple[str, ...]
    coords: Dict[str, np.ndarray] = field(default_factory=dict)
    coord_attrs: Dict[str, dict] = field(default_factory=dict)
    chunks: Optional[Dict[str, int]] = None  # emulates dask chunk lengths for chunks_to_segments

    def __post_init__(self):
        self.data = np.asarray(self.data)
        self.dims = tuple(self.dims)
        assert self.data.ndim == len(self.dims)
        self.coords = {k: np.asarray(v) for k, v in self.coords.items()}

    def coord(self, d):
        if d in self.coords:
            return self.coords[d]
        return np.arange(self.data.shape[self.dims.index(d)])

    def axis(self, d):
        return self.dims.index(d)

    def transpose(self, *dims):
        perm = [self.dims.index(d) for d in dims]
        return Labelled(
            np.transpose(self.data, perm), tuple(dims), dict(self.coords),
            {k: dict(v) for k, v in self.coord_attrs.items()}, self.chunks,
        )


def _move_to_end(lst, el):  # xrft/xrft.py:287-288
    return [i for i in lst if i != el] + [el]


# ----------------------------------------------------------------------------
# coordinates: xrft/xrft.py:139-234, 269-304
# ----------------------------------------------------------------------------
def _diff_coord(coord):  # xrft/xrft.py:195-212 (cftime branch omitted: cftime absent)
    coord = np.asarray(coord)
    if np.issubdtype(coord.dtype, np.datetime64):
        diff = np.diff(coord).astype("timedelta64[ns]").astype("f8")
        return diff / 1e9
    return np.diff(coord)


def _lag_coord(coord):  # xrft/xrft.py:215-234
    coord = np.asarray(coord)
    if coord[-1] > coord[0]:
        coord_data = coord
    else:
        coord_data = np.flip(coord, axis=-1)
    lag = coord_data[len(coord) // 2]
    if np.issubdtype(coord.dtype, np.datetime64):
        return lag.astype("timedelta64[s]").astype("f8")
    return lag


def _get_coordinate_spacing(coord, spacing_tol, name="?"):  # xrft/xrft.py:291-304
    diff = _diff_coord(coord)
    delta = np.abs(diff[0])
    if not np.allclose(diff, diff[0], rtol=spacing_tol):
        raise ValueError(
            "Can't take Fourier transform because coodinate %s is not evenly spaced" % name
        )
    if delta == 0.0:
        raise ValueError(
            "Can't take Fourier transform because spacing in coordinate %s is zero" % name
        )
    return delta


def _is_valid_fft_coord(coord):  # xrft/xrft.py:269-274
    coord = np.asarray(coord)
    return bool(
        pd.api.types.is_numeric_dtype(coord.dtype)
        or pd.api.types.is_datetime64_any_dtype(coord.dtype)
    )


def _freq(N, delta_x, real, shift):  # xrft/xrft.py:139-155
    if real is None:
        fftfreq = [np.fft.fftfreq] * len(N)
    else:
        fftfreq = [np.fft.fftfreq] * (len(N) - 1)
        fftfreq.append(np.fft.rfftfreq)
    k = [f(Nx, dx) for (f, Nx, dx) in zip(fftfreq, N, delta_x)]
    if shift:
        k = [np.fft.fftshift(l) for l in k]
    return k


def _ifreq(N, delta_x, real, shift):  # xrft/xrft.py:158-175
    if real is None:
        fftfreq = [np.fft.fftfreq] * len(N)
    else:
        irfftfreq = lambda Nx, dx: np.fft.fftfreq(2 * (Nx - 1), dx)
        fftfreq = [np.fft.fftfreq] * (len(N) - 1)
        fftfreq.append(irfftfreq)
    k = [f(Nx, dx) for (f, Nx, dx) in zip(fftfreq, N, delta_x)]
    if shift:
        k = [np.fft.fftshift(l) for l in k]
    return k


def _new_name(d, prefix):  # xrft/xrft.py:186
    return prefix + d if d[: len(prefix)] != prefix else d[len(prefix):]


# ----------------------------------------------------------------------------
# window: xrft/xrft.py:39-103
# ----------------------------------------------------------------------------
_WINDOWS = [
    "hann", "hamming", "kaiser", "tukey", "parzen", "taylor", "boxcar", "barthann",
    "bartlett", "blackman", "blackmanharris", "bohman", "chebwin", "cosine", "dpss",
    "exponential", "flattop", "gaussian", "general_cosine", "general_gaussian",
    "general_hamming", "triang", "nuttall",
]


def apply_window(la: Labelled, dims, window_type="hann"):
    """xrft/xrft.py:39-103 -> (window product as broadcastable ndarray, windowed data)."""
    if window_type is True:
        window_type = "hann"
        warnings.warn("boolean window option will be deprecated", FutureWarning)
    elif window_type not in _WINDOWS:
        raise NotImplementedError("Window type {window_type} not supported.")
    if dims is None:
        dims = list(la.dims)
    elif isinstance(dims, str):
        dims = [dims]
    win_func = getattr(sps.windows, window_type)
    # reduce(operator.mul, windows[::-1]) with name-broadcast == outer product laid on la's axes
    wprod = np.ones([1] * la.data.ndim)
    for d in dims[::-1]:
        w = win_func(la.data.shape[la.axis(d)], sym=False)
        shp = [1] * la.data.ndim
        shp[la.axis(d)] = -1
        wprod = wprod * w.reshape(shp)
    return wprod, la.data * wprod


# ----------------------------------------------------------------------------
# detrend: xrft/detrend.py:11-138
# ----------------------------------------------------------------------------
def _detrend_2d_ufunc(arr):  # xrft/detrend.py:100-113
    assert arr.ndim == 2
    N = arr.shape
    col0 = np.ones(N[0] * N[1])
    col1 = np.repeat(np.arange(N[0]), N[1]) + 1
    col2 = np.tile(np.arange(N[1]), N[0]) + 1
    G = np.stack([col0, col1, col2]).transpose()
    d_obs = np.reshape(arr, (N[0] * N[1], 1))
    m_est = np.dot(np.dot(spl.inv(np.dot(G.T, G)), G.T), d_obs)
    d_est = np.dot(G, m_est)
    linear_fit = np.reshape(d_est, N)
    return arr - linear_fit


def _detrend_3d_ufunc(arr):  # xrft/detrend.py:116-138
    assert arr.ndim == 3
    N0, N1, N2 = arr.shape
    i = np.repeat(np.arange(N0), N1 * N2) + 1
    j = np.tile(np.repeat(np.arange(N1), N2), N0) + 1
    k = np.tile(np.arange(N2), N0 * N1) + 1
    col0 = np.ones(N0 * N1 * N2)
    G = np.stack([col0, i, j, k], axis=1)
    d_obs = arr.reshape(-1, 1)
    m_est, _, _, _ = np.linalg.lstsq(G, d_obs, rcond=None)
    d_est = G @ m_est
    return arr - d_est.reshape(N0, N1, N2)


def _vectorize_core(func, data, axes):
    """xr.apply_ufunc(..., input_core_dims=[dim], output_core_dims=[dim], vectorize=True):
    core dims are moved to the end (in `dim` order), func applied per leading index,
    result has core dims LAST (caller transposes back, xrft/xrft.py:427-428)."""
    moved = np.moveaxis(data, axes, list(range(data.ndim - len(axes), data.ndim)))
    lead = moved.shape[: data.ndim - len(axes)]
    out = np.empty(moved.shape, dtype=data.dtype)  # output_dtypes=[da.dtype]
    for idx in np.ndindex(*lead):
        out[idx] = func(moved[idx])
    return out


def detrend(la: Labelled, dim, detrend_type="constant") -> Labelled:
    """xrft/detrend.py:11-97.  Returned dims order equals the input's."""
    if dim is None:
        dim = list(la.dims)
    elif isinstance(dim, str):
        dim = [dim]
    if detrend_type not in ["constant", "linear", None]:
        raise NotImplementedError("%s is not a valid detrending option." % detrend_type)
    if detrend_type is None:
        return la
    axes = [la.axis(d) for d in dim]
    if detrend_type == "constant":  # detrend.py:55
        out = la.data - la.data.mean(axis=tuple(axes), keepdims=True)
        return Labelled(out, la.dims, dict(la.coords), dict(la.coord_attrs), la.chunks)
    if len(dim) == 1:  # detrend.py:64-71 ; scipy returns float64 for ints, keeps f32/f64
        out = sps.detrend(la.data, axis=axes[0]).astype(la.data.dtype, copy=False)
    elif len(dim) == 2:
        core = _vectorize_core(_detrend_2d_ufunc, la.data, axes)
        out = np.moveaxis(core, list(range(la.data.ndim - 2, la.data.ndim)), axes)
    elif len(dim) == 3:
        core = _vectorize_core(_detrend_3d_ufunc, la.data, axes)
        out = np.moveaxis(core, list(range(la.data.ndim - 3, la.data.ndim)), axes)
    else:
        raise NotImplementedError("Only 1D, 2D, and 3D detrending are implemented so far.")
    return Labelled(out, la.dims, dict(la.coords), dict(la.coord_attrs), la.chunks)


# ----------------------------------------------------------------------------
# segments: xrft/xrft.py:106-136
# ----------------------------------------------------------------------------
def _stack_chunks(la: Labelled, dim, suffix="_segment") -> Labelled:
    newdims, newshape, newcoords = [], [], {}
    for d in la.dims:
        n = la.data.shape[la.axis(d)]
        if d in dim:
            if la.chunks is None or d not in la.chunks:
                chunklen = n
            else:
                chunklen = la.chunks[d]
                if n % chunklen != 0:
                    raise ValueError("Chunk lengths need to be the same.")
            coord_rs = la.coord(d).reshape((int(n / chunklen), int(chunklen)))
            newdims += [d + suffix, d]
            newshape += [int(n / chunklen), int(chunklen)]
            newcoords[d + suffix] = np.arange(int(n / chunklen))
            newcoords[d] = coord_rs[0]
        else:
            newdims.append(d)
            newshape.append(n)
            newcoords[d] = la.coord(d)
    return Labelled(la.data.reshape(newshape), tuple(newdims), newcoords)


# ----------------------------------------------------------------------------
# fft: xrft/xrft.py:307-476
# ----------------------------------------------------------------------------
def fft(
    la: Labelled,
    spacing_tol=1e-3,
    dim=None,
    real_dim=None,
    shift=True,
    detrend=None,
    window=None,
    true_phase=True,
    true_amplitude=True,
    chunks_to_segments=False,
    prefix="freq_",
) -> Labelled:
    detrend_type = detrend
    if dim is None:
        dim = list(la.dims)
    elif isinstance(dim, str):
        dim = [dim]
    dim = list(dim)
    if real_dim is not None:
        if real_dim not in la.dims:
            raise ValueError(
                "The dimension along which real FT is taken must be one of the existing dimensions."
            )
        dim = _move_to_end(dim, real_dim)
    if not np.all([_is_valid_fft_coord(la.coord(d)) for d in dim]):  # :277-281
        raise ValueError("All transformed dimensions coordinates must be numerical or datetime.")
    if chunks_to_segments:
        la = _stack_chunks(la, dim)
    rawdims = la.dims
    if real_dim is not None:
        la = la.transpose(*_move_to_end(list(la.dims), real_dim))
    axis_num = [la.axis(d) for d in dim]
    N = [la.data.shape[n] for n in axis_num]
    delta_x = [_get_coordinate_spacing(la.coord(d), spacing_tol, d) for d in dim]
    lag_x = [_lag_coord(la.coord(d)) for d in dim]

    if detrend_type is not None:  # :425-430
        la = globals()["detrend"](la, dim, detrend_type=detrend_type)
    data = la.data
    if window is not None:  # :432-433
        _, data = apply_window(Labelled(data, la.dims, la.coords), dim, window_type=window)

    if real_dim is None:
        fft_fn = np.fft.fftn
    else:
        shift = False
        fft_fn = np.fft.rfftn

    if true_phase:  # :435-442
        reversed_axis = [la.axis(d) for d in dim if la.coord(d)[-1] < la.coord(d)[0]]
        f = fft_fn(np.fft.ifftshift(np.flip(data, axis=reversed_axis), axes=axis_num), axes=axis_num)
    else:
        f = fft_fn(data, axes=axis_num)
    if shift:
        f = np.fft.fftshift(f, axes=axis_num)

    k = _freq(N, delta_x, real_dim, shift)
    newdims = list(la.dims)
    newcoords = {c: v for c, v in la.coords.items() if c not in dim}
    cattrs = {}
    for d, kk in zip(dim, k):
        nn = _new_name(d, prefix)
        newdims[la.axis(d)] = nn
        newcoords[nn] = kk
        cattrs[nn] = {"spacing": kk[1] - kk[0]}

    if true_phase:  # :462-469
        for d, kk, lag in zip(dim, k, lag_x):
            shp = [1] * f.ndim
            shp[la.axis(d)] = -1
            f = f * np.exp(-1j * 2.0 * np.pi * kk * lag).reshape(shp)
            cattrs[_new_name(d, prefix)]["direct_lag"] = lag
    if true_amplitude:  # :471-472
        f = f * np.prod(delta_x)

    out = Labelled(f, tuple(newdims), newcoords, cattrs)
    swap = {d: _new_name(d, prefix) for d in dim}
    return out.transpose(*[swap.get(d, d) for d in rawdims])


# ----------------------------------------------------------------------------
# ifft: xrft/xrft.py:479-646
# ----------------------------------------------------------------------------
def ifft(
    la: Labelled,
    spacing_tol=1e-3,
    dim=None,
    real_dim=None,
    shift=True,
    true_phase=True,
    true_amplitude=True,
    chunks_to_segments=False,
    prefix="freq_",
    lag=None,
) -> Labelled:
    if dim is None:
        dim = list(la.dims)
    elif isinstance(dim, str):
        dim = [dim]
    dim = list(dim)
    if real_dim is not None:
        if real_dim not in la.dims:
            raise ValueError(
                "The dimension along which real IFT is taken must be one of the existing dimensions."
            )
        dim = _move_to_end(dim, real_dim)
    if not np.all([_is_valid_fft_coord(la.coord(d)) for d in dim]):
        raise ValueError("All transformed dimensions coordinates must be numerical or datetime.")
    if lag is None:  # :557-560
        lag = [la.coord_attrs.get(d, {}).get("direct_lag", 0.0) for d in dim]
        warnings.warn("Default ifft's behaviour (lag=None) changed!", FutureWarning)
    else:
        if isinstance(lag, (float, int)):
            lag = [lag]
        if len(dim) != len(lag):
            raise ValueError("dim and lag must have the same length.")
        if not true_phase:
            warnings.warn("Setting lag with true_phase=False does not guarantee accurate ifft.", Warning)
        lag = [la.coord_attrs.get(d, {}).get("direct_lag") if l is None else l for d, l in zip(dim, lag)]

    data = la.data
    if true_phase:  # :574-576
        for d, l in zip(dim, lag):
            shp = [1] * data.ndim
            shp[la.axis(d)] = -1
            data = data * np.exp(1j * 2.0 * np.pi * la.coord(d) * l).reshape(shp)
    la = Labelled(data, la.dims, dict(la.coords), dict(la.coord_attrs), la.chunks)
    if chunks_to_segments:
        la = _stack_chunks(la, dim)
    rawdims = la.dims
    if real_dim is not None:
        la = la.transpose(*_move_to_end(list(la.dims), real_dim))
    fft_fn = np.fft.ifftn if real_dim is None else np.fft.irfftn
    axis_num = [la.axis(d) for d in dim]
    N = [la.data.shape[n] for n in axis_num]

    # sortby(dim): :598
    data = la.data
    coords = dict(la.coords)
    for d in dim:
        order = np.argsort(la.coord(d), kind="stable")
        data = np.take(data, order, axis=la.axis(d))
        coords[d] = la.coord(d)[order]
    la = Labelled(data, la.dims, coords, dict(la.coord_attrs))
    delta_x = [_get_coordinate_spacing(la.coord(d), spacing_tol, d) for d in dim]
    for d in dim:  # :600-606
        l = _lag_coord(la.coord(d)) if d != real_dim else la.coord(d)[0]
        if np.abs(l) > spacing_tol:
            raise ValueError(
                "Inverse Fourier Transform can not be computed because coordinate %s is not centered on zero frequency" % d
            )
    axis_shift = [la.axis(d) for d in dim if d != real_dim]
    f = np.fft.ifftshift(la.data, axes=axis_shift)  # :612-614
    f = fft_fn(f, axes=axis_num)
    if not true_phase:
        f = np.fft.ifftshift(f, axes=axis_num)
    if shift:
        f = np.fft.fftshift(f, axes=axis_num)
    k = _ifreq(N, delta_x, real_dim, shift)

    newdims = list(la.dims)
    newcoords = {c: v for c, v in la.coords.items() if c not in dim}
    cattrs = {}
    spacings = []
    for d, kk, l in zip(dim, k, lag):
        nn = _new_name(d, prefix)
        newdims[la.axis(d)] = nn
        spacings.append(kk[1] - kk[0])
        newcoords[nn] = kk + l  # :637-639
        cattrs[nn] = {"spacing": kk[1] - kk[0]}
    if true_amplitude:  # :641-642
        f = f / np.prod([float(s) for s in spacings])
    out = Labelled(f, tuple(newdims), newcoords, cattrs)
    swap = {d: _new_name(d, prefix) for d in dim}
    return out.transpose(*[swap.get(d, d) for d in rawdims])


# ----------------------------------------------------------------------------
# spectra: xrft/xrft.py:649-874
# ----------------------------------------------------------------------------
def _window_correction_factor(la, dim, scaling, window):  # :649-660
    if window is None:
        raise ValueError("window_correction can only be applied when windowing is turned on.")
    windows, _ = apply_window(la, dim, window_type=window)
    if scaling == "density":
        return (windows ** 2).mean()
    elif scaling == "spectrum":
        return windows.mean() ** 2
    raise ValueError("Unknown {} scaling flag".format(scaling))


def _psd_scaling_factor(ps: Labelled, dims, scaling):  # :663-670
    fs = np.prod([float(ps.coord_attrs[d]["spacing"]) for d in dims])
    if scaling == "density":
        return fs
    elif scaling == "spectrum":
        return fs ** 2
    raise ValueError("Unknown {} scaling flag".format(scaling))


def _psd_real_dim_scaling(la, ps: Labelled, real_dim, updated_dims):  # :673-682
    real = next(d for d in updated_dims if d.endswith(real_dim))
    f = np.full(ps.data.shape[ps.axis(real)], 2.0)
    if la.data.shape[la.axis(real_dim)] % 2 == 0:
        f[0], f[-1] = 1.0, 1.0
    else:
        f[0] = 1.0
    shp = [1] * ps.data.ndim
    shp[ps.axis(real)] = -1
    return f.reshape(shp)


def power_spectrum(la, dim=None, real_dim=None, scaling="density", window_correction=False, **kwargs):
    """xrft/xrft.py:685-750"""
    if "density" in kwargs:
        density = kwargs.pop("density")
        warnings.warn("density flag will be deprecated", FutureWarning)
        scaling = "density" if density else "false_density"
    kwargs.update({"true_amplitude": True, "true_phase": False})
    daft = fft(la, dim=dim, real_dim=real_dim, **kwargs)
    updated_dims = [d for d in daft.dims if (d not in la.dims and "segment" not in d)]
    ps = Labelled(np.abs(daft.data) ** 2, daft.dims, daft.coords, daft.coord_attrs)
    if real_dim is not None:
        ps.data = ps.data * _psd_real_dim_scaling(la, ps, real_dim, updated_dims)
    if scaling != "false_density":
        if window_correction:
            ps.data = ps.data / _window_correction_factor(la, dim, scaling, kwargs.get("window"))
        ps.data = ps.data * _psd_scaling_factor(ps, updated_dims, scaling)
    return ps


def cross_spectrum(la1, la2, dim=None, real_dim=None, scaling="density", window_correction=False,
                   true_phase=True, **kwargs):
    """xrft/xrft.py:753-835"""
    if "density" in kwargs:
        density = kwargs.pop("density")
        warnings.warn("density flag will be deprecated", FutureWarning)
        scaling = "density" if density else "false_density"
    kwargs.update({"true_amplitude": True})
    daft1 = fft(la1, dim=dim, real_dim=real_dim, true_phase=true_phase, **kwargs)
    daft2 = fft(la2, dim=dim, real_dim=real_dim, true_phase=true_phase, **kwargs)
    if daft1.dims != daft2.dims:
        raise ValueError("The two datasets have different dimensions")
    updated_dims = [d for d in daft1.dims if (d not in la1.dims and "segment" not in d)]
    cs = Labelled(daft1.data * np.conj(daft2.data), daft1.dims, daft1.coords, daft1.coord_attrs)
    if real_dim is not None:
        cs.data = cs.data * _psd_real_dim_scaling(la1, cs, real_dim, updated_dims)
    if scaling != "false_density":
        if window_correction:
            cs.data = cs.data / _window_correction_factor(la1, dim, scaling, kwargs.get("window"))
        cs.data = cs.data * _psd_scaling_factor(cs, updated_dims, scaling)
    return cs


def cross_phase(la1, la2, dim=None, true_phase=True, **kwargs):
    """xrft/xrft.py:838-874"""
    cs = cross_spectrum(la1, la2, dim=dim, true_phase=true_phase, **kwargs)
    return Labelled(np.angle(cs.data), cs.dims, cs.coords, cs.coord_attrs)


# ----------------------------------------------------------------------------
# isotropic binning: xrft/xrft.py:877-1187
# ----------------------------------------------------------------------------
def cut_codes(values, nbins):
    """Integer bin codes of ``pd.cut(np.ravel(values), nbins)`` (xrft/xrft.py:921).

    Restatement of pandas ``_nbins_to_bins`` / ``_bins_to_cuts`` for right=True:
    equal-width edges over [min, max], first edge lowered by 0.1% of the range,
    code = searchsorted(edges, v, side='left') - 1.  Checked against pandas itself
    in tests/test_oracle_known_answers.py::test_cut_codes_match_pandas.
    """
    v = np.ravel(values)
    mn, mx = v.min(), v.max()
    if mn == mx:
        mn -= 0.001 * abs(mn) if mn != 0 else 0.001
        mx += 0.001 * abs(mx) if mx != 0 else 0.001
        edges = np.linspace(mn, mx, nbins + 1, endpoint=True)
    else:
        edges = np.linspace(mn, mx, nbins + 1, endpoint=True)
        adj = (mx - mn) * 0.001
        edges[0] -= adj
    codes = edges.searchsorted(v, side="left") - 1
    return codes.astype(np.int64), edges


def _binned_agg(array, indices, num_bins, func):  # xrft/xrft.py:877-907 (numpy_groupies == bincount)
    idx = np.ravel(indices)
    nd = indices.ndim
    lead = array.shape[: array.ndim - nd]
    flat = array.reshape(lead + (-1,))
    out = np.zeros(lead + (num_bins,), dtype=array.dtype if func == "sum" else float)
    counts = np.bincount(idx, minlength=num_bins)
    for i in np.ndindex(*lead):
        if np.iscomplexobj(flat):
            s = np.bincount(idx, weights=flat[i].real, minlength=num_bins) + 1j * np.bincount(
                idx, weights=flat[i].imag, minlength=num_bins
            )
        else:
            s = np.bincount(idx, weights=flat[i], minlength=num_bins)
        if func == "mean":
            s = s / np.where(counts == 0, 1, counts)
        out[i] = s
    return out


def isotropize(ps: Labelled, fftdim, nfactor=4, truncate=True, complx=False) -> Labelled:
    """xrft/xrft.py:948-1010.  Output dims = non-fft dims + ('freq_r',)."""
    k = ps.coord(fftdim[1])
    l = ps.coord(fftdim[0])
    N = [k.size, l.size]
    nbins = int(min(N) / nfactor)
    # freq_r dims order = (fftdim[1], fftdim[0]) (:980): k varies along axis 0
    freq_r = np.sqrt(k[:, None] ** 2 + l[None, :] ** 2)
    codes, _ = cut_codes(freq_r, nbins)
    codes = codes.reshape(freq_r.shape)
    kr = _binned_agg(freq_r, codes, nbins, "mean")
    if truncate:
        kmax = l.max() if k.max() > l.max() else k.max()
        kr = np.where(kr <= kmax, kr, np.nan)
    else:
        warnings.warn("Isotropic wavenumber larger than the Nyquist wavenumber may result.", FutureWarning)
    # move (fftdim[1], fftdim[0]) to the end, as apply_ufunc does with input_core_dims
    others = [d for d in ps.dims if d not in fftdim]
    pst = ps.transpose(*(others + [fftdim[1], fftdim[0]]))
    data = pst.data
    if complx:
        data = data.astype(np.complex128)
    iso = _binned_agg(data, codes, nbins, "sum")
    coords = {c: v for c, v in ps.coords.items() if c in others}
    coords["freq_r"] = kr
    return Labelled(iso, tuple(others) + ("freq_r",), coords)


def isotropic_power_spectrum(la, spacing_tol=1e-3, dim=None, shift=True, detrend=None, scaling="density",
                             window=None, window_correction=False, nfactor=4, truncate=False, **kwargs):
    """xrft/xrft.py:1013-1095"""
    if "density" in kwargs:
        density = kwargs.pop("density")
        scaling = "density" if density else "false_density"
    if dim is None:
        dim = la.dims
    if len(dim) != 2:
        raise ValueError("The Fourier transform should be two dimensional")
    ps = power_spectrum(la, spacing_tol=spacing_tol, dim=dim, shift=shift, detrend=detrend, scaling=scaling,
                        window_correction=window_correction, window=window, **kwargs)
    fftdim = ["freq_" + d for d in dim]
    return isotropize(ps, fftdim, nfactor=nfactor, truncate=truncate)


def isotropic_cross_spectrum(la1, la2, spacing_tol=1e-3, dim=None, shift=True, detrend=None,
                             scaling="density", window=None, window_correction=False, nfactor=4,
                             truncate=False, **kwargs):
    """xrft/xrft.py:1098-1187"""
    if "density" in kwargs:
        density = kwargs.pop("density")
        scaling = "density" if density else "false_density"
    if dim is None:
        dim = la1.dims
        if dim != la2.dims:
            raise ValueError("The two datasets have different dimensions")
    if len(dim) != 2:
        raise ValueError("The Fourier transform should be two dimensional")
    cs = cross_spectrum(la1, la2, spacing_tol=spacing_tol, dim=dim, shift=shift, detrend=detrend,
                        scaling=scaling, window_correction=window_correction, window=window, **kwargs)
    fftdim = ["freq_" + d for d in dim]
    return isotropize(cs, fftdim, nfactor=nfactor, truncate=truncate, complx=True)


def fit_loglog(x, y):  # xrft/xrft.py:1190-1214
    p = np.polyfit(np.log2(x), np.log2(y), 1)
    y_fit = 2 ** (np.log2(x) * p[0] + p[1])
    return y_fit, p[0], p[1]


# ----------------------------------------------------------------------------
# padding: xrft/padding.py:11-446
# ----------------------------------------------------------------------------
def _pad_coord(coord, width):  # xrft/padding.py:263-318
    coord = np.asarray(coord)
    diff = np.diff(coord)
    if not np.allclose(diff, diff[0]):  # xrft/utils.py:11-20
        raise ValueError("Found unevenly spaced coordinates. These coordinates should be evenly spaced.")
    spacing = diff[0]
    n_start, n_end = (width, width) if isinstance(width, (int, np.integer)) else width
    out = np.pad(coord, (n_start, n_end))
    vmin, vmax = coord[0], coord[-1]
    out[:n_start] = vmin - n_start * spacing + np.linspace(0, spacing * (n_start - 1), n_start)
    out[len(out) - n_end:] = vmax + spacing + np.linspace(0, spacing * (n_end - 1), n_end)
    return out


def pad(la: Labelled, pad_width: dict, mode="constant", constant_values=0, **np_pad_kwargs) -> Labelled:
    """xrft/padding.py:157-181"""
    widths = []
    for d in la.dims:
        w = pad_width.get(d, 0)
        widths.append((w, w) if isinstance(w, (int, np.integer)) else tuple(w))
    if mode == "constant":
        data = np.pad(la.data, widths, mode=mode, constant_values=constant_values)
    else:
        data = np.pad(la.data, widths, mode=mode, **np_pad_kwargs)
    coords = dict(la.coords)
    cattrs = {k: dict(v) for k, v in la.coord_attrs.items()}
    for d, w in pad_width.items():
        coords[d] = _pad_coord(la.coord(d), w)
        cattrs.setdefault(d, {})["pad_width"] = w
    return Labelled(data, la.dims, coords, cattrs)


def unpad(la: Labelled, pad_width: Optional[dict] = None) -> Labelled:
    """xrft/padding.py:394-446"""
    if pad_width is None:
        pad_width = {d: a["pad_width"] for d, a in la.coord_attrs.items() if "pad_width" in a}
        if not pad_width:
            raise ValueError("The passed array doesn't seem to be a padded one")
    sl = [slice(None)] * la.data.ndim
    coords = dict(la.coords)
    cattrs = {k: dict(v) for k, v in la.coord_attrs.items()}
    for d, w in pad_width.items():
        if isinstance(w, (int, np.integer)):
            w = (w, w)
        n = la.data.shape[la.axis(d)]
        s = slice(w[0], n - w[1])
        sl[la.axis(d)] = s
        coords[d] = la.coord(d)[s]
        cattrs.get(d, {}).pop("pad_width", None)
    return Labelled(la.data[tuple(sl)], la.dims, coords, cattrs)
