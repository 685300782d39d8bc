"""xrft-facing API (drop-in for the reference's public functions, xrft/__init__.py:1-8).

Host side = the coordinate / metadata bookkeeping of xrft/xrft.py (dims, frequency
coordinates, ``spacing`` / ``direct_lag`` attrs, error behaviour, FutureWarnings) in plain
Python.  Numeric side = xrft_b200.backend (hand-written sm_100a kernels behind the C-ABI);
every array operation the reference performs with numpy/scipy/dask between "coords are
validated" and "result is labelled" is expressed as per-axis vectors + scalars handed to the
device kernels:

  detrend (detrend.py)          -> moments reduce + fp64 plane subtract, fused into the FFT loads
  window  (xrft.py:39-103)      -> per-axis fp64 scipy window vectors, fused into the FFT loads
  ifftshift of the input, phase ramp exp(-i 2 pi k lag) (xrft.py:435-442, 462-469)
                                -> one complex ramp vector per axis, applied in the FFT epilogue
                                   (FFT(ifftshift x)[k] = X[k] exp(+2 pi i k (N//2)/N))
  fftshift, x prod(dx), |F|^2, F conj(G), angle, psd scalings (xrft.py:446-472, 740-748, 825-833, 865)
                                -> index remap + one scalar + mode in the FFT epilogue
  radial binning (xrft.py:877-1010) -> host-built int32 bin LUT (pandas.cut restated) + device bin sum
"""
from __future__ import annotations

import warnings
from collections import OrderedDict
from typing import List, Optional, Sequence

import numbers
import os
import numpy as np
import scipy.signal as sps

from .dataarray import DataArray, Coordinates, LazyPad, LazyIrfft2, LazySegSpectrum, from_any, either_dict_or_kwargs, _is_torch
from . import _lib as L

__all__ = [
    "fft", "ifft", "dft", "idft", "power_spectrum", "cross_spectrum", "cross_phase", "cross_spectrum_and_phase", "isotropize",
    "isotropic_power_spectrum", "isotropic_cross_spectrum", "fit_loglog", "detrend", "pad", "unpad",
]

_WINDOWS = [
    "hann", "hamming", "kaiser", "tukey", "parzen", "taylor", "boxcar", "barthann", "bartlett", "blackman",
    "blackmanharris", "bohman", "chebwin", "cosine", "dpss", "exponential", "flattop", "gaussian", "general_cosine",
    "general_gaussian", "general_hamming", "triang", "nuttall",
]
_real_flag_warning = "`real` flag will be deprecated in future version of xrft.fft and replaced by `real_dim` flag."


# =============================================================================================
# coordinate helpers (xrft/xrft.py:139-234, 269-304) -- pure host numpy
# =============================================================================================
def _freq(N, delta_x, real, shift):  # xrft.py:139-155
    if real is None:
        fns = [np.fft.fftfreq] * len(N)
    else:
        fns = [np.fft.fftfreq] * (len(N) - 1) + [np.fft.rfftfreq]
    k = [f(Nx, dx) for f, Nx, dx in zip(fns, N, delta_x)]
    if shift:
        k = [np.fft.fftshift(l) for l in k]
    return k


def _ifreq(N, delta_x, real, shift):  # xrft.py:158-175
    if real is None:
        fns = [np.fft.fftfreq] * len(N)
    else:
        fns = [np.fft.fftfreq] * (len(N) - 1) + [lambda Nx, dx: np.fft.fftfreq(2 * (Nx - 1), dx)]
    k = [f(Nx, dx) for f, Nx, dx in zip(fns, N, delta_x)]
    if shift:
        k = [np.fft.fftshift(l) for l in k]
    return k


def _new_name(d, prefix):  # xrft.py:186
    return prefix + d if d[: len(prefix)] != prefix else d[len(prefix):]


def _diff_coord(coord):  # xrft.py:195-212
    v = coord.values
    v0 = v[0] if v.ndim else v
    calendar = getattr(v0, "calendar", None)
    if calendar:
        import cftime  # (not installed in this image: tests/test_host_logic.py drives this branch through a stand-in)

        decoded = cftime.date2num(v, "seconds since 1800-01-01 00:00:00", calendar)
        return np.diff(decoded)
    if np.issubdtype(v.dtype, np.datetime64):
        return np.diff(v).astype("timedelta64[ns]").astype("f8") / 1e9
    return np.diff(v)


def _lag_coord(coord):  # xrft.py:215-234
    v = coord.values
    v0 = v[0]
    calendar = getattr(v0, "calendar", None)
    data = v if v[-1] > v[0] else np.flip(v, axis=-1)
    lag = data[len(v) // 2]
    if calendar:
        import cftime

        return cftime.date2num(lag, "seconds since 1800-01-01 00:00:00", calendar)
    if np.issubdtype(v.dtype, np.datetime64):
        return lag.astype("timedelta64[s]").astype("f8")
    return lag


def _is_valid_fft_coord(coord):  # xrft.py:269-274
    v = coord.values
    if np.issubdtype(v.dtype, np.number) or np.issubdtype(v.dtype, np.datetime64) or v.dtype == bool:
        return True
    try:
        return bool(getattr(v.ravel()[0], "calendar", False))
    except Exception:
        return False


def _check_valid_fft_coords(da, dim):  # xrft.py:277-281
    if not np.all([_is_valid_fft_coord(da[d]) for d in dim]):
        raise ValueError("All transformed dimensions coordinates must be numerical or datetime.")


def _get_coordinate_spacing(coord, spacing_tol):  # xrft.py:291-304
    diff = _diff_coord(coord)
    delta = np.abs(diff[0])
    if not np.allclose(diff, diff[0], rtol=spacing_tol):
        raise ValueError("Can't take Fourier transform because coodinate %s is not evenly spaced" % coord.name)
    if delta == 0.0:
        raise ValueError("Can't take Fourier transform because spacing in coordinate %s is zero" % coord.name)
    return delta


def move_to_end(lst, el):  # xrft.py:287-288
    return [i for i in lst if i != el] + [el]


def _stack_chunks(da, dim, suffix="_segment"):  # xrft.py:106-136
    chunks = da.chunks
    data = da.data
    newdims, newcoords, newshape = [], OrderedDict(), []
    for d in da.dims:
        n = da.sizes[d]
        if d in dim:
            axis_num = da.get_axis_num(d)
            ch = chunks[axis_num] if chunks else (n,)
            if np.diff(ch).sum() != 0:
                raise ValueError("Chunk lengths need to be the same.")
            chunklen = ch[0]
            coord_rs = da[d].values.reshape((int(n / chunklen), int(chunklen)))
            newdims += [d + suffix, d]
            newshape += [int(n / chunklen), int(chunklen)]
            newcoords[d + suffix] = np.arange(int(n / chunklen))
            newcoords[d] = coord_rs[0]
        else:
            newdims.append(d)
            newshape.append(n)
            newcoords[d] = da[d].values
    return DataArray(data.reshape(newshape), dims=newdims, coords=newcoords, attrs=da.attrs)


# =============================================================================================
# device plumbing
# =============================================================================================
def _torch():
    import torch
    return torch


def _device_tensor(arr, device=None):
    """numpy / torch -> contiguous CUDA tensor of a supported dtype (numpy promotion rules for ints)."""
    from . import backend as B

    B.require_cuda()
    torch = _torch()
    if _is_torch(arr):
        t = arr
    else:
        a = np.asarray(arr)
        if a.dtype.kind in "iub":
            a = a.astype(np.float64)
        elif a.dtype == np.float16:
            a = a.astype(np.float32)
        elif a.dtype.kind == "f" and a.dtype.itemsize > 8:
            a = a.astype(np.float64)
        elif a.dtype.kind == "c" and a.dtype.itemsize > 16:
            a = a.astype(np.complex128)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype in (torch.int8, torch.int16, torch.int32, torch.int64, torch.uint8, torch.bool):
        t = t.to(torch.float64)
    if t.dtype in (torch.float16, torch.bfloat16):
        t = t.to(torch.float32)
    if not t.is_cuda:
        t = t.cuda() if device is None else t.to(device)
    return t


def _window_vectors(N, window):
    """One periodic scipy window per transformed dim (xrft.py:83-101), fp64 on the host."""
    if window is True:
        window = "hann"
        warnings.warn(
            "Please provide the name of window adhering to scipy.signal.windows. The boolean option will be deprecated in future releases.",
            FutureWarning,
        )
    elif window not in _WINDOWS:
        raise NotImplementedError(
            "Window type {window_type} not supported. Please adhere to scipy.signal.windows for naming convention."
        )
    fn = getattr(sps.windows, window)
    return [np.asarray(fn(int(n), sym=False), dtype=np.float64) for n in N]


def _is_pow2(n):
    return n >= 1 and (n & (n - 1)) == 0


def _spectral_core(x1, x2, ntrans, mode, *, detrend=None, windows=None, keep_half=False, shift=None, ramps=None,
                   weight=None, scale=1.0, lut=None, nbins=0, with_phase=False, pad2=None, seg_axis=None):
    """Transform the last `ntrans` axes of device tensor(s) x1 (,x2) and apply the epilogue.

    Picks the fused 2-D real kernel chain when it applies, otherwise composes
    detrend_window -> (r)fftn -> spectral_post (all CUDA, same C-ABI library).
    """
    from . import backend as B

    torch = _torch()
    det = {None: 0, "constant": 1, "linear": 2}[detrend]
    shift = list(shift) if shift is not None else [False] * ntrans
    ramps = list(ramps) if ramps is not None else [None] * ntrans
    wins = list(windows) if windows is not None else [None] * ntrans
    two = x2 is not None
    is_real = not x1.is_complex()
    tt = lambda v: torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v
    bins_mode = mode in (L.EPI_BINS_POWER, L.EPI_BINS_CROSS)
    # ---- real_dim transform without detrend / window whose input is a deferred zero padding, or whose size the fused spectrum
    #      chain below does not cover: pad predicate + R2C + ramps + scale in one chain of passes (xrftb_fft2r)
    if (is_real and not two and ntrans == 2 and keep_half and mode == L.EPI_COMPLEX and det == 0 and all(w is None for w in wins)
            and weight is None and not any(shift)):
        pny = x1.shape[-2] + (sum(pad2[0]) if pad2 else 0)
        pnx = x1.shape[-1] + (sum(pad2[1]) if pad2 else 0)
        if (pad2 is not None or not B.spectrum2d_supported(pny, pnx, x1.dtype, False)) and B.fft2r_supported(pny, pnx, x1.dtype):
            return B.fft2r_forward(x1, pad2, tt(ramps[0]), tt(ramps[1]), scale)
    if pad2 is not None:   # no fused chain for this shape: carry the padding out
        x1 = B.pad(x1, [(0, 0)] * (x1.ndim - 2) + [pad2[0], pad2[1]], "constant", 0)
    shape = x1.shape

    # ---- fused path: real 2-D, power-of-two sizes
    if (is_real and ntrans == 2 and B.spectrum2d_supported(shape[-2], shape[-1], x1.dtype, two) and seg_axis is None
            and (not bins_mode or nbins <= 1024) and not (keep_half and shift[1])):
        return B.spectrum2d(
            x1, x2, mode, detrend=det, win_y=tt(wins[0]), win_x=tt(wins[1]), keep_half=keep_half, shift_y=shift[0],
            shift_x=shift[1], scale=scale, ramp_y=tt(ramps[0]), ramp_x=tt(ramps[1]), weight_x=tt(weight), lut=lut, nbins=nbins,
            with_phase=with_phase,
        )

    # ---- composed path
    def prep(x):
        if det == 0 and all(w is None for w in wins):
            return x
        if x.is_complex():
            re = B.detrend_window(x.real.contiguous(), ntrans, det, [tt(w) for w in wins])
            im = B.detrend_window(x.imag.contiguous(), ntrans, det, [tt(w) for w in wins])
            return torch.complex(re, im)
        return B.detrend_window(x, ntrans, det, [tt(w) for w in wins])

    axes = list(range(x1.ndim - ntrans, x1.ndim))
    fs = []
    for x in (x1, x2) if two else (x1,):
        x = prep(x)
        fs.append(B.rfftn(x, axes) if is_real else B.fftn(x, axes))
    post_mode = {L.EPI_BINS_POWER: L.EPI_POWER, L.EPI_BINS_CROSS: L.EPI_CROSS}.get(mode, mode)
    out = B.spectral_post(fs[0], fs[1] if two else None, post_mode, ntrans, shape[-1], hermitian=is_real,
                          keep_half=keep_half, shift=shift, ramps=[tt(r) for r in ramps], weight=tt(weight), scale=scale,
                          seg_axis=seg_axis)
    if bins_mode:
        return B.binned_sum(out, lut, nbins, 2)
    if with_phase:   # composed path: the phase is a second epilogue over the same two transforms
        ph = B.spectral_post(fs[0], fs[1], L.EPI_PHASE, ntrans, shape[-1], hermitian=is_real, keep_half=keep_half, shift=shift,
                             ramps=[tt(r) for r in ramps], weight=None, scale=1.0)
        return out, ph
    return out


def _to_last(t, axes):
    """Permute so that `axes` (in the given order) are the trailing axes; return tensor + inverse perm."""
    nd = t.ndim
    lead = [a for a in range(nd) if a not in axes]
    perm = lead + list(axes)
    inv = [0] * nd
    for i, p in enumerate(perm):
        inv[p] = i
    if perm == list(range(nd)):
        return t.contiguous(), None
    from . import backend as B
    return B.permute_flip(t, perm), inv   # the reference's da.transpose (xrft.py:386-396): one CUDA gather, no eager torch


# =============================================================================================
# fft  (xrft/xrft.py:307-476)
# =============================================================================================
def _fft_prepare(da, spacing_tol, dim, real_dim, shift, true_phase, chunks_to_segments, prefix, real):
    """All host-side bookkeeping of xrft.fft up to (not including) the numerics."""
    if isinstance(spacing_tol, bool) or not isinstance(spacing_tol, numbers.Real):  # xrft.py:372-373 (any real number works there)
        raise TypeError("Please provide a float argument")
    if dim is None:
        dim = list(da.dims)
    elif isinstance(dim, str):
        dim = [dim]
    dim = list(dim)
    if real is not None:
        real_dim = real
        warnings.warn(_real_flag_warning, FutureWarning)
    if real_dim is not None:
        if real_dim not in da.dims:
            raise ValueError("The dimension along which real FT is taken must be one of the existing dimensions.")
        dim = move_to_end(dim, real_dim)
    _check_valid_fft_coords(da, dim)
    if chunks_to_segments:
        da = _stack_chunks(da, dim)
    rawdims = da.dims
    if real_dim is not None:
        da = da.transpose(*move_to_end(list(da.dims), real_dim))
        shift = False
    axis_num = [da.get_axis_num(d) for d in dim]
    N = [da.shape[n] for n in axis_num]
    for d in dim:  # xrft.py:412-420
        bad_coords = [cname for cname in da.coords if cname != d and d in da[cname].dims]
        if bad_coords:
            raise ValueError(
                f"The input array contains coordinate variable(s) ({bad_coords}) whose dims include the transform dimension(s) `{d}`. "
                f"Please drop these coordinates (`.drop({bad_coords}`) before invoking xrft."
            )
    delta_x = [_get_coordinate_spacing(da[d], spacing_tol) for d in dim]
    lag_x = [_lag_coord(da[d]) for d in dim]
    reversed_dims = [d for d in dim if da[d].values[-1] < da[d].values[0]] if true_phase else []
    k = _freq(N, delta_x, real_dim, shift)
    k_unshifted = _freq(N, delta_x, real_dim, False)
    return dict(da=da, dim=dim, rawdims=rawdims, axis_num=axis_num, N=N, delta_x=delta_x, lag_x=lag_x, shift=shift,
                real_dim=real_dim, reversed_dims=reversed_dims, k=k, k_unshifted=k_unshifted, prefix=prefix,
                true_phase=true_phase)


def _phase_ramps(P):
    """Per-axis complex vectors (unshifted index) = ifftshift-of-input phase x exp(-i 2 pi k lag)  (xrft.py:435-442, 462-469)."""
    ramps = []
    for N, ku, lag in zip(P["N"], P["k_unshifted"], P["lag_x"]):
        idx = np.arange(len(ku))
        r = np.exp(2j * np.pi * idx * (N // 2) / N) * np.exp(-1j * 2.0 * np.pi * ku * float(lag))
        ramps.append(r.astype(np.complex128))
    return ramps


def _label_output(P, data, new_last_sizes=None, extra_attrs=True):
    """Build the output DataArray: swap dims to freq_*, keep non-transformed coords (xrft.py:449-476)."""
    da, dim, prefix = P["da"], P["dim"], P["prefix"]
    swap = {d: _new_name(d, prefix) for d in dim}
    newdims = [swap.get(d, d) for d in da.dims]
    coords = Coordinates()
    for cname, c in da.coords.items():
        if cname in dim:
            continue
        OrderedDict.__setitem__(coords, cname, c)
    out = DataArray(data, dims=newdims)
    for cname, c in coords.items():
        OrderedDict.__setitem__(out._coords, cname, c)
    for d, kk, lag in zip(dim, P["k"], P["lag_x"]):
        nn = swap[d]
        attrs = {"spacing": kk[1] - kk[0]}
        if P["true_phase"] and extra_attrs:
            attrs["direct_lag"] = lag
        OrderedDict.__setitem__(out._coords, nn, DataArray(kk, dims=(nn,), name=nn, attrs=attrs, _coord=True))
    return out.transpose(*[swap.get(d, d) for d in P["rawdims"]])


_STREAM_MIN_BYTES = 256 << 20    # host inputs at least this large are streamed chunk-wise (H2D | compute | D2H overlap)
_STREAM_CHUNK_BYTES = 128 << 20  # target input bytes per streamed chunk (PCIe-bound: small chunks shorten the pipeline fill/drain)


def _stream_host_chunks(arrs, ntrans, out_host, core):
    """Out-of-core style execution for HOST (numpy) inputs whose transform axes are trailing: the leading (batch)
    axis is cut into chunks that flow through three CUDA streams -- host->device copy of chunk i+1, kernels of
    chunk i and device->host copy of chunk i-1 overlap (the role dask's chunk iteration plays in the reference).
    `core(x1, x2)` runs the device numerics for one chunk; the result lands in `out_host` (pinned if possible)."""
    torch = _torch()
    host = [torch.from_numpy(np.ascontiguousarray(a)) for a in arrs]
    lead = host[0].shape[0]
    per_item = host[0][0].numel() * host[0].element_size()
    step = max(1, min(lead, _STREAM_CHUNK_BYTES // max(per_item, 1)))
    dev = torch.device("cuda", torch.cuda.current_device())
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    cur = torch.cuda.current_stream()
    for st in (s_in, s_cmp, s_out):
        st.wait_stream(cur)
    bufs = [[torch.empty((step,) + tuple(h.shape[1:]), dtype=h.dtype, device=dev) for h in host] for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]   # input buffer b consumed by the kernels
    keep = []
    los = list(range(0, lead, step))

    def upload(i):
        # host -> device copy of chunk i into buffer i % 2 (free once the kernels of chunk i-2 have consumed it)
        lo, b = los[i], i % 2
        hi = min(lead, lo + step)
        with torch.cuda.stream(s_in):
            if i >= 2:
                s_in.wait_event(ev_free[b])
            for h, d in zip(host, bufs[b]):
                d[: hi - lo].copy_(h[lo:hi], non_blocking=True)
            ev_in[b].record(s_in)

    for i, lo in enumerate(los):
        hi = min(lead, lo + step)
        b = i % 2
        # (queueing the upload of chunk i+1 ahead of chunk i's kernels was measured too: 0.52-0.82 of the link probe against
        # 0.78-0.95 in this order on a two-GPU box, profiles/README.md)
        upload(i)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[b])
            xs = [d[: hi - lo] for d in bufs[b]]
            o = core(xs[0], xs[1] if len(xs) == 2 else None)
            ev_free[b].record(s_cmp)
            done = torch.cuda.Event()
            done.record(s_cmp)
        if out_host is None:
            full = (lead,) + tuple(o.shape[1:])
            try:
                out_host = torch.empty(full, dtype=o.dtype, pin_memory=True)
            except Exception:  # pragma: no cover - pinned allocation can fail on small hosts
                out_host = torch.empty(full, dtype=o.dtype)
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            o.record_stream(s_out)
            out_host[lo:hi].copy_(o, non_blocking=True)
        keep.append(o)
        if len(keep) > 3:
            keep.pop(0)
    for st in (s_in, s_cmp, s_out):
        cur.wait_stream(st)
    torch.cuda.current_stream().synchronize()
    return out_host


def _check_out(out, shape, np_dtype):
    """`out=` (extension): a preallocated C-contiguous host buffer of exactly the result's shape and dtype"""
    o = out.numpy() if _is_torch(out) else out
    if _is_torch(out) and out.is_cuda:
        raise ValueError("out= must be a host buffer (numpy array or CPU torch tensor)")
    if not isinstance(o, np.ndarray) or tuple(o.shape) != tuple(shape) or o.dtype != np_dtype or not o.flags["C_CONTIGUOUS"]:
        raise ValueError("out= must be a C-contiguous host array of shape %s and dtype %s" % (tuple(shape), np.dtype(np_dtype)))


def _run_forward(P, das, mode, detrend, window, scale, ramps=None, weight=None, lut=None, nbins=0, out=None, plans=None,
                 with_phase=False, seg_dim=None):
    """Numerics of fft/power/cross for the prepared plan P on one or two DataArrays (already stacked/transposed).
    Container convention: host (numpy) inputs give a numpy result, device (torch CUDA) inputs a torch CUDA result."""
    torch = _torch()
    dim, real_dim = P["dim"], P["real_dim"]
    ntrans = len(dim)
    if ntrans > 3:
        raise NotImplementedError("transforms over more than 3 dimensions are not supported")
    if detrend not in (None, "constant", "linear"):
        raise NotImplementedError("%s is not a valid detrending option. Valid options are: 'constant','linear', or None." % detrend)
    wins = _window_vectors(P["N"], window) if window is not None else None
    # ---- host inputs: stream chunks of the leading axis (numpy in -> numpy out, like the reference)
    nd = P["da"].ndim
    host_in = all(d.lazy_pad is None and not _is_torch(d.data) for d in das)   # (a deferred zero padding is device-resident)
    trailing = list(P["axis_num"]) == list(range(nd - ntrans, nd))
    plans = plans or [P] * len(das)
    any_reversed = any(pl["reversed_dims"] for pl in plans)
    if out is not None:
        W = P["N"][-1] // 2 + 1 if real_dim is not None else P["N"][-1]
        lead_shape = [s_ for a, s_ in enumerate(P["da"].shape) if a not in P["axis_num"]]
        cplx = mode in (L.EPI_COMPLEX, L.EPI_CROSS)
        f32 = all(str(d.data.dtype).endswith("float32") for d in das)
        _check_out(out, lead_shape + list(P["N"][:-1]) + [W], (np.complex64 if cplx else np.float32) if f32 else (np.complex128 if cplx else np.float64))
        if not (host_in and trailing):
            raise ValueError("out= is only supported for host inputs whose transform axes are trailing")
    if (host_in and trailing and nd > ntrans and not any_reversed and not with_phase
            and das[0].data.dtype in (np.float32, np.float64) and all(d.data.dtype == das[0].data.dtype for d in das)
            and das[0].data.nbytes >= _STREAM_MIN_BYTES and das[0].shape[0] >= 2):
        from . import backend as B
        B.require_cuda()
        keep_half = real_dim is not None
        shifts = [P["shift"]] * ntrans if real_dim is None else [False] * ntrans

        # the per-axis vectors go to the device ONCE: a pageable copy inside the chunk loop would block the host until the
        # chunk's own upload has finished, and the copy engine would idle while the host catches up
        dev = torch.device("cuda", torch.cuda.current_device())
        f32 = das[0].data.dtype == np.float32
        rdt, cdt = (torch.float32, torch.complex64) if f32 else (torch.float64, torch.complex128)
        up = lambda v, dt: None if v is None else (v if _is_torch(v) else torch.from_numpy(np.ascontiguousarray(v))).to(device=dev, dtype=dt)
        wins_d = None if wins is None else [up(w, rdt) for w in wins]
        ramps_d = None if ramps is None else [up(r, cdt) for r in ramps]
        weight_d = up(weight, rdt)

        def core(x1, x2):
            return _spectral_core(x1, x2, ntrans, mode, detrend=detrend, windows=wins_d, keep_half=keep_half, shift=shifts,
                                  ramps=ramps_d, weight=weight_d, scale=scale, lut=lut, nbins=nbins)

        oh = None
        if out is not None:
            oh = out if _is_torch(out) else torch.from_numpy(out)
        res = _stream_host_chunks([d.data for d in das], ntrans, oh, core)
        return res.numpy() if out is None or not _is_torch(out) else res
    xs = []
    inv = None
    pad2 = None
    lp = das[0].lazy_pad if len(das) == 1 else None
    if (lp is not None and trailing and ntrans == 2 and real_dim is not None and mode == L.EPI_COMPLEX and detrend is None and window is None
            and not any_reversed and all(w == (0, 0) for w in lp.widths[:-2])):
        # xrft.pad -> xrft.fft(real_dim=...): the zero padding stays deferred, the transform reads the unpadded array
        pad2 = (lp.widths[-2], lp.widths[-1])
        xs.append(lp.base)
    for da, pl in zip(das if pad2 is None else [], plans):
        t = _device_tensor(da.data)
        t, inv = _to_last(t, P["axis_num"])
        if pl["reversed_dims"]:   # each array by ITS OWN coordinate orientation (xrft.py:436-441)
            from . import backend as B
            flip_axes = [t.ndim - ntrans + pl["dim"].index(d) for d in pl["reversed_dims"]]
            t = B.permute_flip(t, list(range(t.ndim)), flip_axes)
        xs.append(t)
    if len(xs) == 2 and xs[0].dtype != xs[1].dtype:
        dt = torch.promote_types(xs[0].dtype, xs[1].dtype)
        xs = [x.to(dt) for x in xs]
    if real_dim is not None and xs[0].is_complex():
        raise ValueError("real_dim requires real input data")
    seg_axis = None
    if seg_dim is not None:   # Welch: mean over the segment axis inside the spectral epilogue
        lead_axes = [a for a in range(nd) if a not in P["axis_num"]]
        seg_axis = lead_axes.index(P["da"].get_axis_num(seg_dim))
    res = _spectral_core(xs[0], xs[1] if len(xs) == 2 else None, ntrans, mode, detrend=detrend, windows=wins,
                         keep_half=real_dim is not None, shift=[P["shift"]] * ntrans if real_dim is None else [False] * ntrans,
                         ramps=ramps, weight=weight, scale=scale, lut=lut, nbins=nbins, with_phase=with_phase, pad2=pad2,
                         seg_axis=seg_axis)
    if seg_dim is not None:   # axes are now [leading axes without the segment axis] + [transform axes]: back to the array's order
        dims_all = list(P["da"].dims)
        cur = [dims_all[a] for a in lead_axes if dims_all[a] != seg_dim] + [dims_all[a] for a in P["axis_num"]]
        perm = [cur.index(d) for d in dims_all if d != seg_dim]
        return res.permute(*perm) if perm != list(range(len(perm))) else res
    if with_phase:
        res = tuple(r.permute(*inv) if inv is not None else r for r in res)
        return tuple(r.cpu().numpy() for r in res) if host_in else res
    if inv is not None and mode not in (L.EPI_BINS_POWER, L.EPI_BINS_CROSS):
        res = res.permute(*inv)
    if host_in:
        res = res.cpu().numpy()
        if out is not None:
            np.copyto(out.numpy() if _is_torch(out) else out, res)
    return res


def fft(da, spacing_tol=1e-3, dim=None, real_dim=None, shift=True, detrend=None, window=None, true_phase=True,
        true_amplitude=True, chunks_to_segments=False, prefix="freq_", real=None):
    """Discrete Fourier transform of `da` along `dim` -- see xrft.fft (xrft/xrft.py:307-476) for the semantics."""
    da = from_any(da)
    P = _fft_prepare(da, spacing_tol, dim, real_dim, shift, true_phase, chunks_to_segments, prefix, real)
    ramps = _phase_ramps(P) if true_phase else None
    scale = float(np.prod(P["delta_x"])) if true_amplitude else 1.0
    out = _run_forward(P, [P["da"]], L.EPI_COMPLEX, detrend, window, scale, ramps=ramps)
    return _label_output(P, out)


def dft(da, dim=None, true_phase=False, true_amplitude=False, **kwargs):  # xrft.py:237-250
    warnings.warn("This function has been renamed and will disappear in the future. Please use `fft` instead", FutureWarning)
    return fft(da, dim=dim, true_phase=true_phase, true_amplitude=true_amplitude, **kwargs)


def idft(daft, dim=None, true_phase=False, true_amplitude=False, **kwargs):  # xrft.py:253-266
    warnings.warn("This function has been renamed and will disappear in the future. Please use `ifft` instead", FutureWarning)
    return ifft(daft, dim=dim, true_phase=true_phase, true_amplitude=true_amplitude, **kwargs)


# =============================================================================================
# ifft  (xrft/xrft.py:479-646)
# =============================================================================================
def ifft(daft, spacing_tol=1e-3, dim=None, real_dim=None, shift=True, true_phase=True, true_amplitude=True,
         chunks_to_segments=False, prefix="freq_", lag=None, real=None):
    """Inverse discrete Fourier transform -- see xrft.ifft (xrft/xrft.py:479-646)."""
    from . import backend as B

    torch = _torch()
    daft = from_any(daft)
    if dim is None:
        dim = list(daft.dims)
    elif isinstance(dim, str):
        dim = [dim]
    dim = list(dim)
    if real is not None:
        real_dim = real
        warnings.warn(_real_flag_warning, FutureWarning)
    if real_dim is not None:
        if real_dim not in daft.dims:
            raise ValueError("The dimension along which real IFT is taken must be one of the existing dimensions.")
        dim = move_to_end(dim, real_dim)
    _check_valid_fft_coords(daft, dim)
    if lag is None:
        lag = [daft[d].attrs.get("direct_lag", 0.0) for d in dim]
        warnings.warn(
            "Default ifft's behaviour (lag=None) changed! Default value of lag was zero (centered output coordinates) and is now set to transformed coordinate's attribute: 'direct_lag'.",
            FutureWarning,
        )
    else:
        if isinstance(lag, (float, int)):
            lag = [lag]
        if len(dim) != len(lag):
            raise ValueError("dim and lag must have the same length.")
        if not true_phase:
            warnings.warn("Setting lag with true_phase=False does not guarantee accurate ifft.", Warning)
        lag = [daft[d].attrs.get("direct_lag") if l is None else l for d, l in zip(dim, lag)]

    # input-side ramp exp(+i 2 pi k lag) is indexed by the (unsorted) input coordinate (xrft.py:574-576)
    in_ramps = {d: np.exp(1j * 2.0 * np.pi * daft[d].values.astype(np.float64) * float(l)) for d, l in zip(dim, lag)} if true_phase else {}
    if chunks_to_segments:
        daft = _stack_chunks(daft, dim)
    rawdims = daft.dims
    if real_dim is not None:
        daft = daft.transpose(*move_to_end(list(daft.dims), real_dim))
    axis_num = [daft.get_axis_num(d) for d in dim]
    N = [daft.shape[n] for n in axis_num]
    # sortby(dim) (xrft.py:598): a permutation per axis, applied to data and ramp together
    orders = {}
    sorted_coords = {}
    for d in dim:
        c = daft[d].values
        order = np.argsort(c, kind="stable")
        orders[d] = order
        sorted_coords[d] = c[order]
    delta_x = []
    for d in dim:
        diff = np.diff(sorted_coords[d].astype(np.float64)) if not np.issubdtype(sorted_coords[d].dtype, np.datetime64) else _diff_coord(DataArray(sorted_coords[d], dims=(d,), name=d))
        delta = np.abs(diff[0])
        if not np.allclose(diff, diff[0], rtol=spacing_tol):
            raise ValueError("Can't take Fourier transform because coodinate %s is not evenly spaced" % d)
        if delta == 0.0:
            raise ValueError("Can't take Fourier transform because spacing in coordinate %s is zero" % d)
        delta_x.append(delta)
    for d in dim:  # xrft.py:600-606
        sc = sorted_coords[d]
        l = sc[len(sc) // 2] if d != real_dim else sc[0]
        if np.abs(l) > spacing_tol:
            raise ValueError("Inverse Fourier Transform can not be computed because coordinate %s is not centered on zero frequency" % d)

    host_in = not _is_torch(daft.data)
    t = _device_tensor(daft.data)
    if not t.is_complex():
        t = t.to(torch.complex64 if t.dtype == torch.float32 else torch.complex128)
    t, inv = _to_last(t, axis_num)
    ntrans = len(dim)
    if ntrans > 3:
        raise NotImplementedError("transforms over more than 3 dimensions are not supported")
    k = _ifreq(N, delta_x, real_dim, shift)
    n_outs = [kk.size for kk in k]
    out_shifts = []
    for n_out in n_outs:
        s = 0
        if not true_phase:
            s += n_out - n_out // 2      # ifftshift  (xrft.py:617-618)
        if shift:
            s += n_out // 2              # fftshift   (xrft.py:620-621)
        out_shifts.append(s % n_out)
    spacings = [kk[1] - kk[0] for kk in k]
    scale = 1.0 / float(np.prod([float(s) for s in spacings])) if true_amplitude else 1.0
    f = None
    if real_dim is not None and ntrans == 2:
        # one chain of passes (xrftb_fft2r): sortby + ifftshift of the strided axis are a cyclic roll of its rows (folded into
        # the loads together with the phase ramps); output shifts and 1 / prod(df) ride on the stores
        ny_ = t.shape[-2]
        oy_, ox_ = orders[dim[0]], orders[dim[1]]
        s0 = int(oy_[0])
        if (np.array_equal(oy_, (np.arange(ny_) + s0) % ny_) and np.array_equal(ox_, np.arange(ox_.size))
                and B.fft2r_supported(ny_, n_outs[1], torch.float32 if t.dtype == torch.complex64 else torch.float64)):
            tr = lambda v: torch.from_numpy(np.ascontiguousarray(v)) if v is not None else None
            # deferred: xrft.unpad of the result narrows the box the inverse transform stores before anything is computed
            f = LazyIrfft2(t, (ny_ // 2 + s0) % ny_, tr(in_ramps[dim[0]]) if true_phase else None,
                           tr(in_ramps[dim[1]]) if true_phase else None, out_shifts, scale)
            if inv is not None or host_in:
                f = f.materialize()
    if f is None:
        # sort + ramp + ifftshift of the input: data movement (gather) + one per-axis complex vector
        for i, d in enumerate(dim):
            ax = t.ndim - ntrans + i
            order = orders[d]
            if not np.array_equal(order, np.arange(order.size)):
                t = t.index_select(ax, torch.as_tensor(order, device=t.device))
        ramps = []
        shifts = []
        for i, d in enumerate(dim):
            r = in_ramps[d][orders[d]] if true_phase else None
            ramps.append(r)
            shifts.append(2 if d != real_dim else 0)  # ifftshift on the non-real axes (xrft.py:608-614)
        kin = list(t.shape[-ntrans:])
        t = B.spectral_post(t, None, L.EPI_COMPLEX, ntrans, kin[-1], hermitian=False, keep_half=False, shift=shifts,
                            ramps=[torch.from_numpy(np.ascontiguousarray(r)) if r is not None else None for r in ramps],
                            weight=None, scale=1.0)
        axes = list(range(t.ndim - ntrans, t.ndim))
        f = B.ifftn(t, axes) if real_dim is None else B.irfftn(t, axes)
        f = B.roll_scale(f, ntrans, out_shifts, scale)
    if inv is not None:
        f = f.permute(*inv)
    if host_in:
        f = f.cpu().numpy()

    swap = {d: _new_name(d, prefix) for d in dim}
    out = DataArray(f, dims=[swap.get(d, d) for d in daft.dims])
    for cname, c in daft.coords.items():
        if cname not in dim:
            OrderedDict.__setitem__(out._coords, cname, c)
    for d, kk, l in zip(dim, k, lag):
        nn = swap[d]
        OrderedDict.__setitem__(out._coords, nn, DataArray(kk + l, dims=(nn,), name=nn, attrs={"spacing": kk[1] - kk[0]}, _coord=True))
    return out.transpose(*[swap.get(d, d) for d in rawdims])


# =============================================================================================
# spectra  (xrft/xrft.py:649-874)
# =============================================================================================
def _window_correction_factor(da, dim, scaling, window):  # xrft.py:649-660
    if window is None:
        raise ValueError("window_correction can only be applied when windowing is turned on.")
    if dim is None:
        dim = list(da.dims)
    elif isinstance(dim, str):
        dim = [dim]
    ws = _window_vectors([da.sizes[d] for d in dim], window)
    if scaling == "density":
        return float(np.prod([(w ** 2).mean() for w in ws]))  # mean of a separable product = product of means
    elif scaling == "spectrum":
        return float(np.prod([w.mean() for w in ws]) ** 2)
    raise ValueError("Unknown {} scaling flag".format(scaling))


def _psd_scaling_factor(spacings, scaling):  # xrft.py:663-670
    fs = float(np.prod([float(s) for s in spacings]))
    if scaling == "density":
        return fs
    elif scaling == "spectrum":
        return fs ** 2
    raise ValueError("Unknown {} scaling flag".format(scaling))


def _real_dim_weights(n_real, n_half):  # xrft.py:673-682
    f = np.full(n_half, 2.0)
    if n_real % 2 == 0:
        f[0], f[-1] = 1.0, 1.0
    else:
        f[0] = 1.0
    return f


def _spectrum_scale(P, da, dim_arg, real_dim, scaling, window_correction, window):
    """Scalar multiplying |F|^2 or F conj(G): (prod dx)^2 [true_amplitude] / window factor x psd factor."""
    scale = float(np.prod(P["delta_x"])) ** 2
    if scaling != "false_density":
        if window_correction:
            scale /= _window_correction_factor(da, dim_arg, scaling, window)
        updated = [kk[1] - kk[0] for kk in P["k"]]
        scale *= _psd_scaling_factor(updated, scaling)
    return scale


def power_spectrum(da, dim=None, real_dim=None, scaling="density", window_correction=False, **kwargs):
    """xrft.power_spectrum (xrft/xrft.py:685-750), one fused device pass for real 2-D power-of-two fields."""
    da0 = from_any(da)
    if "density" in kwargs:
        density = kwargs.pop("density")
        warnings.warn(
            "density flag will be deprecated in future version of xrft.power_spectrum and replaced by scaling flag. "
            + 'density=True should be replaced by scaling="density" and density=False will not be maintained.\nscaling flag is ignored !',
            FutureWarning,
        )
        scaling = "density" if density else "false_density"
    if "real" in kwargs:
        real_dim = kwargs.pop("real")
        warnings.warn(_real_flag_warning, FutureWarning)
    kwargs.update({"true_amplitude": True, "true_phase": False})
    P, out = _spectrum(da0, None, L.EPI_POWER, dim, real_dim, scaling, window_correction, kwargs)
    return _label_spectrum(P, out)


def _spectrum(da1, da2, mode, dim, real_dim, scaling, window_correction, kwargs, bins=None, with_phase=False):
    kw = dict(kwargs)
    out_buf = kw.pop("out", None)  # extension: preallocated (pinned) host result buffer for streamed host inputs
    detrend_t = kw.pop("detrend", None)
    window = kw.pop("window", None)
    true_phase = kw.pop("true_phase", True)
    kw.pop("true_amplitude", None)
    spacing_tol = kw.pop("spacing_tol", 1e-3)
    shift = kw.pop("shift", True)
    c2s = kw.pop("chunks_to_segments", False)
    prefix = kw.pop("prefix", "freq_")
    if kw:
        raise TypeError("fft() got an unexpected keyword argument '%s'" % next(iter(kw)))
    P = _fft_prepare(da1, spacing_tol, dim, real_dim, shift, true_phase, c2s, prefix, None)
    das = [P["da"]]
    if da2 is not None:
        P2 = _fft_prepare(da2, spacing_tol, dim, real_dim, shift, true_phase, c2s, prefix, None)
        out_dims1 = [_new_name(d, prefix) if d in P["dim"] else d for d in P["rawdims"]]
        out_dims2 = [_new_name(d, prefix) if d in P2["dim"] else d for d in P2["rawdims"]]
        if out_dims1 != out_dims2:
            raise ValueError("The two datasets have different dimensions")
        if P2["da"].shape != P["da"].shape:
            raise ValueError("The two datasets have different dimensions")
        das.append(P2["da"])
    if scaling not in ("density", "spectrum", "false_density"):
        raise ValueError("Unknown {} scaling flag".format(scaling))
    scale = _spectrum_scale(P, da1, dim, P["real_dim"], scaling, window_correction, window)
    ramps = None
    if da2 is not None and true_phase:
        r1, r2 = _phase_ramps(P), _phase_ramps(P2)
        ramps = [a * np.conj(b) for a, b in zip(r1, r2)]
        if all(np.allclose(r, 1.0, rtol=0, atol=1e-15) for r in ramps):
            ramps = None
    weight = None
    if P["real_dim"] is not None:
        n_real = da1.sizes[P["real_dim"]]   # len(da[real_dim]) of the un-segmented array, like the reference (xrft.py:678)
        weight = _real_dim_weights(n_real, P["N"][-1] // 2 + 1)
    lut, nbins = (None, 0) if bins is None else bins(P)
    plans = [P, P2] if da2 is not None else None
    seg_dims = [d for d in P["da"].dims if d.endswith("_segment")] if c2s else []
    if (seg_dims and mode in (L.EPI_POWER, L.EPI_CROSS) and not with_phase and bins is None and out_buf is None
            and all(_is_torch(d.data) and d.data.is_cuda for d in das) and not any(pl["reversed_dims"] for pl in (plans or [P]))):
        # chunks_to_segments on device data: the per-segment spectra are deferred, so that a following .mean over a segment
        # axis (Welch) is reduced inside the spectral epilogue instead of being a pass over spectra written to memory
        def run(seg_dim):
            return _run_forward(P, das, mode, detrend_t, window, scale, ramps=ramps, weight=weight, plans=plans, seg_dim=seg_dim)

        shape = list(P["da"].shape)
        if P["real_dim"] is not None:
            shape[P["axis_num"][-1]] = P["N"][-1] // 2 + 1
        nat_dims = [_new_name(d, prefix) if d in P["dim"] else d for d in P["da"].dims]
        return P, LazySegSpectrum(run, shape, nat_dims, seg_dims)
    out = _run_forward(P, das, mode, detrend_t, window, scale, ramps=ramps, weight=weight, lut=lut, nbins=nbins, out=out_buf,
                       plans=plans, with_phase=with_phase)
    return P, out


def _label_spectrum(P, out):
    return _label_output(P, out, extra_attrs=P["true_phase"])


def cross_spectrum(da1, da2, dim=None, real_dim=None, scaling="density", window_correction=False, true_phase=True, **kwargs):
    """xrft.cross_spectrum (xrft/xrft.py:753-835)."""
    da1, da2 = from_any(da1), from_any(da2)
    if "real" in kwargs:
        real_dim = kwargs.pop("real")
        warnings.warn(_real_flag_warning, FutureWarning)
    if "density" in kwargs:
        density = kwargs.pop("density")
        warnings.warn(
            "density flag will be deprecated in future version of xrft.cross_spectrum and replaced by scaling flag. "
            + 'density=True should be replaced by scaling="density" and density=False will not be maintained.\nscaling flag is ignored !',
            FutureWarning,
        )
        scaling = "density" if density else "false_density"
    kwargs.update({"true_amplitude": True, "true_phase": true_phase})
    P, out = _spectrum(da1, da2, L.EPI_CROSS, dim, real_dim, scaling, window_correction, kwargs)
    return _label_spectrum(P, out)


def cross_spectrum_and_phase(da1, da2, dim=None, real_dim=None, scaling="density", window_correction=False, true_phase=True, **kwargs):
    """(cross_spectrum(da1, da2, ...), cross_phase(da1, da2, ...)) from ONE pass over the two fields.  The reference computes
    the phase by running the whole cross-spectrum pipeline a second time (xrft/xrft.py:865-869 calls cross_spectrum, which
    runs two more fft pipelines); here the angle is a second store of the same epilogue.  Arguments as cross_spectrum."""
    da1, da2 = from_any(da1), from_any(da2)
    if "real" in kwargs:
        real_dim = kwargs.pop("real")
        warnings.warn(_real_flag_warning, FutureWarning)
    if "density" in kwargs:
        scaling = "density" if kwargs.pop("density") else "false_density"
    kwargs.update({"true_amplitude": True, "true_phase": true_phase})
    P, (cs, ph) = _spectrum(da1, da2, L.EPI_CROSS, dim, real_dim, scaling, window_correction, kwargs, with_phase=True)
    cs, cp = _label_spectrum(P, cs), _label_spectrum(P, ph)
    if da1.name and da2.name:
        cp.name = "{}_{}_phase".format(da1.name, da2.name)
    return cs, cp


def cross_phase(da1, da2, dim=None, true_phase=True, **kwargs):
    """xrft.cross_phase (xrft/xrft.py:838-874): angle of the cross spectrum in one fused pass."""
    da1, da2 = from_any(da1), from_any(da2)
    real_dim = kwargs.pop("real_dim", None)
    scaling = kwargs.pop("scaling", "density")
    window_correction = kwargs.pop("window_correction", False)
    if "real" in kwargs:
        real_dim = kwargs.pop("real")
        warnings.warn(_real_flag_warning, FutureWarning)
    if "density" in kwargs:
        density = kwargs.pop("density")
        warnings.warn("density flag will be deprecated in future version of xrft.cross_spectrum", FutureWarning)
        scaling = "density" if density else "false_density"
    kwargs.update({"true_amplitude": True, "true_phase": true_phase})
    P, out = _spectrum(da1, da2, L.EPI_PHASE, dim, real_dim, scaling, window_correction, kwargs)
    cp = _label_spectrum(P, out)
    if da1.name and da2.name:
        cp.name = "{}_{}_phase".format(da1.name, da2.name)
    return cp


# =============================================================================================
# isotropic spectra  (xrft/xrft.py:877-1187)
# =============================================================================================
def _cut_codes(values, nbins):
    """Integer codes of pandas.cut(ravel(values), nbins) (xrft.py:921), restated: equal-width edges over
    [min, max], first edge lowered by 0.1 % of the range, right-closed bins."""
    v = np.ravel(values)
    mn, mx = float(v.min()), float(v.max())
    if mn == mx:
        mn -= 0.001 * abs(mn) if mn != 0 else 0.001
        mx += 0.001 * abs(mx) if mx != 0 else 0.001
        edges = np.linspace(mn, mx, nbins + 1, endpoint=True)
    else:
        edges = np.linspace(mn, mx, nbins + 1, endpoint=True)
        edges[0] -= (mx - mn) * 0.001
    return (edges.searchsorted(v, side="left") - 1).astype(np.int64)


_LUT_TENSORS = OrderedDict()    # id(codes) -> (codes, int32 LUT tensor laid out like the spectrum): same object every call
_RADIAL_CACHE = OrderedDict()   # (k bytes, l bytes, nfactor, truncate) -> (codes, nbins, kr): the LUT of a grid is built once


def _radial_bins(k, l, nfactor, truncate):
    """freq_r, LUT codes [len(k), len(l)], nbins and the bin-mean radius coordinate (xrft.py:975-991)."""
    key = (k.tobytes(), l.tobytes(), nfactor, bool(truncate))
    hit = _RADIAL_CACHE.get(key)
    if hit is not None:
        if not truncate:
            warnings.warn("Isotropic wavenumber larger than the Nyquist wavenumber may result.", FutureWarning)
        return hit
    res = _radial_bins_build(k, l, nfactor, truncate)
    _RADIAL_CACHE[key] = res
    while len(_RADIAL_CACHE) > 8:
        _RADIAL_CACHE.popitem(last=False)
    return res


def _radial_bins_build(k, l, nfactor, truncate):
    N = [k.size, l.size]
    nbins = int(min(N) / nfactor)
    freq_r = np.sqrt(k[:, None] ** 2 + l[None, :] ** 2)
    codes = _cut_codes(freq_r, nbins).reshape(freq_r.shape)
    cnt = np.bincount(codes.ravel(), minlength=nbins)
    kr = np.bincount(codes.ravel(), weights=freq_r.ravel(), minlength=nbins) / np.where(cnt == 0, 1, cnt)
    if truncate:
        kmax = l.max() if k.max() > l.max() else k.max()
        kr = np.where(kr <= kmax, kr, np.nan)
    else:
        warnings.warn("Isotropic wavenumber larger than the Nyquist wavenumber may result.", FutureWarning)
    return codes, nbins, kr


def isotropize(ps, fftdim, nfactor=4, truncate=True, complx=False):
    """Azimuthal sum of a 2-D (cross-)spectrum (xrft/xrft.py:948-1010) on the device."""
    from . import backend as B

    torch = _torch()
    ps = from_any(ps)
    k = ps[fftdim[1]].values.astype(np.float64)
    l = ps[fftdim[0]].values.astype(np.float64)
    codes, nbins, kr = _radial_bins(k, l, nfactor, truncate)
    others = [d for d in ps.dims if d not in fftdim]
    pst = ps.transpose(*(others + [fftdim[1], fftdim[0]]))
    t = _device_tensor(pst.data).contiguous()
    if complx and not t.is_complex():
        t = t.to(torch.complex128 if t.dtype == torch.float64 else torch.complex64)
    iso = B.binned_sum(t, torch.from_numpy(codes.astype(np.int32)), nbins, 2)
    if not complx:
        iso = iso.to(t.dtype)  # output_dtypes=[array.dtype] (xrft.py:931)
    if not _is_torch(ps.data):
        iso = iso.cpu().numpy()
    out = DataArray(iso, dims=others + ["freq_r"], name=ps.name)
    for cname, c in ps.coords.items():
        if cname not in fftdim and not any(d in fftdim for d in c.dims):
            OrderedDict.__setitem__(out._coords, cname, c)
    OrderedDict.__setitem__(out._coords, "freq_r", DataArray(kr, dims=("freq_r",), name="freq_r", _coord=True))
    # truncate=True -> dropna("freq_r") looks at the DATA only: nothing is dropped
    return out


def _iso_common(da1, da2, mode, spacing_tol, dim, shift, detrend, scaling, window, window_correction, nfactor, truncate, kwargs):
    if "density" in kwargs:
        density = kwargs.pop("density")
        scaling = "density" if density else "false_density"
    if dim is None:
        dim = da1.dims
        if da2 is not None and dim != da2.dims:
            raise ValueError("The two datasets have different dimensions")
    if len(dim) != 2:
        raise ValueError("The Fourier transform should be two dimensional")
    dim = list(dim)
    state = {}

    def bins(P):
        # frequency coordinates of the OUTPUT grid; LUT is laid out like the spectrum's trailing axes (dim[0], dim[1])
        kk = dict(zip(P["dim"], P["k"]))
        k, l = kk[dim[1]].astype(np.float64), kk[dim[0]].astype(np.float64)
        codes, nbins, kr = _radial_bins(k, l, nfactor, truncate)  # codes[i_k, i_l]
        state["kr"], state["nbins"] = kr, nbins
        order = [P["dim"].index(dim[0]), P["dim"].index(dim[1])]
        import torch
        ck = (id(codes), order[0])
        hit = _LUT_TENSORS.get(ck)
        if hit is None or hit[0] is not codes:
            lut = codes.T if order == [0, 1] else codes  # -> [axis of P.dim[0]][axis of P.dim[1]]
            hit = (codes, torch.from_numpy(np.ascontiguousarray(lut.astype(np.int32))))
            _LUT_TENSORS[ck] = hit
            while len(_LUT_TENSORS) > 8:
                _LUT_TENSORS.popitem(last=False)
        return hit[1], nbins

    kw = dict(kwargs)
    kw.update(dict(spacing_tol=spacing_tol, shift=shift, detrend=detrend, window=window, true_amplitude=True))
    if da2 is None:
        kw["true_phase"] = False
    else:
        kw.setdefault("true_phase", True)
    P, out = _spectrum(da1, da2, mode, dim, None, scaling, window_correction, kw, bins=bins)
    others = [d for d in P["da"].dims if d not in P["dim"]]
    if da2 is None and str(P["da"].data.dtype).endswith("float32"):
        # output_dtypes=[array.dtype] (xrft.py:931); accumulation itself is fp64
        out = out.astype(np.float32) if isinstance(out, np.ndarray) else out.float()
    res = DataArray(out, dims=others + ["freq_r"])
    for cname, c in P["da"].coords.items():
        if cname not in P["dim"] and not any(d in P["dim"] for d in c.dims):
            OrderedDict.__setitem__(res._coords, cname, c)
    OrderedDict.__setitem__(res._coords, "freq_r", DataArray(state["kr"], dims=("freq_r",), name="freq_r", _coord=True))
    return res


def isotropic_power_spectrum(da, spacing_tol=1e-3, dim=None, shift=True, detrend=None, scaling="density", window=None,
                             window_correction=False, nfactor=4, truncate=False, **kwargs):
    """xrft.isotropic_power_spectrum (xrft/xrft.py:1013-1095); the radial-bin sum is fused into the FFT epilogue."""
    da = from_any(da)
    return _iso_common(da, None, L.EPI_BINS_POWER, spacing_tol, dim, shift, detrend, scaling, window, window_correction,
                       nfactor, truncate, kwargs)


def isotropic_cross_spectrum(da1, da2, spacing_tol=1e-3, dim=None, shift=True, detrend=None, scaling="density", window=None,
                             window_correction=False, nfactor=4, truncate=False, **kwargs):
    """xrft.isotropic_cross_spectrum (xrft/xrft.py:1098-1187)."""
    da1, da2 = from_any(da1), from_any(da2)
    return _iso_common(da1, da2, L.EPI_BINS_CROSS, spacing_tol, dim, shift, detrend, scaling, window, window_correction,
                       nfactor, truncate, kwargs)


def fit_loglog(x, y):  # xrft.py:1190-1214 (host post-processing)
    p = np.polyfit(np.log2(x), np.log2(y), 1)
    y_fit = 2 ** (np.log2(x) * p[0] + p[1])
    return y_fit, p[0], p[1]


# =============================================================================================
# detrend  (xrft/detrend.py:11-97)
# =============================================================================================
def detrend(da, dim, detrend_type="constant"):
    """xrft.detrend: mean or least-squares (hyper)plane removal over 1-3 dims, computed on the device."""
    from . import backend as B

    da = from_any(da)
    if dim is None:
        dim = list(da.dims)
    elif isinstance(dim, str):
        dim = [dim]
    if detrend_type not in ["constant", "linear", None]:
        raise NotImplementedError(
            "%s is not a valid detrending option. Valid options are: 'constant','linear', or None." % detrend_type
        )
    if detrend_type is None:
        return da
    if detrend_type == "linear":
        chunks = da.chunks
        if chunks:
            axis_chunks = [chunks[da.get_axis_num(d)] for d in dim]
            if not all(len(ac) == 1 for ac in axis_chunks):
                raise ValueError("Contiguous chunks required for detrending.")
        if len(dim) > 3:
            raise NotImplementedError("Only 1D, 2D, and 3D detrending are implemented so far.")
    elif len(dim) > 3:
        raise NotImplementedError("constant detrend over more than 3 dims is not supported on the device path")
    torch = _torch()
    axes = [da.get_axis_num(d) for d in dim]
    t = _device_tensor(da.data)
    t, inv = _to_last(t, axes)
    det = 1 if detrend_type == "constant" else 2
    if t.is_complex():
        re = B.detrend_window(t.real.contiguous(), len(dim), det)
        im = B.detrend_window(t.imag.contiguous(), len(dim), det)
        out = torch.complex(re, im)
    else:
        out = B.detrend_window(t, len(dim), det)
    if inv is not None:
        out = out.permute(*inv)
    if not _is_torch(da.data):
        out = out.cpu().numpy()
    return da._replace(data=out)


# =============================================================================================
# padding  (xrft/padding.py:11-446)
# =============================================================================================
def _get_spacing(coord):  # xrft/utils.py:11-20
    diff = _diff_coord(coord)
    if not np.allclose(diff, diff[0]):
        raise ValueError(f"Found unevenly spaced coordinates '{coord.name}'. These coordinates should be evenly spaced.")
    return diff[0]


def _pad_coordinate(values, width, spacing):  # xrft/padding.py:263-318
    n_start, n_end = (width, width) if isinstance(width, (int, np.integer)) else width
    out = np.pad(values, (n_start, n_end))
    vmin, vmax = values[0], values[-1]
    out[:n_start] = vmin - n_start * spacing + np.linspace(0, spacing * (n_start - 1), n_start)
    out[len(out) - n_end:] = vmax + spacing + np.linspace(0, spacing * (n_end - 1), n_end)
    return out


def pad(da, pad_width=None, mode="constant", stat_length=None, constant_values=0, end_values=None, reflect_type=None,
        **pad_width_kwargs):
    """xrft.pad (xrft/padding.py:11-181): pad data and linearly extrapolate evenly spaced coordinates."""
    da = from_any(da)
    pad_width = either_dict_or_kwargs(pad_width, pad_width_kwargs, "pad")
    bad_coords = []
    for coord in pad_width.keys():  # padding.py:184-215
        d = da[coord].dims[0]
        bad_coords += [c for c in da.coords if d in da[c].dims and c != coord]
    if bad_coords:
        bad = "'" + "', '".join(bad_coords) + "'"
        raise ValueError("Please, drop the following coordinates from the passed DataArray " + f"before trying to pad it: {bad}.")
    if _is_torch(da.data) and da.data.is_cuda:
        # device-resident data is padded on the device (xrftb_pad): constant / edge / reflect / symmetric / wrap.  The statistic
        # modes of numpy.pad (mean, median, minimum, maximum, linear_ramp, empty) have no device kernel and fail loudly
        from . import backend as B
        if mode not in B.PAD_MODES or stat_length is not None or end_values is not None or reflect_type not in (None, "even"):
            raise NotImplementedError(
                f"pad(mode={mode!r}, stat_length={stat_length}, end_values={end_values}, reflect_type={reflect_type}) is not available for "
                f"device-resident data; supported modes: {sorted(B.PAD_MODES)} (reflect_type='even')")
        if mode == "constant" and not np.isscalar(constant_values):
            raise NotImplementedError("pad: per-axis constant_values are not available for device-resident data")
        widths = []
        for d in da.dims:
            w = pad_width.get(d, 0)
            widths.append((int(w), int(w)) if isinstance(w, (int, np.integer)) else (int(w[0]), int(w[1])))
        if mode == "constant" and constant_values == 0 and da.data.dtype in (_torch().float32, _torch().float64):
            data = LazyPad(da.data.contiguous(), widths)   # deferred: xrft.fft reads the unpadded array through load predicates
        else:
            data = B.pad(da.data, widths, mode, constant_values if mode == "constant" else 0)
        padded = da._replace(data=data)
        for d in pad_width:
            if d in padded._coords:
                OrderedDict.__delitem__(padded._coords, d)
    else:   # host-resident data stays on the host: numpy.pad, as in the reference (pure data movement where the data lives)
        padded = da.pad(pad_width, mode, stat_length, constant_values, end_values, reflect_type)
    for d in pad_width:
        cvals = da[d].values
        spacing = _get_spacing(da[d])
        new = _pad_coordinate(cvals, pad_width[d], spacing)
        attrs = dict(da[d].attrs)
        attrs.update({"pad_width": pad_width[d]})
        OrderedDict.__setitem__(padded._coords, d, DataArray(new, dims=(d,), name=d, attrs=attrs, _coord=True))
    return padded


def _pad_width_to_slice(pad_width, size):  # padding.py:425-446
    if isinstance(pad_width, (int, np.integer)):
        pad_width = (pad_width, pad_width)
    return slice(pad_width[0], size - pad_width[1])


def unpad(da, pad_width=None, **pad_width_kwargs):
    """xrft.unpad (xrft/padding.py:321-422)."""
    da = from_any(da)
    if pad_width is None and not pad_width_kwargs:
        pad_width = {d: c.attrs["pad_width"] for d, c in da.coords.items() if "pad_width" in c.attrs}
        if not pad_width:
            raise ValueError(
                "The passed array doesn't seem to be a padded one: the 'pad_width' attribute was missing on every one of its coordinates. "
            )
    else:
        pad_width = either_dict_or_kwargs(pad_width, pad_width_kwargs, "pad")
    slices = {d: _pad_width_to_slice(pad_width[d], da.coords[d].size) for d in pad_width}
    out = da.isel(indexers=slices)
    for d in pad_width:
        c = out.coords[d]
        if "pad_width" in c.attrs:
            c2 = c._replace()
            c2.attrs.pop("pad_width")
            OrderedDict.__setitem__(out._coords, d, c2)
    return out
