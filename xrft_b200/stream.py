"""Out-of-core execution: a lazy chunk iterator in, a lazy result iterator out.

The role dask's chunk iteration plays in the reference (`xr.apply_ufunc(..., dask="parallelized")`, xrft/xrft.py:925-943;
chunks of the non-transform axes are independent, xrft/xrft.py:32-36): arrays larger than device (or host) memory flow
through the GPU chunk by chunk.  `chunks` is any iterable of DataArrays with host (numpy / memory-mapped) data -- it is
consumed lazily, one chunk ahead -- and the results are yielded as DataArrays with numpy data in the same order.

Three CUDA streams overlap the host->device copy of chunk i+1 (from a pinned staging buffer), the kernels of chunk i and
the device->host copy of result i-1.  `func` is any function of this package that maps device DataArrays to a device
DataArray (power_spectrum, cross_spectrum, isotropic_power_spectrum, fft, ...); with several input iterables the chunks
are zipped (cross spectra).  There is no CPU fallback: the numerics run in the CUDA kernels behind the C-ABI.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Iterable, Iterator

import numpy as np

from .dataarray import DataArray, from_any


def _coords_of(da: DataArray):
    c = OrderedDict()
    for name, v in da.coords.items():
        c[name] = v
    return c


def stream(func: Callable, *chunk_iterables: Iterable, **kwargs) -> Iterator[DataArray]:
    """Yield func(chunk, ..., **kwargs) for every chunk (tuple of zipped chunks) of the input iterable(s)."""
    import torch
    from . import backend as B

    B.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    its = [iter(c) for c in chunk_iterables]

    def upload():
        """next chunk tuple -> (host DataArrays, device tensors, event) or None at the end; copies run on s_in"""
        try:
            host = [from_any(next(it)) for it in its]
        except StopIteration:
            return None
        tens = []
        with torch.cuda.stream(s_in):
            for h in host:
                a = np.ascontiguousarray(np.asarray(h.data))
                if a.dtype.kind in "iub":
                    a = a.astype(np.float64)
                try:
                    pinned = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].reshape(-1)).dtype, pin_memory=True)
                    pinned.numpy()[...] = a           # staging copy (the source may be a memory map)
                except Exception:  # pragma: no cover - pinned allocation can fail on small hosts
                    pinned = torch.from_numpy(a)
                t = torch.empty(pinned.shape, dtype=pinned.dtype, device=dev)
                t.copy_(pinned, non_blocking=True)
                tens.append((t, pinned))
            ev = torch.cuda.Event()
            ev.record(s_in)
        return host, tens, ev

    def launch(item):
        """kernels of one chunk on s_cmp, result copy on s_out -> (result DataArray on the device, pinned result, event)"""
        host, tens, ev = item
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev)
            das = [DataArray(t, dims=h.dims, coords=_coords_of(h), name=h.name, attrs=h.attrs, chunks=h._chunks) for h, (t, _) in zip(host, tens)]
            res = func(*das, **kwargs)
            rt = res.data                      # deferred results are computed here, on s_cmp
            done = torch.cuda.Event()
            done.record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            rt = rt.contiguous()
            try:
                hbuf = torch.empty(rt.shape, dtype=rt.dtype, pin_memory=True)
            except Exception:  # pragma: no cover
                hbuf = torch.empty(rt.shape, dtype=rt.dtype)
            hbuf.copy_(rt, non_blocking=True)
            rt.record_stream(s_out)
            out_ev = torch.cuda.Event()
            out_ev.record(s_out)
        return res, hbuf, out_ev, tens

    nxt = upload()
    pending = None
    while nxt is not None or pending is not None:
        cur = launch(nxt) if nxt is not None else None
        nxt = upload() if nxt is not None else None      # chunk i+1 travels while chunk i computes
        if pending is not None:
            res, hbuf, out_ev, _keep = pending
            out_ev.synchronize()
            yield res._replace(data=hbuf.numpy())
        pending = cur
