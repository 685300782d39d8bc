"""Multi-GPU data parallelism: the role dask plays in the reference (chunks of the non-transform axes are
independent; xrft/xrft.py:32-36), rebuilt as one process per GPU.

Transforms never cross ranks (transform axes are unchunked in the reference too: test_xrft.py:166-170), so
fft / power_spectrum / cross_* need NO collective: each rank works on a contiguous block of the outermost
non-transform axis.  The only exchange step is the reduction of isotropic (radial-bin) spectra over the
sharded axis: one all-reduce of `nbins` float64 values (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from .dataarray import DataArray, from_any


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n` items owned by `rank`: sizes differ by at most one, in rank order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_shard(da, dim: str, rank: Optional[int] = None, world: Optional[int] = None) -> DataArray:
    """This rank's block of `da` along `dim` (rank/world default to torch.distributed's)."""
    da = from_any(da)
    rank, world = _rank_world(rank, world)
    lo, hi = shard_bounds(da.sizes[dim], rank, world)
    return da.isel({dim: slice(lo, hi)})


def _rank_world(rank, world):
    if rank is not None and world is not None:
        return rank, world
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class BinsComm:
    """The C-ABI communicator of the radial-bin all-reduce (include/xrft_b200.h, xrftb_comm_* / xrftb_allreduce_bins ->
    ncclAllReduce on the caller's stream).  The NCCL unique id travels from rank 0 through torch.distributed (any
    backend); one communicator per process, created on first use on the current CUDA device."""

    _cached = None

    def __init__(self, rank: int, world: int):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib as L

        self.lib = L.load()
        self.rank, self.world = rank, world
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            L.check(self.lib.xrftb_comm_unique_id(idbuf), "xrftb_comm_unique_id")
        if world > 1:
            box = [idbuf.raw if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            idbuf = C.create_string_buffer(box[0], 128)
        self.handle = C.c_void_p()
        with torch.cuda.device(torch.cuda.current_device()):
            L.check(self.lib.xrftb_comm_init(C.byref(self.handle), world, rank, idbuf), "xrftb_comm_init")
        self.device = torch.cuda.current_device()

    @classmethod
    def get(cls, rank: int, world: int) -> "BinsComm":
        import torch

        c = cls._cached
        if c is None or (c.rank, c.world, c.device) != (rank, world, torch.cuda.current_device()):
            if c is not None:
                c.close()
            c = cls._cached = cls(rank, world)
        return c

    def allreduce(self, t):
        """in-place sum over the ranks of a contiguous float64 CUDA tensor, on torch's current stream"""
        import ctypes as C
        import torch
        from . import _lib as L

        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        with torch.cuda.device(t.device):
            rc = self.lib.xrftb_allreduce_bins(self.handle, C.c_void_p(t.data_ptr()), t.numel(),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream))
        L.check(rc, "xrftb_allreduce_bins")
        return t

    def close(self):
        if self.handle:
            self.lib.xrftb_comm_destroy(self.handle)
            self.handle = None
        if BinsComm._cached is self:
            BinsComm._cached = None


_LAST = ("none", 1)


def last_collective():
    """(which collective the last allreduce_sum used, number of ranks) -- evidence for the tests and bench.py"""
    return _LAST


def allreduce_sum(t):
    """In-place sum over the ranks of a float64 tensor.  CUDA tensors go through the C-ABI (xrftb_allreduce_bins: NCCL
    over NVLink); host tensors through torch.distributed (gloo, the CPU tests).  No-op semantics for one rank without an
    initialised process group."""
    global _LAST
    import torch.distributed as dist

    rank, world = _rank_world(None, None)
    initialised = dist.is_available() and dist.is_initialized()
    if t.is_cuda and (initialised or world == 1):
        BinsComm.get(rank, world).allreduce(t)
        _LAST = ("xrftb_allreduce_bins (NCCL)", world)
    elif initialised and world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        _LAST = ("torch.distributed (%s)" % dist.get_backend(), world)
    else:
        _LAST = ("none", 1)
    return t


def _column_sums(vals):
    """float64 sums over the leading axis of a [items][nbins] array: the radial-bin kernel of the C-ABI on the device
    (every item mapped onto the identity LUT), plain torch on the host (gloo tests)."""
    import torch

    if vals.is_cuda and not vals.is_complex():
        from . import backend as B

        nb = vals.shape[-1]
        lut = torch.arange(nb, dtype=torch.int32, device=vals.device).repeat(vals.shape[0])
        return B.binned_sum(vals.reshape(1, -1), lut, nb, 1).reshape(nb)
    return vals.to(torch.float64).sum(dim=0)


def sharded_isotropic_mean(da, shard_dim: str, dim: Sequence[str], compute: Optional[Callable] = None, presharded: bool = False,
                           **kwargs) -> DataArray:
    """mean over every non-transform axis of isotropic_power_spectrum(da, dim=...), with `shard_dim` split over
    the ranks: each rank reduces its own block to `nbins` partial sums, then ONE all-reduce (sum + count).
    presharded=True: `da` already IS this rank's block of the global array (each rank generated / loaded its own)."""
    import torch

    if compute is None:
        from .api import isotropic_power_spectrum as compute
    da = from_any(da)
    mine = da if presharded else local_shard(da, shard_dim)
    n_local = 1
    for d in mine.dims:
        if d not in dim:
            n_local *= mine.sizes[d]
    if n_local > 0:
        iso = compute(mine, dim=list(dim), **kwargs)
        vals = iso.data
        if not isinstance(vals, torch.Tensor):
            vals = torch.as_tensor(np.asarray(vals))
        part = _column_sums(vals.reshape(-1, vals.shape[-1]))
        freq_r = iso["freq_r"].values
    else:  # a rank may own nothing when there are more ranks than items
        raise ValueError("every rank must own at least one item of the sharded axis")
    buf = torch.empty(part.numel() + 1, dtype=torch.float64, device=part.device)
    buf[:-1] = part
    buf[-1] = float(n_local)
    allreduce_sum(buf)
    mean = buf[:-1] / buf[-1]
    return DataArray(mean, dims=["freq_r"], coords={"freq_r": freq_r})
