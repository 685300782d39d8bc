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


def allreduce_sum(t):
    """In-place sum over ranks of a torch tensor (no-op when torch.distributed is not initialised)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def sharded_isotropic_mean(da, shard_dim: str, dim: Sequence[str], compute: Optional[Callable] = None, **kwargs) -> DataArray:
    """mean over every non-transform axis of isotropic_power_spectrum(da, dim=...), with `shard_dim` split over
    the ranks: each rank reduces its own block to `nbins` partial sums, then ONE all-reduce (sum + count)."""
    import torch

    if compute is None:
        from .api import isotropic_power_spectrum as compute
    da = from_any(da)
    mine = local_shard(da, shard_dim)
    n_local = 1
    for d in mine.dims:
        if d not in dim:
            n_local *= mine.sizes[d]
    if n_local > 0:
        iso = compute(mine, dim=list(dim), **kwargs)
        vals = iso.data
        if not isinstance(vals, torch.Tensor):
            vals = torch.as_tensor(np.asarray(vals))
        part = vals.reshape(-1, vals.shape[-1]).to(torch.float64).sum(dim=0)
        freq_r = iso["freq_r"].values
    else:  # a rank may own nothing when there are more ranks than items
        raise ValueError("every rank must own at least one item of the sharded axis")
    buf = torch.cat([part, torch.tensor([float(n_local)], dtype=torch.float64, device=part.device)])
    allreduce_sum(buf)
    mean = buf[:-1] / buf[-1]
    return DataArray(mean, dims=["freq_r"], coords={"freq_r": freq_r})
