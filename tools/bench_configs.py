#!/usr/bin/env python
"""Secondary BASELINE.json configs (3, 4, 5) through the public API; prints one JSON line per config.
(bench.py keeps the driver contract for the headline config 2.)  Run under torchrun for config 4 at N > 1."""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xrft_b200 as xrft
from xrft_b200 import shard

warnings.simplefilter("ignore")
which = sys.argv[1] if len(sys.argv) > 1 else "3,4,5"
if len(sys.argv) > 2:   # fused-chain chunk (batch items per kernel chain)
    from xrft_b200 import backend as _B
    _B.set_fused_chunk(int(sys.argv[2]))
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def timeit(fn, reps=3, warm=2):
    for _ in range(warm): out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def emit(d):
    if rank == 0: print(json.dumps(d), flush=True)


if "3" in which:  # cross_spectrum + cross_phase of two 2048^2 x 512 fields
    T, n = 512, 2048
    g = torch.Generator(device=dev).manual_seed(1)
    a = torch.randn((T, n, n), generator=g, device=dev)
    b = torch.roll(a, shifts=(3, 5), dims=(1, 2)) + 0.5 * torch.randn((T, n, n), generator=g, device=dev)
    c = {"time": np.arange(T) * 1.0, "y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0}
    da, db = xrft.DataArray(a, dims=["time", "y", "x"], coords=c), xrft.DataArray(b, dims=["time", "y", "x"], coords=c)
    ms_cs, cs = timeit(lambda: xrft.cross_spectrum(da, db, dim=["y", "x"], detrend="constant", window="hann"))
    del cs
    ms_cp, cp = timeit(lambda: xrft.cross_phase(da, db, dim=["y", "x"], detrend="constant", window="hann"))
    pts = T * n * n
    emit({"config": 3, "workload": "cross_spectrum + cross_phase, two 2048^2 x 512 float32 fields, detrend=constant, window=hann",
          "cross_spectrum_GPts_s": pts / ms_cs / 1e6, "cross_phase_GPts_s": pts / ms_cp / 1e6, "both_GPts_s": pts / (ms_cs + ms_cp) / 1e6,
          "ms": [ms_cs, ms_cp], "points_per_field": pts, "phase_range": [float(cp.data.min()), float(cp.data.max())]})
    del a, b, da, db, cp
    torch.cuda.empty_cache()

if "4" in which:  # isotropic_power_spectrum of 512^2 planes x (64 chunks x 512 z), sharded over ranks, one all-reduce
    chunks, z, n = 64, 512, 512
    lo, hi = shard.shard_bounds(chunks, rank, world)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.randn((hi - lo, z, n, n), generator=g, device=dev)
    c = {"chunk": np.arange(lo, hi) * 1.0, "z": np.arange(z) * 1.0, "y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0}
    da = xrft.DataArray(x, dims=["chunk", "z", "y", "x"], coords=c)

    def step():
        iso = xrft.isotropic_power_spectrum(da, dim=["y", "x"], detrend="constant", window="hann")
        part = iso.data.reshape(-1, iso.shape[-1]).double().sum(dim=0)
        buf = torch.cat([part, torch.tensor([float(iso.data.numel() // iso.shape[-1])], dtype=torch.float64, device=dev)])
        shard.allreduce_sum(buf)
        return buf[:-1] / buf[-1]

    if world > 1: dist.barrier()
    ms, mean = timeit(step)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pts = chunks * z * n * n  # whole job
    local_pts = (hi - lo) * z * n * n
    emit({"config": 4, "workload": "isotropic_power_spectrum 512^2 planes x (64 x 512), detrend=constant, window=hann, mean over (chunk, z)",
          "n_gpus": world, "GPts_s": (local_pts * world) / float(t.item()) / 1e6, "ms": float(t.item()), "points_total": local_pts * world,
          "nbins": int(mean.numel()), "collective": "one all-reduce of nbins+1 float64" if world > 1 else "none (1 rank)",
          "mean_head": [float(v) for v in mean[:3]]})
    del x, da
    torch.cuda.empty_cache()

if "5" in which:  # pad -> rfft -> irfft -> unpad, float64, + Parseval
    n = int(os.environ.get("C5_N", "4096")); p = int(os.environ.get("C5_PAD", str(n // 2)))
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn((n, n), generator=g, device=dev, dtype=torch.float64)
    da = xrft.DataArray(x, dims=["y", "x"], coords={"y": np.arange(n) * 0.5, "x": np.arange(n) * 0.5})

    def rt():
        padded = xrft.pad(da, x=p, y=p)
        ft = xrft.fft(padded, real_dim="x")
        back = xrft.ifft(ft, real_dim="freq_x")
        return xrft.unpad(back, {"x": p, "y": p}), padded

    ms, (un, padded) = timeit(rt, reps=2, warm=1)
    err = float((un.data - x).abs().max() / x.abs().max())
    ps = xrft.power_spectrum(padded, real_dim="x")
    pars = float(ps.data.sum()) * ps["freq_x"].attrs["spacing"] * ps["freq_y"].attrs["spacing"]
    ref = float((padded.data ** 2).mean())
    N = n + 2 * p
    emit({"config": 5, "workload": f"pad({p}) -> rfft -> irfft -> unpad of {n}^2 float64 (padded grid {N}^2)", "ms_round_trip": ms,
          "GPts_s_padded_grid": N * N / ms / 1e6, "round_trip_max_rel_err": err, "parseval_rel_err": abs(pars - ref) / ref})
if world > 1:
    dist.destroy_process_group()
