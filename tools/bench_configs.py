#!/usr/bin/env python
"""BASELINE.json configs 3, 4 and 5 through the public API: device-resident synthetic inputs, CUDA-event timing, the
SURVEY.md section 8(d) algorithmic bytes per point, the fraction of the measured HBM peak and a parity check of a
sub-batch against the oracle.  bench.py imports run_config3/4/5 and emits their dicts under `other_configs`; as a script
it prints one JSON line per config (torchrun for config 4 at N > 1):  python tools/bench_configs.py [3,4,5]"""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def _peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def _timeit(fn, reps=3, warm=2, sync=None):
    import torch
    out = None
    for _ in range(warm):
        del out
        out = fn()
    if sync:
        sync()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        del out
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def _relerr(a, b):
    den = np.linalg.norm(np.asarray(b).ravel())
    return float(np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / (den if den > 0 else 1.0))


def _cpu_threads(work, nitems):
    """time `work(i)` for i in range(nitems) on a thread pool of the host cores (BLAS threads = 1): the reference's
    chunk-parallel CPU model (dask threaded scheduler over independent chunks).  Returns (seconds, threads)."""
    import contextlib
    import time
    from concurrent.futures import ThreadPoolExecutor
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=1)
    except Exception:  # pragma: no cover
        ctx = contextlib.nullcontext()
    workers = max(1, min(os.cpu_count() or 1, 32, nitems))
    t0 = time.perf_counter()
    with ctx:
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(work, range(nitems)))
    return time.perf_counter() - t0, workers


def _roof(gpts, bytes_per_point):
    peak, src = _peak()
    gbs = gpts * bytes_per_point
    return {"algorithmic_bytes_per_point": bytes_per_point, "achieved_GBs": gbs, "frac_of_hbm_peak": gbs / peak, "peak_GBs": peak, "peak_source": src}


def run_config3(dev, T=512, n=2048, parity=True):
    """cross_spectrum + cross_phase of two 2048^2 x 512 float32 fields, detrend='constant', window='hann' (xrft.py:753-874)"""
    import torch
    import xrft_b200 as xrft
    from xrft_b200 import _lib as L
    warnings.simplefilter("ignore")
    lib = L.load()
    g = torch.Generator(device=dev).manual_seed(1)
    a = torch.randn((T, n, n), generator=g, device=dev)
    b = torch.roll(a, shifts=(3, 5), dims=(1, 2))
    b += 0.5 * torch.randn((T, n, n), generator=g, device=dev)
    c = {"time": np.arange(T) * 1.0, "y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0}
    da, db = xrft.DataArray(a, dims=["time", "y", "x"], coords=c), xrft.DataArray(b, dims=["time", "y", "x"], coords=c)
    kw = dict(dim=["y", "x"], detrend="constant", window="hann")
    pts = T * n * n
    lib.xrftb_launch_count(1)
    ms_cs, cs = _timeit(lambda: xrft.cross_spectrum(da, db, **kw))
    res = {"config": 3, "workload": "cross_spectrum + cross_phase, two %d^2 x %d float32 fields, detrend=constant, window=hann" % (n, T),
           "points_per_field": pts, "unit": "GPoints/s per field"}
    par = {}
    if parity:
        from oracle import xrft_oracle as O
        c1 = {k: (v[:1] if k == "time" else v) for k, v in c.items()}
        ref = O.cross_spectrum(O.Labelled(a[:1].cpu().numpy().astype(np.float64), ("time", "y", "x"), c1),
                               O.Labelled(b[:1].cpu().numpy().astype(np.float64), ("time", "y", "x"), c1), **kw).data
        par["cross_spectrum_relerr_slice0"] = _relerr(cs.data[:1].cpu().numpy(), ref)
    del cs
    ms_cp, cp = _timeit(lambda: xrft.cross_phase(da, db, **kw))
    if parity:
        d = np.abs(np.angle(np.exp(1j * (cp.data[:1].cpu().numpy() - np.angle(ref)))))
        par["cross_phase_max_abs_err_slice0"] = float(d[np.abs(ref) > 1e-4 * np.abs(ref).max()].max())
        del ref
    del cp
    res.update({"cross_spectrum": dict(value=pts / ms_cs / 1e6, ms=ms_cs, **_roof(pts / ms_cs / 1e6, 16.0)),
                "cross_phase": dict(value=pts / ms_cp / 1e6, ms=ms_cp, **_roof(pts / ms_cp / 1e6, 12.0))})
    if hasattr(xrft, "cross_spectrum_and_phase"):
        ms_b, both = _timeit(lambda: xrft.cross_spectrum_and_phase(da, db, **kw))
        del both
        res["cross_spectrum_and_phase"] = dict(value=pts / ms_b / 1e6, ms=ms_b, **_roof(pts / ms_b / 1e6, 20.0))
    res["gpu_launches"] = int(lib.xrftb_launch_count(0))
    res["parity"] = par
    if parity:   # the reference's CPU path (oracle port) on a bounded sample: one slice pair per worker thread
        from oracle import xrft_oracle as O
        ns = max(1, min(os.cpu_count() or 1, 16, T))
        ah, bh = a[:ns].cpu().numpy(), b[:ns].cpu().numpy()
        c1 = {k: (v[:1] if k == "time" else v) for k, v in c.items()}

        def work(i):
            la = O.Labelled(ah[i:i + 1].astype(np.float64), ("time", "y", "x"), c1)
            lb = O.Labelled(bh[i:i + 1].astype(np.float64), ("time", "y", "x"), c1)
            O.cross_spectrum(la, lb, **kw); O.cross_phase(la, lb, **kw)

        sec, thr = _cpu_threads(work, ns)
        res["cpu_baseline"] = {"value": ns * n * n / sec / 1e9, "unit": "GPoints/s per field (cross_spectrum + cross_phase)", "cores": thr, "kind": "port",
                               "sample": "%d slice pairs of %d^2 float32 (%.1f s)" % (ns, n, sec)}
    del a, b, da, db
    torch.cuda.empty_cache()
    return res


def run_config4(dev, rank=0, world=1, chunks=64, z=512, n=512, parity=True):
    """isotropic_power_spectrum of 512^2 planes x (64 chunks x 512 z) float32, detrend='constant', window='hann', mean
    over (chunk, z); `chunk` sharded over the ranks, ONE all-reduce of nbins + 1 float64 (xrft.py:1013-1095)."""
    import torch
    import xrft_b200 as xrft
    from xrft_b200 import shard, _lib as L
    warnings.simplefilter("ignore")
    lib = L.load()
    lo, hi = shard.shard_bounds(chunks, rank, world)
    x = torch.empty((hi - lo, z, n, n), dtype=torch.float32, device=dev)
    for i in range(hi - lo):   # seeded per chunk: the same global array whatever the number of ranks
        g = torch.Generator(device=dev).manual_seed(100 + lo + i)
        x[i].normal_(generator=g)
    x += 0.25
    c = {"chunk": np.arange(lo, hi) * 1.0, "z": np.arange(z) * 1.0, "y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0}
    da = xrft.DataArray(x, dims=["chunk", "z", "y", "x"], coords=c)
    kw = dict(detrend="constant", window="hann")

    def step():
        # this rank's block is `da` itself: shard_dim of length (hi - lo) split over one rank
        return shard.sharded_isotropic_mean(da, "chunk", ["y", "x"], presharded=True, **kw)

    def sync():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    lib.xrftb_launch_count(1)
    ms, mean = _timeit(step, sync=sync)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    pts = chunks * z * n * n
    gpts = pts / ms / 1e6
    res = {"config": 4, "workload": "isotropic_power_spectrum %d^2 planes x (%d x %d) float32, detrend=constant, window=hann, mean over (chunk, z)" % (n, chunks, z),
           "n_gpus": world, "scaling": "strong (64 chunks split over the ranks)", "value": gpts, "unit": "GPoints/s (whole job)", "ms": ms,
           "points_total": pts, "nbins": int(mean.shape[-1]), "collective": "%s over %d rank(s)" % shard.last_collective(),
           "gpu_launches": int(lib.xrftb_launch_count(0))}
    res.update(_roof(gpts / world, 4.0))
    if parity and rank == 0:
        from oracle import xrft_oracle as O
        sub = x[0, :4].cpu().numpy().astype(np.float64)
        cc = {"z": c["z"][:4], "y": c["y"], "x": c["x"]}
        ref = O.isotropic_power_spectrum(O.Labelled(sub, ("z", "y", "x"), cc), dim=["y", "x"], **kw).data
        got = xrft.isotropic_power_spectrum(xrft.DataArray(x[0, :4], dims=["z", "y", "x"], coords=cc), dim=["y", "x"], **kw).values
        res["parity"] = {"iso_relerr_4_planes": _relerr(got, ref)}
        if world == 1:
            npl = 4 * max(1, min(os.cpu_count() or 1, 32))
            xh = x.reshape(-1, n, n)[:npl].cpu().numpy()
            c4 = {"z": np.arange(4.0), "y": c["y"], "x": c["x"]}

            def work(i):
                O.isotropic_power_spectrum(O.Labelled(xh[4 * i:4 * i + 4].astype(np.float64), ("z", "y", "x"), c4), dim=["y", "x"], **kw)

            sec, thr = _cpu_threads(work, npl // 4)
            res["cpu_baseline"] = {"value": npl * n * n / sec / 1e9, "unit": "GPoints/s", "cores": thr, "kind": "port",
                                   "sample": "%d planes of %d^2 float32 (%.1f s)" % (npl, n, sec)}
    del x, da
    torch.cuda.empty_cache()
    return res


def run_config5(dev, n=8192, p=4096, cpu=True):
    """pad -> fft(real_dim) -> ifft(real_dim) -> unpad of 8192^2 float64 (padded grid 16384^2) + Parseval (padding.py, xrft.py:307-646)"""
    import torch
    import xrft_b200 as xrft
    from xrft_b200 import _lib as L
    warnings.simplefilter("ignore")
    lib = L.load()
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn((n, n), generator=g, device=dev, dtype=torch.float64)
    da = xrft.DataArray(x, dims=["y", "x"], coords={"y": np.arange(n) * 0.5, "x": np.arange(n) * 0.5})

    def rt():
        padded = xrft.pad(da, x=p, y=p)
        ft = xrft.fft(padded, real_dim="x")
        back = xrft.ifft(ft, real_dim="freq_x")
        un = xrft.unpad(back, {"x": p, "y": p})
        un.data   # pad and ifft are deferred until their data is asked for: the round trip is computed here
        return un

    lib.xrftb_launch_count(1)
    ms, un = _timeit(rt, reps=3, warm=2)
    launches = int(lib.xrftb_launch_count(0))
    err = float((un.data - x).abs().max() / x.abs().max())
    del un
    padded = xrft.pad(da, x=p, y=p)
    ps = xrft.power_spectrum(padded, real_dim="x")
    pars = float(ps.data.sum()) * ps["freq_x"].attrs["spacing"] * ps["freq_y"].attrs["spacing"]
    ref = float((padded.data ** 2).mean())
    N = n + 2 * p
    gpts = N * N / ms / 1e6
    res = {"config": 5, "workload": "pad(%d) -> rfft -> irfft -> unpad of %d^2 float64 (padded grid %d^2)" % (p, n, N), "ms_round_trip": ms,
           "value": gpts, "unit": "GPoints/s (padded-grid points)", "gpu_launches": launches,
           "parity": {"round_trip_max_rel_err": err, "parseval_rel_err": abs(pars - ref) / ref}}
    res.update(_roof(gpts, 32.0))
    if cpu:
        import time
        from oracle import xrft_oracle as O
        nh, ph = min(n, 4096), min(p, 2048)     # bounded sample: a quarter-size grid, scaled by points
        xh = x[:nh, :nh].cpu().numpy()
        la = O.Labelled(xh, ("y", "x"), {"y": np.arange(nh) * 0.5, "x": np.arange(nh) * 0.5})
        t0 = time.perf_counter()
        pd = O.pad(la, {"x": ph, "y": ph})
        back = O.unpad(O.ifft(O.fft(pd, real_dim="x"), real_dim="freq_x"), {"x": ph, "y": ph})
        sec = time.perf_counter() - t0
        Nh = nh + 2 * ph
        res["cpu_baseline"] = {"value": Nh * Nh / sec / 1e9, "unit": "GPoints/s (padded-grid points)", "cores": 1, "kind": "port",
                               "sample": "pad(%d) -> fft -> ifft -> unpad of %d^2 float64 (%.1f s, numpy pocketfft, one array = one thread)" % (ph, nh, sec),
                               "round_trip_max_rel_err": float(np.abs(back.data - xh).max() / np.abs(xh).max())}
    del x, da, padded, ps
    torch.cuda.empty_cache()
    return res


def main():
    import torch
    which = sys.argv[1] if len(sys.argv) > 1 else "3,4,5"
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if "3" in which and rank == 0:
        print(json.dumps(run_config3(dev)), flush=True)
    if "4" in which:
        r = run_config4(dev, rank, world)
        if rank == 0:
            print(json.dumps(r), flush=True)
    if "5" in which and rank == 0:
        print(json.dumps(run_config5(dev)), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
