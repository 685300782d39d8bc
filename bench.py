#!/usr/bin/env python
"""bench.py -- BASELINE.json headline metric: GPoints/s of detrend + window + FFT + PSD.

Workload (configs[1]): power_spectrum over (y, x) of a (time, 4096, 4096) float32 field, detrend='linear',
window='hann', density scaling, fftshifted full spectrum.  A "step" is one pass of that hot path over the
whole resident dataset (time = 1024 when it fits the GPU's free memory next to its output, else the largest
power of two that does; the value actually used is in config.time).  Points = real-space input grid points.

  value  device-resident inputs (torch CUDA tensor inside the DataArray), CUDA-event timing
  e2e    the same call with HOST (pinned) numpy inputs and a device->host copy of the result, every step
  roofline  achieved HBM bandwidth of the dominant kernel (per-kernel CUDA events inside the timed region)
  cpu_baseline  the oracle (numpy restatement of the reference) on the host cores, bounded sample

`--impl reference` times the reference's CPU path (oracle port; the reference itself cannot be imported in
this image: xarray/dask are absent) on a bounded sample of the same workload.
N > 1 (torchrun): every rank runs the same per-GPU job on its own shard (independent time chunks, no
data-path collective), barrier + max-over-ranks timing, weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "GPoints/s (detrend+window+FFT+PSD)"
UNIT = "GPoints/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--time", type=int, default=1024, help="time slices resident per GPU (config: 1024)")
    ap.add_argument("--ny", type=int, default=4096)
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--e2e-time", type=int, default=64, help="time slices per end-to-end step (host buffers)")
    ap.add_argument("--chunk", type=int, default=32, help="time slices per fused kernel chain (bounds the workspace: 64 MiB per 4096^2 slice)")
    ap.add_argument("--cpu-slices", type=int, default=0, help="slices in the CPU sample (0 = one per worker)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip BASELINE configs 3, 4, 5 (other_configs)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores (emulates dask's threaded scheduler over time chunks)
# ------------------------------------------------------------------------------------------------
def make_slice(seed, ny, nx):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((ny, nx), dtype=np.float32)
    x += (0.3 * np.arange(nx, dtype=np.float32) - 0.7 * np.arange(ny, dtype=np.float32)[:, None] + 5.0)
    return x


def cpu_sample(ny, nx, nslices, workers):
    """Time power_spectrum(detrend='linear', window='hann') of `nslices` slices with `workers` threads."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import xrft_oracle as O

    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # pragma: no cover
        threadpool_limits = None
    slices = [make_slice(1234 + i, ny, nx) for i in range(nslices)]
    coords = {"y": np.arange(ny) * 1.0, "x": np.arange(nx) * 1.0}

    def work(i):
        la = O.Labelled(slices[i][None], ("time", "y", "x"), coords)
        ps = O.power_spectrum(la, dim=["y", "x"], detrend="linear", window="hann")
        return float(ps.data[0, 0, 0])

    import contextlib

    ctx = threadpool_limits(limits=1) if threadpool_limits else contextlib.nullcontext()
    t0 = time.perf_counter()
    with ctx:
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(work, range(nslices)))
    dt = time.perf_counter() - t0
    return nslices * ny * nx / dt / 1e9, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = min(cores, 32)
    ns = args.cpu_slices or workers
    # one slice costs ~4-8 s on one core (dense least-squares detrend).  Every one of the W + K steps runs: the first step is
    # timed, and if W + K steps of that size would not fit the budget the sample per step shrinks (fewer slices per step,
    # the metric is per point) instead of the run being cut short
    vals = []
    budget_s = 200.0
    nsteps = args.warmup + args.steps
    v, dt = cpu_sample(args.ny, args.nx, ns, workers)
    if dt * nsteps > budget_s:
        ns = max(1, int(ns * budget_s / (dt * nsteps)))
        workers = min(workers, ns)
    for i in range(nsteps):
        v, dt = cpu_sample(args.ny, args.nx, ns, workers)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d for _, d in vals])) * 1e3
    sample = f"{ns} slices of {args.ny}x{args.nx} float32 per step, {workers} threads (BLAS threads = 1)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 in / f64 arithmetic (numpy>=2 promotion, SURVEY F6)", "data": "synthetic",
        # the GPU arm's config; a step here is a bounded sample of that workload (the full job would take hours on the host)
        "config": {"workload": "power_spectrum 2-D 4096x4096x1024-time float32, detrend='linear', window='hann', 1 GPU (BASELINE configs[1])",
                   "time": args.time, "ny": args.ny, "nx": args.nx, "slices_per_step": ns,
                   "sample": f"{ns} of {args.time} time slices per step, one slice per worker thread"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 - 0.05 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # timed region shorter than the sampling period: take the nearest samples
            for t, line in self.rows[-3:]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# host link probe (the end-to-end number's denominator)
# ------------------------------------------------------------------------------------------------
def pcie_probe(dev, world, barrier, nbytes=1 << 30):
    """pinned H2D alone, D2H alone, both concurrently (two streams); all ranks at once, slowest rank reported"""
    import torch
    import torch.distributed as dist
    try:
        h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    except Exception:
        return None
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def timed(fns, reps=3):
        for f in fns:
            f()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            for f in fns:
                f()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return nbytes / float(t.item()) / 1e9

    out = {"h2d_GBs": timed([h2d]), "d2h_GBs": timed([d2h]), "both_each_GBs": timed([h2d, d2h]), "ranks_at_once": world}
    del h_in, h_out, d_a, d_b
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import scipy.signal  # noqa: F401  (import cost outside the timed region)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # keep stdout to the single JSON line: NCCL's banner / debug output goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import xrft_b200 as xrft
    from xrft_b200 import backend as B, _lib as L

    lib = L.load()
    ny, nx = args.ny, args.nx
    slice_b = ny * nx * 4
    free, _total = torch.cuda.mem_get_info()
    T = args.time
    while T > 1 and (2 * T * slice_b + (6 << 30)) > free:
        T //= 2
    B.set_fused_chunk(args.chunk)
    # ---- synthetic resident dataset: N(0,1) + 0.3 j - 0.7 i + 5, seeded per chunk
    x = torch.empty((T, ny, nx), dtype=torch.float32, device=dev)
    plane = (0.3 * torch.arange(nx, device=dev, dtype=torch.float32) - 0.7 * torch.arange(ny, device=dev, dtype=torch.float32)[:, None] + 5.0)
    for c0 in range(0, T, 16):
        g = torch.Generator(device=dev).manual_seed(1234 + rank * 100003 + c0)
        x[c0:c0 + 16].normal_(generator=g)
        x[c0:c0 + 16] += plane
    coords = {"time": np.arange(T, dtype=np.float64), "y": np.arange(ny) * 1.0, "x": np.arange(nx) * 1.0}
    da = xrft.DataArray(x, dims=["time", "y", "x"], coords=coords).chunk({"time": args.chunk})

    def step():
        ps = xrft.power_spectrum(da, dim=["y", "x"], detrend="linear", window="hann")
        return ps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    for _ in range(max(args.warmup, 3)):
        del out
        out = step()
    barrier()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.25)
    lib.xrftb_launch_count(1)
    lib.xrftb_profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        del out
        out = step()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    launches = int(lib.xrftb_launch_count(0))
    pms = (ctypes.c_double * 4)()
    pcnt = (ctypes.c_long * 4)()
    lib.xrftb_profile_end(pms, pcnt)
    clk = clocks.stop(t0, t1) if rank == 0 else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    pts_step = T * ny * nx
    value = world * pts_step / (ms_step * 1e-3) / 1e9
    checksum = float(out[0, ny // 2, nx // 2 + 1].values) if rank == 0 else 0.0
    del out

    # ---- roofline of the dominant kernel (achieved algorithmic bytes / CUDA-event duration)
    last_path = int(lib.xrftb_spectrum2d_last_path())
    if last_path == 2:
        # columns-first chain, z mode: pass 1 reads the f32 input and writes the packed column spectra Z (ny x nx/2 c64 =
        # 4 B/point) | pass 2 reads Z and writes the f32 output | completion tables O(nx) per slice | no mirror pass
        names = ["rowline_fix_kernel (column-line completion tables)", "rowsz_power_kernel (pass 2: column separation + row FFT + |F|^2 + mirrored row)",
                 "cols_async_kernel<ColsR2CPack> (pass 1: detrend + window + packed column FFT, TMA in / TMA out)", "mirror_fill_kernel (not launched)"]
        bpp = [0.0, 4.0 + 4.0, 4.0 + 4.0, 0.0]
    elif last_path == 1:
        # columns-first chain: completion tables (O(nx) per slice) | pass 2: read the c64 half spectrum rows, write the f32
        # output | pass 1: read the f32 input, write the c64 half spectrum (ny/2+1 of ny rows) | no mirror pass
        names = ["rowline_fix_kernel (column-line completion tables)", "rows2c_power_kernel (pass 2: row C2C + |F|^2 + mirrored row)",
                 "cols_async_kernel<ColsR2CPack> (pass 1: detrend + window + column R2C)", "mirror_fill_kernel (not launched)"]
        halfy = (ny // 2 + 1) / ny
        bpp = [0.0, 8.0 * halfy + 4.0, 4.0 + 8.0 * halfy, 0.0]
    else:
        names = ["moments_kernel", "rows2_kernel<RowsR2CFused>", "cols_kernel<ColsFused POWER>", "mirror_fill_kernel"]
        # algorithmic bytes per real-space point each kernel must move: read f32 | read f32 + write c64 half spectrum |
        # read half spectrum + write the direct half of the f32 output | read + write the mirrored half
        half = (nx // 2 + 1) / nx
        bpp = [4.0, 4.0 + 8.0 * half, 8.0 * half + 4.0 * half, 4.0 * (1 - half) * 2]
    peak, peak_src = peaks()
    tot_ms = [float(pms[i]) for i in range(4)]
    dom = int(np.argmax(tot_ms)) if sum(tot_ms) > 0 else 1
    n_launch = int(pcnt[dom]) or 1
    pts_launch = pts_step * args.steps / n_launch
    avg_ms = tot_ms[dom] / n_launch if n_launch else float("nan")
    achieved = bpp[dom] * pts_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    # DRAM traffic of the dominant kernel: from the ncu capture in profiles/traffic.json, quoted only while the kernel sources
    # are the ones the capture was taken from (source hash stored beside it); otherwise null
    traffic, traffic_note = None, "no ncu capture for the current kernel sources"
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            import hashlib
            tj = json.load(open(tp))
            h = hashlib.sha256()
            for f in ("fft_core.cuh", "fft_kernels.cuh"):
                h.update(open(os.path.join(ROOT, "xrft_b200", "csrc", f), "rb").read())
            if tj.get("kernel_source_sha16") == h.hexdigest()[:16]:
                bpp_t = tj.get("kernels", {}).get(names[dom], {}).get("dram_bytes_per_point", None)
                if bpp_t is not None:
                    traffic = bpp_t * pts_launch
                    traffic_note = "ncu capture of %s (profiles/traffic.json), same kernel sources" % tj.get("captured")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bpp[dom] * pts_launch, "avg_launch_ms": avg_ms,
                "traffic": traffic, "traffic_source": traffic_note,
                "kernel_share_of_step": {names[i]: tot_ms[i] / (ms_total if ms_total else 1) for i in range(4)},
                "pipeline": {"algorithmic_bytes_per_point": 8.0, "achieved": 8.0 * value / world, "frac": 8.0 * value / world / peak}}

    # ---- end to end: host (pinned) buffers in, result copied back, every step
    e2e = None
    if not args.no_e2e:
        probe = pcie_probe(dev, world, barrier)
        Te = min(args.e2e_time, T)
        hin = torch.empty((Te, ny, nx), dtype=torch.float32).pin_memory()
        hin.copy_(x[:Te])
        hout = torch.empty((Te, ny, nx), dtype=torch.float32).pin_memory()
        hda = xrft.DataArray(hin.numpy(), dims=["time", "y", "x"], coords={"time": np.arange(Te) * 1.0, "y": coords["y"], "x": coords["x"]})

        def e2e_step():
            # host numpy in -> host numpy out: the API streams time chunks (H2D | kernels | D2H overlapped);
            # `out=` is the preallocated pinned result buffer
            ps = xrft.power_spectrum(hda, dim=["y", "x"], detrend="linear", window="hann", out=hout)
            torch.cuda.synchronize()
            return ps

        for _ in range(2):
            e2e_step()
        barrier()
        tt0 = time.perf_counter()
        nst = max(3, min(args.steps, 5))
        for _ in range(nst):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - tt0) / nst
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        each_way = Te * slice_b / float(tt.item()) / 1e9   # GB/s per direction per GPU (slowest rank)
        e2e = {"value": world * Te * ny * nx / float(tt.item()) / 1e9, "unit": UNIT, "h2d_bytes_per_step": Te * slice_b,
               "d2h_bytes_per_step": Te * slice_b, "slices_per_step": Te, "ms_per_step": float(tt.item()) * 1e3,
               "path": "xrft_b200.power_spectrum(DataArray(pinned numpy), out=pinned) -> chunk-streamed H2D | C-ABI kernels | D2H",
               # the limiter: 4 B/point in and 4 B/point out over the host link.  Probe = pinned 1 GiB copies on two streams,
               # every rank at once (min over ranks), measured in this run just before the end-to-end steps
               "link_GBs_each_way": each_way, "pcie_probe": probe,
               "pcie_frac": each_way / probe["both_each_GBs"] if probe and probe.get("both_each_GBs") else None}
        del hin, hout
    del x
    torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        workers = min(cores, 32)
        ns = args.cpu_slices or workers
        v, dt = cpu_sample(ny, nx, ns, workers)
        cpu = {"value": v, "unit": UNIT, "cores": workers, "kind": "port",
               "sample": f"{ns} slices of {ny}x{nx} float32 ({dt:.1f} s), oracle power_spectrum(detrend='linear', window='hann'), {workers} threads"}

    # ---- the other BASELINE configs (3: cross spectrum + phase, 4: isotropic spectrum sharded over the ranks + the NCCL
    #      radial-bin all-reduce, 5: padded rfft/irfft round trip), each with its SURVEY section 8(d) bytes per point
    other = None
    if not args.no_other:
        from tools import bench_configs as BC
        other = {}
        try:
            r4 = BC.run_config4(dev, rank, world)
            if rank == 0:
                other["config4_isotropic_power_spectrum"] = r4
            if rank == 0 and world == 1:
                other["config3_cross_spectrum_phase"] = BC.run_config3(dev)
                other["config5_rfft_irfft_padded_f64"] = BC.run_config5(dev)
        except Exception as exc:  # the headline line must still be printed
            other["error"] = repr(exc)[:500]
        barrier()

    if rank == 0:
        sys.stdout.flush()
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "power_spectrum 2-D 4096x4096x1024-time float32, detrend='linear', window='hann', 1 GPU (BASELINE configs[1])",
                       "time": T, "ny": ny, "nx": nx, "chunk_time": args.chunk, "points_per_step_per_gpu": pts_step,
                       "l2": "inputs (%.1f GiB/step) far exceed the 126 MB L2; no flush needed" % (T * slice_b / 2 ** 30),
                       "parallelism": f"time sharded over {world} GPU(s), no data-path collective"},
            "gpu_launches": launches, "clocks": clk, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
            "other_configs": other, "checksum": checksum,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE line, the JSON: file descriptor 1 is pointed at stderr for the whole run (native libraries -- the
    # NCCL banner -- write to the descriptor, not to sys.stdout) and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    import builtins
    orig_print = builtins.print

    def capture(*a, **k):
        if k.get("file") in (None, sys.stdout):
            lines.append(" ".join(str(x) for x in a))
        else:
            orig_print(*a, **k)

    builtins.print = capture
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    finally:
        builtins.print = orig_print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in lines:
        if ln.startswith("{"):
            orig_print(ln, flush=True)
        else:
            orig_print(ln, file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
