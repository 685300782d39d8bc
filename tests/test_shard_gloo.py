"""N > 1 path on CPU: world_size-2 gloo processes exercise the sharder (partition + the one all-reduce).
The per-rank numerics are the oracle here (test infrastructure standing in for the GPU kernels)."""
import os
import socket
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_iso(da, dim, **kw):
    from oracle import xrft_oracle as O
    from xrft_b200 import DataArray

    la = O.Labelled(da.values, da.dims, {d: da[d].values for d in da.dims})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = O.isotropic_power_spectrum(la, dim=dim, **kw)
    return DataArray(r.data, dims=r.dims, coords={"freq_r": r.coords["freq_r"]})


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from xrft_b200 import DataArray
    from xrft_b200 import shard

    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)  # same data on every rank; each takes its own block
    x = rng.standard_normal((5, 3, 32, 32))
    da = DataArray(x, dims=["chunk", "z", "y", "x"], coords={"chunk": np.arange(5.), "z": np.arange(3.), "y": np.arange(32.), "x": np.arange(32.)})
    out = shard.sharded_isotropic_mean(da, "chunk", ["y", "x"], compute=_oracle_iso, detrend="constant", window="hann")
    lo, hi = shard.shard_bounds(5, rank, world)
    q.put((rank, lo, hi, out.values.copy(), out["freq_r"].values.copy()))
    dist.destroy_process_group()


def test_shard_bounds_partition():
    from xrft_b200.shard import shard_bounds
    for n in (1, 5, 8, 64, 1024, 1025):
        for w in (1, 2, 3, 8):
            if n < w:
                continue
            blocks = [shard_bounds(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_world2_gloo_isotropic_allreduce():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 3, 3, 5)
    np.testing.assert_allclose(res[0][3], res[1][3], rtol=0, atol=0)  # all-reduce: identical on both ranks
    # equals the un-sharded answer
    sys.path.insert(0, ROOT)
    from xrft_b200 import DataArray
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 3, 32, 32))
    da = DataArray(x, dims=["chunk", "z", "y", "x"], coords={"chunk": np.arange(5.), "z": np.arange(3.), "y": np.arange(32.), "x": np.arange(32.)})
    full = _oracle_iso(da, ["y", "x"], detrend="constant", window="hann").values.reshape(-1, res[0][3].size).mean(axis=0)
    np.testing.assert_allclose(res[0][3], full, rtol=1e-12)
