"""Worker of tests/test_gpu_configs.py::test_sharded_isotropic_mean_nccl (launched by torch.distributed.run, one rank per
GPU): BASELINE config 4's exchange step -- isotropic_power_spectrum of this rank's chunks on the CUDA kernels, then ONE
all-reduce of nbins + 1 float64 through the C-ABI (xrftb_allreduce_bins -> ncclAllReduce) -- against the un-sharded
oracle mean (xrft/xrft.py:1013-1095 + .mean over the sharded axis, xrft/tests/test_xrft.py:1011-1013)."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

import xrft_b200 as xrft  # noqa: E402
from xrft_b200 import shard  # noqa: E402
from oracle import xrft_oracle as O  # noqa: E402

rng = np.random.default_rng(0)   # the same array on every rank; each takes its own block of `chunk`
nchunk, nz, n = 5, 6, 128
x = (rng.standard_normal((nchunk, nz, n, n)) + 0.3).astype(np.float32)
c = {"chunk": np.arange(nchunk) * 1.0, "z": np.arange(nz) * 1.0, "y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0}
da = xrft.DataArray(torch.from_numpy(x).cuda(), dims=["chunk", "z", "y", "x"], coords=c)
out = shard.sharded_isotropic_mean(da, "chunk", ["y", "x"], detrend="constant", window="hann")
assert shard.last_collective() == ("xrftb_allreduce_bins (NCCL)", world), shard.last_collective()
ref = O.isotropic_power_spectrum(O.Labelled(x.astype(np.float64), ("chunk", "z", "y", "x"), c), dim=["y", "x"], detrend="constant", window="hann")
full = ref.data.reshape(-1, ref.data.shape[-1]).mean(axis=0)
got = out.values
err = np.linalg.norm(got - full) / np.linalg.norm(full)
assert err < 1e-3, err
np.testing.assert_allclose(out["freq_r"].values, ref.coords["freq_r"], rtol=1e-12)
# every rank holds the same reduced result
t = torch.from_numpy(np.ascontiguousarray(got)).cuda()
lo, hi = t.clone(), t.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert torch.equal(lo, hi)
print(f"NCCL_ISO_OK rank={rank}/{world} err={err:.2e}", flush=True)
dist.destroy_process_group()
