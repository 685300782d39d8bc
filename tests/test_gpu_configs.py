"""BASELINE.json configs 3, 4 and 5 at (or near) their full sizes, CUDA path vs the oracle on the same seeded inputs,
plus the size-independent properties of each path.  Tolerances are north_star's: 1e-3 relative for float32, 1e-6 for
float64 (normwise).  Reference call sites: xrft/xrft.py:753-874 (cross spectrum / phase), :948-1095 (isotropic power
spectrum), :307-476 / :479-646 (fft / ifft with real_dim), xrft/padding.py:157-181, 394-446 (pad / unpad)."""
import os
import subprocess
import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import xrft_b200 as xrft  # noqa: E402
from xrft_b200 import DataArray  # noqa: E402
from oracle import xrft_oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
warnings.simplefilter("ignore")


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0)


def synthetic_field(N, rng, slope=-3.0, amp=10.0):
    """power-law field of the reference's isotropic tests (xrft/tests/test_xrft.py:845-914)"""
    k = np.fft.fftshift(np.fft.fftfreq(N, 1.0))
    kk, ll = np.meshgrid(k, k)
    K = np.sqrt(kk ** 2 + ll ** 2)
    with np.errstate(divide="ignore"):
        spec = np.where(K > 0, amp * K ** (slope - 1.0), 0.0)
    F = np.sqrt(spec) * np.exp(1j * rng.uniform(-np.pi, np.pi, (N, N)))
    return np.real(np.fft.ifft2(np.fft.ifftshift(F))) * N


# ------------------------------------------------------------------------------------------- config 4
@pytest.mark.parametrize("detrend", ["constant", "linear", None])
def test_config4_isotropic_power_spectrum_512_f32(detrend):
    """512^2 float32 planes, batch 96 (> the 64 of one chunk), hann: every plane's radial spectrum vs the oracle."""
    import torch
    rng = np.random.default_rng(4)
    n, B = 512, 96
    base = np.stack([synthetic_field(n, rng) for _ in range(4)])
    x = (base[np.arange(B) % 4] + 0.2 * rng.standard_normal((B, n, n)) + 0.01 * np.arange(n) + 0.5).astype(np.float32)
    x = x.reshape(3, 32, n, n)
    c = {"chunk": np.arange(3.0), "z": np.arange(32.0), "y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0}
    kw = dict(dim=["y", "x"], detrend=detrend, window="hann")
    out = xrft.isotropic_power_spectrum(DataArray(torch.from_numpy(x).cuda(), dims=["chunk", "z", "y", "x"], coords=c), **kw)
    ref = O.isotropic_power_spectrum(O.Labelled(x.astype(np.float64), ("chunk", "z", "y", "x"), c), **kw)
    assert out.dims == ref.dims == ("chunk", "z", "freq_r") and out.shape == (3, 32, 128)
    np.testing.assert_allclose(out["freq_r"].values, ref.coords["freq_r"], rtol=1e-12)
    got = out.values.astype(np.float64)
    per_plane = np.linalg.norm(got - ref.data, axis=-1) / np.linalg.norm(ref.data, axis=-1)
    assert per_plane.max() < 1e-3, per_plane.max()
    # the k^-3 spectrum spans six decades: bins far below the norm must be right too (float32: 1e-3 of each bin)
    np.testing.assert_allclose(got, ref.data, rtol=2e-3, atol=1e-9 * ref.data.max())
    # sum conservation (xrft/tests/test_xrft.py:963): radial sum == sum of the 2-D spectrum
    ps = xrft.power_spectrum(DataArray(torch.from_numpy(x[0, :4]).cuda(), dims=["z", "y", "x"], coords={"z": c["z"][:4], "y": c["y"], "x": c["x"]}), **kw)
    np.testing.assert_allclose(got[0, :4].sum(-1), ps.values.astype(np.float64).sum((-2, -1)), rtol=1e-4)


def test_config4_host_input_and_f64():
    """the same path from host (numpy) input and in float64 (1e-6)"""
    rng = np.random.default_rng(41)
    n = 512
    x = np.stack([synthetic_field(n, rng) for _ in range(3)]) + 0.1 * rng.standard_normal((3, n, n))
    c = {"t": np.arange(3.0), "y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0}
    kw = dict(dim=["y", "x"], detrend="constant", window="hann")
    ref = O.isotropic_power_spectrum(O.Labelled(x, ("t", "y", "x"), c), **kw)
    out = xrft.isotropic_power_spectrum(DataArray(x, dims=["t", "y", "x"], coords=c), **kw)
    assert relerr(out.values, ref.data) < 1e-6
    out32 = xrft.isotropic_power_spectrum(DataArray(x.astype(np.float32), dims=["t", "y", "x"], coords=c), **kw)
    ref32 = O.isotropic_power_spectrum(O.Labelled(x.astype(np.float32).astype(np.float64), ("t", "y", "x"), c), **kw)
    assert relerr(out32.values, ref32.data) < 1e-3


# ------------------------------------------------------------------------------------------- config 3
def test_config3_cross_spectrum_and_phase_2048():
    """two 2048^2 x 4 float32 fields, detrend='constant', window='hann' (config 3's call), vs the oracle"""
    import torch
    T, n = 4, 2048
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn((T, n, n), generator=g, device="cuda") + 1.5
    b = torch.roll(a, shifts=(3, 5), dims=(1, 2)) + 0.5 * torch.randn((T, n, n), generator=g, device="cuda") - 0.5
    c = {"time": np.arange(T) * 1.0, "y": np.arange(n) * 1.0, "x": np.arange(n) * 1.0}
    da, db = DataArray(a, dims=["time", "y", "x"], coords=c), DataArray(b, dims=["time", "y", "x"], coords=c)
    kw = dict(dim=["y", "x"], detrend="constant", window="hann")
    la = O.Labelled(a.cpu().numpy().astype(np.float64), ("time", "y", "x"), c)
    lb = O.Labelled(b.cpu().numpy().astype(np.float64), ("time", "y", "x"), c)
    ref = O.cross_spectrum(la, lb, **kw)
    cs = xrft.cross_spectrum(da, db, **kw)
    assert cs.dims == ref.dims
    for t in range(T):
        assert relerr(cs.values[t], ref.data[t]) < 1e-3
    ph = xrft.cross_phase(da, db, **kw)
    d = np.abs(np.angle(np.exp(1j * (ph.values - np.angle(ref.data)))))
    sig = np.abs(ref.data) > 1e-4 * np.abs(ref.data).max()
    # float32: the angle of a cell is as accurate as its cross spectrum relative to its own magnitude
    assert d[sig].max() < 5e-3, d[sig].max()
    assert np.abs(ph.values).max() <= np.pi + 1e-6
    # C(-k) = conj(C(k)); phase antisymmetric
    v = cs.values[0]
    np.testing.assert_allclose(v[1:, 1:], np.conj(v[1:, 1:][::-1, ::-1]), rtol=1e-4, atol=1e-6 * np.abs(v).max())
    # both outputs from one pass over the two fields: identical to the two separate calls
    cs2, ph2 = xrft.cross_spectrum_and_phase(da, db, **kw)
    assert cs2.dims == cs.dims and ph2.dims == ph.dims
    np.testing.assert_array_equal(cs2.values, cs.values)
    np.testing.assert_array_equal(ph2.values, ph.values)
    # cross spectrum of a field with itself is its power spectrum
    ps = xrft.power_spectrum(da, **kw)
    cc = xrft.cross_spectrum(da, da, **kw)
    assert relerr(cc.values.real, ps.values) < 1e-5 and np.abs(cc.values.imag).max() <= 1e-6 * np.abs(ps.values).max()


# ------------------------------------------------------------------------------------------- config 5
def test_config5_round_trip_8192_padded_f64():
    """pad(4096) -> fft(real_dim) -> ifft(real_dim) -> unpad of 8192^2 float64 (padded grid 16384^2): round trip 1e-6,
    Parseval, and the forward half spectrum against numpy's rfft2 (the reference's backend) on the full grid."""
    import torch
    n, p = 8192, 4096
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64)
    da = DataArray(x, dims=["y", "x"], coords={"y": np.arange(n) * 0.5, "x": np.arange(n) * 0.5})
    padded = xrft.pad(da, x=p, y=p)
    assert padded.shape == (n + 2 * p, n + 2 * p) and padded["x"].attrs["pad_width"] == p
    ft = xrft.fft(padded, real_dim="x")
    assert ft.shape == (n + 2 * p, (n + 2 * p) // 2 + 1)
    back = xrft.ifft(ft, real_dim="freq_x")
    un = xrft.unpad(back, {"x": p, "y": p})
    assert un.shape == (n, n)
    err = float((un.data - x).abs().max() / x.abs().max())
    assert err < 1e-6, err
    np.testing.assert_allclose(un["x"].values, da["x"].values, atol=1e-9)
    # Parseval: sum |X|^2 dk dl (one-sided weights) == mean(x^2) per unit area -> power_spectrum(real_dim) sums to it
    ps = xrft.power_spectrum(padded, real_dim="x")
    e_spec = float(ps.data.sum()) * ps["freq_x"].attrs["spacing"] * ps["freq_y"].attrs["spacing"]
    np.testing.assert_allclose(e_spec, float((padded.data ** 2).mean()), rtol=1e-10)
    del ps, back, un
    # forward parity on the full padded grid: true_phase ramp + dx dy + fftshift(y) applied to numpy's rfft2 by the oracle
    xp = np.zeros((n + 2 * p, n + 2 * p))
    xp[p:p + n, p:p + n] = x.cpu().numpy()
    cy = padded["y"].values; cx = padded["x"].values
    ref = O.fft(O.Labelled(xp, ("y", "x"), {"y": cy, "x": cx}), real_dim="x")
    got = ft.values
    assert relerr(got, ref.data) < 1e-6, relerr(got, ref.data)


# ------------------------------------------------------------------- multi-GPU: CUDA kernels + NCCL all-reduce
def test_sharded_isotropic_mean_nccl():
    """torchrun-spawned ranks (2 when the box has >= 2 GPUs, else 1): sharded_isotropic_mean with the CUDA kernels and
    the NCCL all-reduce behind the C-ABI (xrftb_allreduce_bins) equals the un-sharded oracle mean on every rank."""
    import torch
    nproc = 2 if torch.cuda.device_count() >= 2 else 1
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_iso_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert r.stdout.count("NCCL_ISO_OK") == nproc, r.stdout[-2000:]
