"""Quick perf probe of the fused config-2 path (not the bench contract)."""
import sys, time
import numpy as np, torch, scipy.signal as sps
sys.path.insert(0, ".")
from xrft_b200 import backend as B, _lib as L

ny = nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dt = torch.float32 if (len(sys.argv) <= 4 or sys.argv[4] == "f32") else torch.float64
g = torch.Generator(device="cuda").manual_seed(1234)
x = torch.randn((T, ny, nx), generator=g, device="cuda", dtype=dt)
x += 0.3 * torch.arange(nx, device="cuda", dtype=dt) - 0.7 * torch.arange(ny, device="cuda", dtype=dt)[:, None] + 5
wy = torch.from_numpy(sps.windows.hann(ny, sym=False)); wx = torch.from_numpy(sps.windows.hann(nx, sym=False))
chunks = [int(c) for c in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]
mode = sys.argv[6] if len(sys.argv) > 6 else "power"
x2 = torch.roll(x, shifts=(3, 5), dims=(1, 2)) + 0.5 if mode != "power" else None
MODE = {"power": L.EPI_POWER, "cross": L.EPI_CROSS, "phase": L.EPI_PHASE}[mode]
lib = L.load()
for ch in chunks:
    mw = (2 << 30) if ch == 0 else lib.xrftb_spectrum2d_workspace(0 if dt == torch.float32 else 1, ny, nx, 0 if mode == "power" else 1, ch)
    def run():
        return B.spectrum2d(x, x2, MODE, detrend=2, win_y=wy, win_x=wx, shift_y=True, shift_x=True, scale=1.0 / (ny * nx), max_work_bytes=mw)
    for _ in range(2): out = run()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pts = T * ny * nx
    import ctypes
    lib.xrftb_profile_begin(); out = run(); pm = (ctypes.c_double * 4)(); pc = (ctypes.c_long * 4)(); lib.xrftb_profile_end(pm, pc)
    err = ""
    if mode == "power":
        # accuracy of slice 0 against a float64 numpy evaluation of the reference formulas (closed-form plane, f32 round, window, fft2)
        xd = x[0].double().cpu().numpy()
        ii = np.arange(ny)[:, None] - 0.5 * (ny - 1); jj = np.arange(nx)[None, :] - 0.5 * (nx - 1)
        pl = xd.mean() + ii * ((ii * xd).sum() / (nx * ny * (ny * ny - 1) / 12)) + jj * ((jj * xd).sum() / (ny * nx * (nx * nx - 1) / 12))
        d = (xd - pl).astype(np.float32 if dt == torch.float32 else np.float64).astype(np.float64)
        ref = np.fft.fftshift(np.abs(np.fft.fft2(d * wy.numpy()[:, None] * wx.numpy()[None, :])) ** 2) / (ny * nx)
        got = out[0].double().cpu().numpy()
        err = f" relL2={np.linalg.norm(got - ref) / np.linalg.norm(ref):.2e} maxrel={np.abs(got - ref).max() / ref.max():.2e}"
    print(f"{ny}x{nx}x{T} {dt} chunk={ch}:{err} {ms:.3f} ms/step  {pts / ms / 1e6:.1f} GPts/s | us/16slices-equiv: " + " ".join(f"{n}={pm[i] * 1e3 * 16 / T:.0f}" for i, n in enumerate(["mom", "rows", "cols", "mirror"])))
