"""Calibration of the HBM roofline by direction: write-only (fill), read-only (sum), copy; CUDA-event timed, 4 GiB buffers."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
def t(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
gb = n * 4 / 1e9
print(f"fill  (write only): {gb / t(lambda: a.fill_(1.5)):.0f} GB/s")
print(f"zero  (memset)    : {gb / t(lambda: a.zero_()):.0f} GB/s")
print(f"sum   (read only) : {gb / t(lambda: a.sum()):.0f} GB/s")
print(f"copy  (r + w)     : {2 * gb / t(lambda: b.copy_(a)):.0f} GB/s (read+write bytes)")
print(f"add_  (r + w same): {2 * gb / t(lambda: a.add_(1.0)):.0f} GB/s (read+write bytes)")
c = a[: n // 2].view(-1, 4096)
print(f"flip  (r + w, reversed rows): {gb / t(lambda: torch.flip(c, dims=(1,))):.0f} GB/s (read+write bytes)")
