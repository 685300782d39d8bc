import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xrft_b200 import backend as B
for shape, axes, dt in [((64, 360, 720), (1, 2), np.float32), ((64, 360, 720), (1, 2), np.float64), ((8192, 3600), (1,), np.float32), ((4096, 1000), (1,), np.float64), ((64, 1080, 2160), (1, 2), np.float32)]:
    x = np.random.default_rng(1).standard_normal(shape).astype(dt)
    t = torch.from_numpy(x).cuda()
    for _ in range(2): y = B.rfftn(t, axes=list(axes))
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): y = B.rfftn(t, axes=list(axes))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    yb = B.irfftn(y, axes=list(axes))
    e0.record()
    for _ in range(10): yb = B.irfftn(y, axes=list(axes))
    e1.record(); torch.cuda.synchronize()
    msi = e0.elapsed_time(e1) / 10
    err = float((yb - t).abs().max())
    print(f"{shape} axes={axes} {np.dtype(dt).name}: rfftn {ms:.3f} ms ({np.prod(shape)/ms/1e6:.1f} GPoints/s)  irfftn {msi:.3f} ms  round-trip max err {err:.1e}", flush=True)
