import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xrft_b200 import backend as B
x = np.random.default_rng(1).standard_normal((256, 360, 720)).astype(np.float32)
t = torch.from_numpy(x).cuda()
for _ in range(3):
    y = B.rfftn(t, axes=[1, 2])
torch.cuda.synchronize()
print("done")
