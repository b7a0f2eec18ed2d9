"""Small driver for ncu captures: runs one hot-path piece a few times.
    python tools/prof_kernels.py fwd|train [P]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fsnerf_b200 import ops  # noqa: E402
from fsnerf_b200.engine import HotPath  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dev = torch.device("cuda:0")
ops.require_device(0)
hp = HotPath(device=dev)
R = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
g = torch.Generator().manual_seed(0)
o = (torch.tensor([0.0, 0, 4.0]) + 0.05 * torch.randn(R, 3, generator=g)).to(dev)
d = torch.nn.functional.normalize(torch.tensor([0.0, 0, -1.0]) + 0.3 * torch.randn(R, 3, generator=g), dim=-1).to(dev)
gt = torch.rand(R, 3, generator=g).to(dev)
reps = 3
for i in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if mode == "fwd":
        hp.render(o, d)
    else:
        hp.train_step(o, d, gt)
    torch.cuda.synchronize()
    print(mode, R, "rays:", round((time.perf_counter() - t0) * 1e3, 3), "ms")
