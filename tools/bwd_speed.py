"""Per-kernel times of one training step (tuning aid; honours FSNERF_DEBUG_FLAGS)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fsnerf_b200 import ops
from fsnerf_b200.engine import HotPath
dev = torch.device("cuda:0")
hp = HotPath(device=dev)
R = 4096
g = torch.Generator().manual_seed(0)
o = (torch.tensor([0.0, 0, 4.0]) + 0.05 * torch.randn(R, 3, generator=g)).to(dev)
d = torch.nn.functional.normalize(torch.tensor([0.0, 0, -1.0]) + 0.3 * torch.randn(R, 3, generator=g), dim=-1).to(dev)
gt = torch.rand(R, 3, generator=g).to(dev)
for _ in range(3):
    hp.train_step(o, d, gt)
torch.cuda.synchronize()
ops.profile_enable(True)
n = 5
for _ in range(n):
    hp.train_step(o, d, gt)
prof = ops.profile_read()
ops.profile_enable(False)
print(f"FSNERF_DEBUG_FLAGS={os.environ.get('FSNERF_DEBUG_FLAGS','0')}:", {k: round(v[0] / n, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:5]})
