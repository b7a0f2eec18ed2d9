"""Compositing kernels at render scale vs the HBM roofline (tuning aid)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fsnerf_b200 import ops
dev = torch.device("cuda:0")
Rc, S = 262144, 192
g = torch.Generator(device=dev).manual_seed(0)
raw = torch.rand(Rc, S, 4, device=dev, generator=g)
e = torch.sort(2 + 4 * torch.rand(Rc, S + 1, device=dev, generator=g), -1).values
ts, te = e[:, :-1].contiguous(), e[:, 1:].contiguous()
d_rgb = torch.rand(Rc, 3, device=dev, generator=g)
bk = torch.ones(3, device=dev)
for _ in range(3):
    ops.composite_forward(raw, ts, te, bkgd=bk); ops.composite_backward(raw, ts, te, d_rgb, bkgd=bk)
torch.cuda.synchronize()
ops.profile_enable(True)
for _ in range(10):
    ops.composite_forward(raw, ts, te, bkgd=bk); ops.composite_backward(raw, ts, te, d_rgb, bkgd=bk)
p = ops.profile_read(); ops.profile_enable(False)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6467.1
f, b = p["composite_fwd"][0] / 10, p["composite_bwd"][0] / 10
print(f"fwd {f:.3f} ms {Rc*(28*S+20)/f/1e6:.0f} GB/s ({Rc*(28*S+20)/f/1e6/peak:.2f}) | bwd {b:.3f} ms {Rc*(40*S+40)/b/1e6:.0f} GB/s ({Rc*(40*S+40)/b/1e6/peak:.2f})")
