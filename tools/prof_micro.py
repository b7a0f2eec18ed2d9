"""Driver for ncu captures of the bandwidth-bound kernels at render scale (262 144 rays: every launch
moves far more than the 126 MB L2): ray generation, stratified sampling, sample_pdf (random and
deterministic u), compositing forward / backward.  Two repetitions; profile the second."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fsnerf_b200 import ops, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
ops.require_device(0)
R, Sc, Sf = 262144, 64, 128
g = torch.Generator(device=dev).manual_seed(0)
pose = torch.from_numpy(syn.orbit_poses(4)[1]).to(dev)[None].contiguous()
H = W = 800
focal = syn.focal_from_fov(W)
us, up = torch.rand(R, Sc, device=dev, generator=g), torch.rand(R, Sf, device=dev, generator=g)
w = torch.rand(R, Sc, device=dev, generator=g) ** 4
raw = torch.rand(R, Sc + Sf, 4, device=dev, generator=g)
e = torch.sort(2 + 4 * torch.rand(R, Sc + Sf + 1, device=dev, generator=g), -1).values
ts, te = e[:, :-1].contiguous(), e[:, 1:].contiguous()
d_rgb = torch.rand(R, 3, device=dev, generator=g)
bk = torch.ones(3, device=dev)
for rep in range(2):
    ops.gen_rays(pose, H, W, focal, first_id=0, n_rays=R)
    tsc, _ = ops.sample_stratified(R, Sc, 2.0, 6.0, us)
    ops.sample_pdf(tsc, w, Sf, 6.0, up, want_aux=False)
    ops.sample_pdf(tsc, w, Sf, 6.0, None, want_aux=False)
    ops.composite_forward(raw, ts, te, bkgd=bk)
    ops.composite_backward(raw, ts, te, d_rgb, bkgd=bk)
    torch.cuda.synchronize()
print("ok")
