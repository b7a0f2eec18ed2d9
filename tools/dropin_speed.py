"""ms/step of the reference's train loop on the drop-in modules, with the in-kernel jitter
(default) and with explicit torch.rand uniforms (tuning aid; CPU time per step printed too)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from fsnerf_b200 import ops, synthetic as syn
from fsnerf_b200.core.models import NeRF
from fsnerf_b200.render.rendering import HierarchicalEstimator, render_rays
dev = torch.device("cuda:0")
kw = {"pos_fn": {"n_freqs": 10, "log_space": True}, "dir_fn": {"n_freqs": 4, "log_space": True}}
torch.manual_seed(42)
coarse, fine = NeRF(3, 3, 8, 256, [4], **kw).to(dev), NeRF(3, 3, 8, 256, [4], **kw).to(dev)
est = HierarchicalEstimator(near=2.0, far=6.0, n_coarse=64, n_fine=128, proposal_model=coarse)
opt = torch.optim.Adam(list(fine.parameters()) + list(coarse.parameters()), lr=5e-4)
pose = torch.from_numpy(syn.orbit_poses(4)[1]).to(dev)[None].contiguous()
ro, rd, _ = ops.gen_rays(pose, 400, 400, syn.focal_from_fov(400), first_id=30000, n_rays=4096)
gt = torch.rand(4096, 3, device=dev)
def step(explicit):
    if explicit:
        est.set_uniforms(torch.rand(4096, 64, device=dev), torch.rand(4096, 128, device=dev))
    (rgb, *_, extras), _, _ = render_rays(ro, rd, est, fine, train=True, white_bkgd=True, device=dev)
    loss = F.mse_loss(rgb, gt) + F.mse_loss(extras["rgb_coarse"], gt)
    loss.backward(); opt.step(); opt.zero_grad()
    return loss
for rep in range(2):
    for explicit in (False, True):
        for _ in range(5): step(explicit)
        torch.cuda.synchronize(); t0 = time.perf_counter(); cpu = 0.0
        for _ in range(40):
            c0 = time.perf_counter(); l = step(explicit); cpu += time.perf_counter() - c0
            l.item()
        torch.cuda.synchronize()
        print(f"explicit_u={explicit}: {(time.perf_counter() - t0) * 25:.3f} ms/step, host enqueue {cpu * 25:.3f} ms/step")
