"""CPU-side profile (cProfile) of the reference's train loop on the drop-in modules (tuning aid)."""
import cProfile, os, pstats, sys, time, torch
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
def step(sync):
    (rgb, *_, extras), _, _ = render_rays(ro, rd, est, fine, train=True, white_bkgd=True, device=dev)
    loss = F.mse_loss(rgb, gt) + F.mse_loss(extras["rgb_coarse"], gt)
    loss.backward(); opt.step(); opt.zero_grad()
    if sync:
        return loss.item()
for _ in range(3): step(True)
for sync in (True, False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): step(sync)
    torch.cuda.synchronize(); print("sync per step" if sync else "no sync", "ms/step", (time.perf_counter() - t0) * 50)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step(True)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
