import os, sys, time, torch
sys.path.insert(0, "/root/repo")
from fsnerf_b200 import ops, synthetic as syn
from fsnerf_b200.core.models import NeRF
from fsnerf_b200.engine import HotPath
from fsnerf_b200.render.rendering import HierarchicalEstimator, render_rays
dev = torch.device("cuda:0")
kw = {"pos_fn": {"n_freqs": 10, "log_space": True}, "dir_fn": {"n_freqs": 4, "log_space": True}}
coarse, fine = NeRF(3, 3, 8, 256, [4], **kw).to(dev).eval(), NeRF(3, 3, 8, 256, [4], **kw).to(dev).eval()
est = HierarchicalEstimator(near=2.0, far=6.0, n_coarse=64, n_fine=128, proposal_model=coarse).eval()
pose = torch.from_numpy(syn.orbit_poses(4)[1]).to(dev)[None].contiguous()
ro, rd, _ = ops.gen_rays(pose, 800, 800, syn.focal_from_fov(800), first_id=100000, n_rays=65536)
hp = HotPath(device=dev)
def t(fn, n=4):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
with torch.no_grad():
    print("HotPath.render ms:", t(lambda: hp.render(ro, rd)))
    print("render_rays ms:", t(lambda: render_rays(ro, rd, est, fine, white_bkgd=True, device=dev)))
    ops.profile_enable(True)
    render_rays(ro, rd, est, fine, white_bkgd=True, device=dev)
    print({k: round(v[0], 3) for k, v in ops.profile_read().items()})
    ops.profile_enable(False)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        render_rays(ro, rd, est, fine, white_bkgd=True, device=dev); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
