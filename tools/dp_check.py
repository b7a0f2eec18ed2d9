"""N-GPU data-parallel check (run under torchrun): the all-reduced gradient of the
ray-sharded step equals the single-GPU gradient of the same global batch.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dp_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fsnerf_b200 import parallel, synthetic as syn  # noqa: E402
from fsnerf_b200.engine import HotPath  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
G, Sc, Sf = 2048, 64, 128
H = W = 64
poses, imgs, focal = syn.make_views(4, H, W, seed=42)
rng = np.random.default_rng(0)
ids = rng.permutation(4 * H * W)[:G]
o = np.concatenate([syn.camera_rays(p, H, W, focal)[0].reshape(-1, 3) for p in poses])[ids]
d = np.concatenate([syn.camera_rays(p, H, W, focal)[1].reshape(-1, 3) for p in poses])[ids]
gt = imgs.reshape(-1, 3)[ids]
us, up = rng.random((G, Sc), dtype=np.float32), rng.random((G, Sf), dtype=np.float32)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
a, b = parallel.shard_range(G, rank, world)
hp = HotPath(n_coarse=Sc, n_fine=Sf, device=dev)
ls = hp.train_step(cu(o[a:b]), cu(d[a:b]), cu(gt[a:b]), cu(us[a:b]), cu(up[a:b]), global_rays=G, apply_update=False)
ls = ls.clone()
dist.all_reduce(ls)
ok = True
if rank == 0:
    hp1 = HotPath(n_coarse=Sc, n_fine=Sf, device=dev, world_size=1)
    ls1 = hp1.train_step(cu(o), cu(d), cu(gt), cu(us), cu(up), global_rays=G, apply_update=False)
    rel = ((hp.grads - hp1.grads).norm() / hp1.grads.norm()).item()
    dl = (ls - ls1).abs().max().item() / ls1.abs().max().item()
    print(f"dp{world} vs single GPU: grad rel err {rel:.3e}, loss rel err {dl:.3e}")
    ok = rel < 1e-3 and dl < 1e-5
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
