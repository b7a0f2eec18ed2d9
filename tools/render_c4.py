"""BASELINE.json configs[3]: full-resolution 800x800 render of 200 test poses through the drop-in
``render_path`` (src/render/rendering.py:180-248), the flattened F*H*W pixel range partitioned across
the ranks with no collective.  Run alone (1 GPU) or under torchrun (N GPUs):
    python tools/render_c4.py [--frames 200] [--chunk 65536]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29540 tools/render_c4.py
Prints one JSON line (rank 0): Mrays/s over the whole job incl. ray generation and the device->host
copy of every frame (render_path returns numpy arrays like the reference)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fsnerf_b200 import synthetic as syn  # noqa: E402
from fsnerf_b200.core.models import NeRF  # noqa: E402
from fsnerf_b200.render.rendering import HierarchicalEstimator, render_path  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--chunk", type=int, default=65536)
ap.add_argument("--size", type=int, default=800)
args = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
H = W = args.size
hwf = (H, W, syn.focal_from_fov(W))
poses = torch.from_numpy(syn.orbit_poses(args.frames))
kw = {"pos_fn": {"n_freqs": 10, "log_space": True}, "dir_fn": {"n_freqs": 4, "log_space": True}}
torch.manual_seed(42)
coarse, fine = NeRF(3, 3, 8, 256, [4], **kw).to(dev).eval(), NeRF(3, 3, 8, 256, [4], **kw).to(dev).eval()
est = HierarchicalEstimator(near=2.0, far=6.0, n_coarse=64, n_fine=128, proposal_model=coarse).eval()
render_path(poses[:1], hwf, 2.0, 6.0, args.chunk, fine, est, white_bkgd=True, device=dev)  # warm-up (allocator, kernels)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
out = render_path(poses, hwf, 2.0, 6.0, args.chunk, fine, est, white_bkgd=True, device=dev, rank=rank, world_size=world)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
n_local = out[0].reshape(-1, 3).shape[0]
if rank == 0:
    total = args.frames * H * W
    print(json.dumps({"config": f"C4: {args.frames} poses {H}x{W}, coarse 64 + fine 64+128 samples, chunk {args.chunk}",
                      "n_gpus": world, "rays": total, "rays_this_rank": n_local, "seconds": dt.item(),
                      "render_mrays_per_s": total / dt.item() / 1e6, "host_output_gb": total * 16 / 1e9}))
if world > 1:
    dist.destroy_process_group()
