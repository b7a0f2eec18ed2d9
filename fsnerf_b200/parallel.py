"""Data-parallel plumbing of the hot path: rays shard across
ranks with ONE collective per step (SUM all-reduce of the flat fp32 gradient of
both networks); full-image rendering partitions pixels with no collective.
Pure host logic + torch.distributed — exercised on CPU with the gloo backend
(tests/test_parallel_cpu.py) and on GPUs with NCCL (bench.py, tools/dp_check.py).
"""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous slice [start, stop) of n items owned by `rank` (sizes differ by <= 1)"""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_batch(global_batch: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """rank's slice of a [G, ...] global batch (the seeded pixel permutation, rays, uniforms)"""
    a, b = shard_range(global_batch.shape[0], rank, world)
    return global_batch[a:b]


def loss_grad_scale(global_rays: int) -> float:
    """Every rank seeds d(loss)/d(rgb) with 2*(rgb-gt)/(3*G): the SUM over ranks of the
    local gradients is then the gradient of F.mse_loss over the GLOBAL batch
    (reference: src/run-nerf.py:256)."""
    return 1.0 / (3.0 * global_rays)


def allreduce_gradients(flat_grads: torch.Tensor, group=None) -> torch.Tensor:
    """the path's only exchange: in-place SUM of the flat gradient buffer"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return flat_grads


def pixel_partition(n_frames: int, H: int, W: int, rank: int, world: int):
    """per-frame flattened pixel ranges rendered by `rank`:
    [(frame, first_pixel, last_pixel_exclusive), ...] covering its slice of F*H*W"""
    start, stop = shard_range(n_frames * H * W, rank, world)
    out = []
    for f in range(n_frames):
        a, b = max(start, f * H * W), min(stop, (f + 1) * H * W)
        if a < b:
            out.append((f, a - f * H * W, b - f * H * W))
    return out
