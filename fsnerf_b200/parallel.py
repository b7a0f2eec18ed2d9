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


def allreduce_module_gradients(models, group=None) -> None:
    """Data-parallel exchange for the reference-style loop (``loss.backward()`` on the drop-in
    modules, then ``optimizer.step()``): SUM of every parameter gradient over the ranks.  A drop-in
    ``NeRF``'s gradients are views of ONE flat buffer written by its backward kernels, so each
    network costs one collective; any other module falls back to a coalesced copy."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    for m in models:
        flat = getattr(m, "_last_flat_grad", None)
        params = [p for p in m.parameters() if p.grad is not None]
        if flat is not None and hasattr(m, "_layout") and len(params) == len(m._layout) and all(
                p.grad.data_ptr() == flat.data_ptr() + 4 * o and p.grad.is_contiguous()
                for (o, _), p in zip(m._layout, m._param_list())):
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            continue
        if not params:
            continue
        buf = torch.cat([p.grad.reshape(-1) for p in params])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        o = 0
        for p in params:
            n = p.grad.numel()
            p.grad.copy_(buf[o:o + n].view_as(p.grad))
            o += n


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
