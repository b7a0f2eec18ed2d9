"""Drop-in for the reference's ``utils.utilities`` ray helpers
(/root/reference/src/utils/utilities.py:36-134) — same names, argument meaning
and return structure; the arithmetic runs in the CUDA ray kernels."""
from typing import List, Tuple

import torch
from torch import Tensor

from .. import ops


def get_rays(pose: Tensor, hwf: Tuple[int, int, float],
             device: torch.device = torch.device("cuda")) -> Tuple[Tensor, Tensor]:
    """reference: src/utils/utilities.py:36-82.  pose [4,4] or [3,4] ->
    (origins [H,W,3], dirs [H,W,3]); pixel (h, w) is row h*W + w when flattened.
    (The reference returns origins as a stride-0 expand of pose[:3,-1]; ours is
    a materialised tensor with the same values.)"""
    H, W, focal = hwf
    H, W = int(H), int(W)
    device = torch.device(device)
    if device.type != "cuda":
        raise ops._lib.FsnerfError("get_rays: fsnerf_b200 has no CPU path; pass a CUDA device")
    pose = pose.to(device=device, dtype=torch.float32)
    o, d, _ = ops.gen_rays(pose[None].contiguous(), H, W, float(focal), first_id=0, n_rays=H * W)
    return o.reshape(H, W, 3), d.reshape(H, W, 3)


def to_ndc(rays_o: Tensor, rays_d: Tensor, hwf: Tuple[int, int, float],
           near: float) -> Tuple[Tensor, Tensor]:
    """reference: src/utils/utilities.py:84-120 (directions are not renormalised)."""
    H, W, focal = hwf
    return ops.to_ndc(rays_o, rays_d, int(H), int(W), float(focal), float(near))


def get_chunks(inputs: Tensor, chunksize: int) -> List[Tensor]:
    """reference: src/utils/utilities.py:122-134 (views; last chunk ragged)."""
    return [inputs[i:i + chunksize] for i in range(0, inputs.shape[0], chunksize)]
