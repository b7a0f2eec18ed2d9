"""Drop-in for the reference's ``nerfdata.datasets.llff.LLFFDataset``
(src/nerfdata/datasets/llff.py:16-112): same constructor, attributes (``imgs, poses, hwf,
near, far, ndc, img_mode, rays_o, rays_d, rgb, aabb``), ``__getitem__`` / ``__len__``.

Ray bounds: ``near, far = 0.9*min_bound, max_bound`` or ``0, 1`` under NDC (llff.py:47-53).
The per-ray table the reference builds eagerly on the host (llff.py:59-90: ``get_rays`` per
pose, NDC warp with near = 1.0, ROI box / 2**3) is built here on first use by the B200
ray-generation kernel (no CPU path); training should not touch it at all and use
``device_loader()`` — images and poses resident in HBM, rays generated per batch.
"""
from typing import Tuple

import numpy as np
import torch
from torch import Tensor
from torch.utils.data import Dataset

from ..loader import DeviceRayLoader


class LLFFDataset(Dataset):
    def __init__(self, imgs: np.array, poses: np.array, min_bound: float, max_bound: float,
                 hwf: Tuple[int, int, float], white_bkgd: bool = False, img_mode: bool = False,
                 ndc: bool = True, device="cuda") -> None:
        super(LLFFDataset, self).__init__()
        self.imgs = torch.tensor(imgs, dtype=torch.float32)
        self.poses = torch.tensor(poses, dtype=torch.float32)
        self.hwf = hwf
        self.white_bkgd = white_bkgd
        self.img_mode = img_mode
        self.ndc = ndc
        self.device = device
        if not ndc:
            self.near = min_bound * 0.9
            self.far = max_bound * 1.0
        else:
            self.near = 0.0
            self.far = 1.0
        self._samples = None

    # ---- per-ray table (llff.py:59-90), built lazily on the device
    def _build_samples(self):
        if self._samples is None:
            from ... import ops
            H, W, focal = self.hwf
            dev = torch.device(self.device)
            poses = self.poses.to(dev)
            rays_o, rays_d, _ = ops.gen_rays(poses, H, W, focal, n_rays=poses.shape[0] * H * W, ndc=self.ndc,
                                             ndc_near=1.0)
            if self.ndc:
                lo = torch.minimum(rays_o.min(dim=0)[0], (rays_o + rays_d).min(dim=0)[0])
                hi = torch.maximum(rays_o.max(dim=0)[0], (rays_o + rays_d).max(dim=0)[0])
                aabb = torch.hstack([lo, hi]) / 2 ** (4 - 1)
            else:
                aabb = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5])
            self._samples = (rays_o, rays_d, self.imgs.reshape(-1, 3), aabb)
        return self._samples

    rays_o = property(lambda self: self._build_samples()[0])
    rays_d = property(lambda self: self._build_samples()[1])
    rgb = property(lambda self: self._build_samples()[2])
    aabb = property(lambda self: self._build_samples()[3])

    def device_loader(self, batch_size: int, seed=None, device=None) -> DeviceRayLoader:
        """the B200-native replacement of DataLoader(self, batch_size, shuffle=True)"""
        return DeviceRayLoader(self.imgs, self.poses, self.hwf, batch_size, ndc=self.ndc, ndc_near=1.0,
                               seed=seed, device=device or self.device)

    def __getitem__(self, idx: int) -> Tuple[Tensor, Tensor, Tensor]:
        if self.img_mode:
            return self.imgs[idx], self.poses[idx]
        return self.rays_o[idx], self.rays_d[idx], self.rgb[idx]

    def __len__(self) -> int:
        if self.img_mode:
            return self.imgs.shape[0]
        H, W, _ = self.hwf
        return self.imgs.shape[0] * H * W
