"""B200-native batch source.

The reference feeds the train loop through ``DataLoader(dataset, batch_size, shuffle=True)``
over a host-resident table of per-ray samples (src/nerfdata/datasets/llff.py:59-105,
src/nerfdata/utils/splitter.py:123-132): a global shuffle of the flattened ray table, one
Python ``__getitem__`` per ray and a default collate — it caps out around 1e4-1e5 rays/s.
``DeviceRayLoader`` keeps the *images and poses* resident in HBM instead (a ray table is 7x
larger than the images) and turns a seeded permutation of pixel ids into ``(rays_o, rays_d,
rgb_gt)`` batches with ONE launch of the ray-generation kernel (``fsnerf_gen_rays``: pixel id
-> view, row, column bit-exact; NDC warp fused under ``ndc``).  Same iteration contract as the
reference loader: every epoch is a fresh permutation of all ``V*H*W`` rays, cut into batches of
``batch_size`` with a ragged last batch, yielding ``[B,3] x 3`` fp32 tensors (on the device).
"""
import math

import torch

from .. import ops


class DeviceRayLoader:
    def __init__(self, imgs, poses, hwf, batch_size, ndc=False, ndc_near=1.0, seed=None, device="cuda",
                 drop_last=False):
        self.device = torch.device(device)
        ops.require_device(self.device.index if self.device.index is not None else torch.cuda.current_device())
        self.imgs = torch.as_tensor(imgs, dtype=torch.float32).to(self.device).contiguous()  # [V,H,W,3]
        self.poses = torch.as_tensor(poses, dtype=torch.float32).to(self.device).contiguous()  # [V,3|4,4]
        self.H, self.W, self.focal = int(hwf[0]), int(hwf[1]), float(hwf[2])
        assert self.imgs.shape[1:] == (self.H, self.W, 3), "images must be [V,H,W,3]"
        self.batch_size = int(batch_size)
        self.ndc, self.ndc_near = bool(ndc), float(ndc_near)
        self.drop_last = drop_last
        self.n_rays = self.imgs.shape[0] * self.H * self.W
        self.gen = torch.Generator(device=self.device)
        if seed is not None:
            self.gen.manual_seed(int(seed))

    def __len__(self):
        n = self.n_rays / self.batch_size
        return math.floor(n) if self.drop_last else math.ceil(n)

    def batch(self, pixel_ids):
        """(rays_o, rays_d, rgb_gt) of the given flat pixel ids (id = (view*H + row)*W + col)"""
        return ops.gen_rays(self.poses, self.H, self.W, self.focal, pixel_ids=pixel_ids, images=self.imgs,
                            ndc=self.ndc, ndc_near=self.ndc_near)

    def __iter__(self):
        perm = torch.randperm(self.n_rays, generator=self.gen, device=self.device)
        for i in range(len(self)):
            yield self.batch(perm[i * self.batch_size:(i + 1) * self.batch_size])
