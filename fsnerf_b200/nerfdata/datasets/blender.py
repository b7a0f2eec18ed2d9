"""Drop-in for the reference's ``nerfdata.datasets.blender.BlenderDataset``
(src/nerfdata/datasets/blender.py:72-277; ``datasets/dataset.py:SyntheticRealistic`` is the
same class under another name).

On-disk format (blender.py:217-258): ``<root>/<scene>/transforms_<split>.json`` with
``camera_angle_x`` and per frame ``file_path`` + 4x4 ``transform_matrix``; RGBA PNGs;
``focal = 0.5 W / tan(0.5 camera_angle_x)``; near/far = 2/6; white background = alpha blend
onto 1; ``path_poses`` = 90 cameras on the radius-4.0311289 sphere at colatitude 50 deg
(blender.py:260-277).

Position taken on a reference defect (blender.py:124-125): the view-selection line
``x = x[x[:, -1, -1] > 0]`` indexes a [N,3] array with three subscripts and raises
IndexError, so the reference class cannot be constructed as shipped.  Its comment states the
intent ("remove poses with negative z-coordinates"); this mirror implements that intent and
maps the K-means picks back to frame indices.  ``n_imgs=None`` keeps every view.
"""
import json
import os
from typing import Tuple

import numpy as np
import torch
from sklearn.cluster import KMeans
from torch import Tensor
from torch.utils.data import Dataset

from ..loader import DeviceRayLoader


def pose_from_spherical(radius: float, theta: float, phi: float) -> torch.Tensor:
    """camera-to-world looking at the origin from spherical coordinates (degrees):
    rot_z(phi) @ rot_x(theta) @ translate_z(radius)  (blender.py:20-70)"""
    th, ph = theta / 180.0 * np.pi, phi / 180.0 * np.pi
    t = torch.eye(4)
    t[2, 3] = radius
    rx = torch.Tensor([[1, 0, 0, 0], [0, np.cos(th), -np.sin(th), 0], [0, np.sin(th), np.cos(th), 0], [0, 0, 0, 1]])
    rz = torch.Tensor([[np.cos(ph), -np.sin(ph), 0, 0], [np.sin(ph), np.cos(ph), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    return rz @ (rx @ t)


class BlenderDataset(Dataset):
    def __init__(self, scene: str, split: str, n_imgs: int = None, img_mode: bool = False,
                 white_bkgd: bool = False, root: str = os.path.join("..", "datasets", "synthetic"),
                 device="cuda") -> None:
        super().__init__()
        self.scene, self.split, self.root, self.device = scene, split, root, device
        self.near, self.far = 2.0, 6.0
        self.ndc = False
        self.img_mode = img_mode
        imgs, poses, hwf = self._load()
        self.path_poses = torch.stack([pose_from_spherical(4.0311289, 50.0, phi)
                                       for phi in np.linspace(0, 360, 90, endpoint=False)], 0)
        self.hwf = hwf
        if white_bkgd:
            imgs = imgs[..., :3] * imgs[..., -1:] + (1.0 - imgs[..., -1:])
        else:
            imgs = imgs[..., :3]
        idx = np.random.randint(0, imgs.shape[0])  # view used for visual comparisons
        self.testimg, self.testpose = imgs[idx], poses[idx]
        if n_imgs is not None:
            keep = self.select_views(poses, n_imgs)
            imgs, poses = imgs[keep], poses[keep]
        self.imgs, self.poses = imgs, poses
        self.aabb = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5])  # for the occupancy-grid estimator
        self._samples = None

    @staticmethod
    def select_views(poses: Tensor, n_imgs: int) -> np.ndarray:
        """n_imgs views covering the scene: K-means over the camera positions above the ground
        plane, the view closest to every centre (blender.py:122-135, see the module docstring)"""
        pos = poses[:, :3, 3].numpy()
        upper = np.nonzero(pos[:, 2] > 0)[0]
        x = pos[upper]
        km = KMeans(n_clusters=n_imgs, n_init=10).fit(x)
        dist = np.linalg.norm(x - km.cluster_centers_[km.labels_], axis=1)
        return upper[[int(np.argmin(np.where(km.labels_ == k, dist, np.inf))) for k in range(n_imgs)]]

    def _load(self) -> Tuple[Tensor, Tensor, Tuple[int, int, float]]:
        from PIL import Image
        path = os.path.join(self.root, self.scene)
        with open(os.path.join(path, f"transforms_{self.split}.json"), "r") as f:
            meta = json.load(f)
        poses = np.stack([np.array(fr["transform_matrix"]) for fr in meta["frames"]], 0).astype(np.float32)
        imgs = np.stack([np.asarray(Image.open(os.path.join(path, fr["file_path"] + ".png")))
                         for fr in meta["frames"]], 0)
        imgs = (imgs / 255.0).astype(np.float32)
        H, W = imgs.shape[1:3]
        focal = 0.5 * W / np.tan(0.5 * meta["camera_angle_x"])
        return torch.from_numpy(imgs), torch.from_numpy(poses), (H, W, float(focal))

    # ---- per-ray table (blender.py:176-193), built lazily by the ray-generation kernel
    def _build_data(self):
        if self._samples is None:
            from ... import ops
            H, W, focal = self.hwf
            poses = self.poses.to(torch.device(self.device))
            rays_o, rays_d, _ = ops.gen_rays(poses, H, W, focal, n_rays=poses.shape[0] * H * W)
            self._samples = (rays_o, rays_d, self.imgs.reshape(-1, 3))
        return self._samples

    rays_o = property(lambda self: self._build_data()[0])
    rays_d = property(lambda self: self._build_data()[1])
    rgb = property(lambda self: self._build_data()[2])

    def device_loader(self, batch_size: int, seed=None, device=None) -> DeviceRayLoader:
        """the B200-native replacement of DataLoader(self, batch_size, shuffle=True)"""
        return DeviceRayLoader(self.imgs, self.poses, self.hwf, batch_size, seed=seed, device=device or self.device)

    def __len__(self) -> int:
        if self.img_mode:
            return len(self.imgs)
        return self.imgs.shape[0] * self.hwf[0] * self.hwf[1]

    def __getitem__(self, idx: int):
        if self.img_mode:
            return self.imgs[idx], self.poses[idx]
        return self.rays_o[idx], self.rays_d[idx], self.rgb[idx]
