"""Drop-in for the reference's ``nerfdata.utils.splitter.Splitter``
(src/nerfdata/utils/splitter.py:13-388): LLFF scene folder -> poses / bounds / intrinsics ->
pose-based train/val/test split -> datasets and loaders, plus the spiral path for the video.

Same constructor, ``split()``, ``get_datasets()``, ``get_dataloaders()`` and attributes
(``poses [N,3,4]``, ``hwf``, ``min_bound``, ``max_bound``, ``path_poses [120,3,4]``,
``img_paths``, ``train_ids / val_ids / test_ids``).  Differences, all additive:
``root`` (the reference hard-codes ``../datasets/llff/``, splitter.py:183), images are decoded
with PIL (imageio is not in this image), and the train loader is the device-resident
``DeviceRayLoader`` (row a12) instead of a host ``DataLoader`` over per-ray items.

On-disk format (splitter.py:174-231): ``<root>/<scene>/poses_bounds.npy`` = [N,17]: a 3x5
camera-to-world in LLFF axes (down, right, back | position | H,W,focal) + near/far depth
bounds; ``images_8/`` = the frames at 1/8 resolution, sorted by name.
"""
import os
from typing import Optional

import numpy as np
from sklearn.cluster import KMeans

from ..datasets import llff


def _imread(path):
    from PIL import Image
    return np.asarray(Image.open(path))


def _unit(v):
    return v / np.linalg.norm(v)


def _look_at(z, up, pos):
    """[3,4] camera-to-world with viewing axis z, approximate up vector and position"""
    z = _unit(z)
    x = _unit(np.cross(up, z))
    y = _unit(np.cross(z, x))
    return np.stack([x, y, z, pos], axis=1)


def _average_pose(poses):
    """[3,5]: mean position, summed viewing / up axes, intrinsics column of pose 0"""
    centre = poses[:, :3, 3].mean(0)
    z = _unit(poses[:, :3, 2].sum(0))
    up = poses[:, :3, 1].sum(0)
    return np.concatenate([_look_at(z, up, centre), poses[0, :3, -1:]], 1)


def _recentre(poses):
    """express every pose in the frame of the average pose (splitter.py:283-302)"""
    out = poses.copy()
    last = np.array([[0, 0, 0, 1.0]])
    mean44 = np.concatenate([_average_pose(poses)[:3, :4], last], axis=0)
    all44 = np.concatenate([poses[:, :3, :4], np.broadcast_to(last, (len(poses), 1, 4))], axis=1)
    out[:, :3, :4] = (np.linalg.inv(mean44) @ all44)[:, :3, :4]
    return out


def _spiral(c2w, poses, bounds, n_views=120, n_rots=2, zrate=0.5):
    """spiral of look-at cameras around the average pose (splitter.py:337-388)"""
    up = _unit(poses[:, :3, 1].sum(0))
    close, far = bounds.min() * 0.9, bounds.max() * 5.0
    dt = 0.75
    focal = 1.0 / ((1.0 - dt) / close + dt / far)
    rads = np.append(np.percentile(np.abs(poses[:, :3, 3]), 90, 0), 1.0)
    hwf = c2w[:, 4:5]
    out = []
    for theta in np.linspace(0.0, 2.0 * np.pi * n_rots, n_views + 1)[:-1]:
        c = c2w[:3, :4] @ (np.array([np.cos(theta), -np.sin(theta), -np.sin(theta * zrate), 1.0]) * rads)
        z = _unit(c - c2w[:3, :4] @ np.array([0, 0, -focal, 1.0]))
        out.append(np.concatenate([_look_at(z, up, c), hwf], 1))
    return np.stack(out, 0)


class Splitter:
    def __init__(self, dataset_type: str, scene: str, strategy: str = "pose_based", n_training_views=-1,
                 val_ratio: float = 0.15, test_ratio: float = 0.15, random_seed: Optional[int] = None,
                 root: str = "../datasets/llff/"):
        self.dataset_type = dataset_type
        self.scene = scene
        self.strategy = strategy
        self.n_training_views = n_training_views
        self.val_ratio = val_ratio
        self.test_ratio = test_ratio
        self.random_seed = random_seed
        self.root = root
        self.image_paths = []
        self.poses = np.empty((0, 3, 4))
        self.train_ids = self.val_ids = self.test_ids = None
        self._load_dataset()

    # ------------------------------------------------------------------ split
    def split(self):
        """test, then val, then train views: K-means over camera positions, the view closest to
        every centre is taken (splitter.py:49-72, 134-160)."""
        available = np.arange(len(self.poses))
        self.test_ids, available = self._select_pose_based(available, int(self.test_ratio * len(self.poses)))
        self.val_ids, available = self._select_pose_based(available, int(self.val_ratio * len(self.poses)))
        if self.n_training_views < 0:
            self.train_ids = available
        else:
            assert self.n_training_views > 0, \
                "ValueError, the specified number of training images must be greater than zero."
            self.train_ids, _ = self._select_pose_based(available, self.n_training_views)

    def _select_pose_based(self, available_idxs: np.ndarray, n_samples: int):
        x = self.poses[available_idxs, :3, 3]
        km = KMeans(n_clusters=n_samples, n_init=10, random_state=self.random_seed).fit(x)
        dist = np.linalg.norm(x - km.cluster_centers_[km.labels_], axis=1)
        picked = np.array([np.argmin(np.where(km.labels_ == k, dist, np.inf)) for k in range(n_samples)], dtype=int)
        chosen = available_idxs[picked]
        return chosen, np.array([i for i in available_idxs if i not in chosen])

    # ------------------------------------------------------------------ datasets
    def get_datasets(self, train_img_mode: bool = False, **kwargs):
        assert self.train_ids is not None, "Split the source data before building the datasets."
        white_bkgd, ndc = kwargs.get("white_bkgd", False), kwargs.get("ndc", False)
        device = kwargs.get("device", "cuda")

        def build(ids, img_mode):
            return llff.LLFFDataset(self._load_img_files(self.img_paths[ids]), self.poses[ids], self.min_bound,
                                    self.max_bound, self.hwf, white_bkgd, img_mode, ndc, device=device)
        test, val = build(self.test_ids, True), build(self.val_ids, True)
        return build(self.train_ids, train_img_mode), val, test

    def get_dataloaders(self, train_batch_size, train_img_mode=False, **kwargs):
        """(train, val, test) loaders.  val/test iterate images like the reference's
        ``DataLoader(batch_size=1, shuffle=True)``; the train loader is the device-resident ray
        source (``.dataset`` keeps the attributes run-nerf.py reads: hwf, near, far, aabb)."""
        from torch.utils.data import DataLoader
        train_set, val_set, test_set = self.get_datasets(train_img_mode, **kwargs)
        if train_img_mode:
            train_loader = DataLoader(train_set, batch_size=train_batch_size, shuffle=True)
        else:
            train_loader = train_set.device_loader(train_batch_size, seed=self.random_seed)
            train_loader.dataset = train_set
        return (train_loader, DataLoader(val_set, batch_size=1, shuffle=True),
                DataLoader(test_set, batch_size=1, shuffle=True))

    # ------------------------------------------------------------------ loading
    def _load_dataset(self):
        if self.dataset_type == "llff":
            self._load_llff_dataset()
        else:
            raise ValueError(f"Dataset of type '{self.dataset_type}' is not supported.")

    def _load_llff_dataset(self):
        base = os.path.normpath(self.root)
        assert os.path.isdir(base), f"LLFF dataset folder {os.path.abspath(base)} not found."
        assert self.scene in os.listdir(base), f"Scene '{self.scene}' not found in local LLFF dataset folder."
        data = np.load(os.path.join(base, self.scene, "poses_bounds.npy"))
        poses = data[:, :15].reshape(-1, 3, 5).astype(np.float64)  # [N,3,5]
        bounds = data[:, 15:].astype(np.float32)                    # [N,2]
        img_dir = os.path.normpath(os.path.join(base, self.scene, "images_8/"))
        assert os.path.isdir(img_dir), f"Images folder path {os.path.abspath(img_dir)} not found."
        self.img_paths = np.array([os.path.abspath(os.path.join(img_dir, f)) for f in sorted(os.listdir(img_dir))
                                   if f.endswith(("JPG", "jpg", "png"))])
        assert len(self.img_paths) == poses.shape[0], "Mismath between the number of images and poses"
        H, W = _imread(self.img_paths[0]).shape[:2]
        poses[:, 0, 4], poses[:, 1, 4] = H, W
        poses[:, 2, 4] *= 1.0 / 8.0  # the frames are the 1/8 resolution set
        # LLFF axes (down, right, back) -> (right, up, back)
        poses = np.concatenate([poses[:, :, 1:2], -poses[:, :, 0:1], poses[:, :, 2:]], axis=2).astype(np.float32)
        self.postprocess_poses(poses, bounds)

    def postprocess_poses(self, poses: np.ndarray, bounds: np.ndarray, factor: int = 4, bd_factor: float = 0.75,
                          recenter: bool = True, ndc: bool = True):
        """scale so that the nearest bound sits at 1/bd_factor, recentre on the average pose,
        build the spiral (splitter.py:304-327).  NB ``min_bound`` / ``max_bound`` are the
        extrema of the whole pose array (incl. its intrinsics column), as in the reference."""
        scale = 1.0 if bd_factor is None else 1.0 / (bounds.min() * bd_factor)
        poses[..., :3, 3] *= scale
        bounds *= scale
        if recenter:
            poses = _recentre(poses)
        self.path_poses = _spiral(_average_pose(poses), poses, bounds)[:, :3, :4]
        hwf = poses[0, :3, -1]
        self.hwf = (int(hwf[0]), int(hwf[1]), float(hwf[2]))
        self.poses = poses[:, :3, :4]
        self.min_bound = poses.min()
        self.max_bound = poses.max()

    def _load_img_files(self, img_paths):
        return np.stack([_imread(p)[..., :3] / 255.0 for p in img_paths], axis=0)
