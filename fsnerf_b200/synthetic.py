"""Procedural synthetic scenes in the reference's on-disk formats (fixtures for
tests and bench.py — there is no network for the real datasets).

Blender format as read by /root/reference/src/nerfdata/datasets/blender.py:217-258
(``transforms_{split}.json`` with ``camera_angle_x`` + per-frame
``transform_matrix``, RGBA PNGs); cameras on the radius-4.0311289 sphere of
``pose_from_spherical`` (blender.py:51-70,260-277), near/far 2/6
(blender.py:104-105).  The scene is a few coloured spheres ray-traced
analytically on the CPU with the reference's ``get_rays`` convention.
Host-side numpy only; nothing here is on the timed hot path.
"""
import json
import os

import numpy as np

FOV_X = 0.6911112  # Blender synthetic set
RADIUS = 4.0311289
NEAR, FAR = 2.0, 6.0
f32 = np.float32

SPHERES = [  # centre, radius, rgb
    ((0.0, 0.0, 0.0), 0.9, (0.85, 0.25, 0.2)),
    ((0.9, 0.4, 0.3), 0.45, (0.2, 0.6, 0.9)),
    ((-0.7, -0.6, 0.5), 0.4, (0.3, 0.8, 0.35)),
    ((0.1, -0.8, -0.6), 0.35, (0.95, 0.8, 0.2)),
]
LIGHT = np.array([0.4, 0.5, 0.77], f32) / np.linalg.norm([0.4, 0.5, 0.77])


def pose_from_spherical(radius, theta_deg, phi_deg):
    """reference: src/nerfdata/datasets/blender.py:51-70 (c2w = rot_phi @ rot_theta @ trans_t)."""
    t = np.eye(4, dtype=f32)
    t[2, 3] = radius
    th, ph = np.deg2rad(theta_deg), np.deg2rad(phi_deg)
    rt = np.array([[1, 0, 0, 0], [0, np.cos(th), -np.sin(th), 0], [0, np.sin(th), np.cos(th), 0],
                   [0, 0, 0, 1]], f32)
    rp = np.array([[np.cos(ph), -np.sin(ph), 0, 0], [np.sin(ph), np.cos(ph), 0, 0], [0, 0, 1, 0],
                   [0, 0, 0, 1]], f32)
    return (rp @ rt @ t).astype(f32)


def focal_from_fov(W, fov_x=FOV_X):
    """reference: blender.py:250-251  focal = 0.5*W/tan(0.5*fov_x)"""
    return float(0.5 * W / np.tan(0.5 * fov_x))


def camera_rays(pose, H, W, focal):
    """numpy restatement of the reference get_rays convention (src/utils/utilities.py:54-82)."""
    i, j = np.meshgrid(np.arange(W, dtype=f32), np.arange(H, dtype=f32), indexing="xy")
    dirs = np.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -np.ones_like(i)], -1)
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    rd = dirs @ pose[:3, :3].T
    ro = np.broadcast_to(pose[:3, 3], rd.shape)
    return ro.astype(f32), rd.astype(f32)


def trace(ro, rd, white_bkgd=True):
    """analytic render -> rgb[...,3], alpha[...]"""
    shape = rd.shape[:-1]
    ro, rd = ro.reshape(-1, 3).astype(np.float64), rd.reshape(-1, 3).astype(np.float64)
    best_t = np.full(len(rd), np.inf)
    rgb = np.ones((len(rd), 3)) if white_bkgd else np.zeros((len(rd), 3))
    for c, r, col in SPHERES:
        oc = ro - np.array(c)
        b = (oc * rd).sum(-1)
        disc = b * b - ((oc * oc).sum(-1) - r * r)
        hit = disc > 0
        t = -b - np.sqrt(np.where(hit, disc, 0))
        hit &= (t > 0) & (t < best_t)
        n = (ro + t[:, None] * rd - np.array(c)) / r
        shade = 0.35 + 0.65 * np.clip((n * LIGHT).sum(-1), 0, 1)
        rgb[hit] = (np.array(col)[None] * shade[:, None])[hit]
        best_t = np.where(hit, t, best_t)
    alpha = np.isfinite(best_t).astype(f32)
    return rgb.reshape(*shape, 3).astype(f32), alpha.reshape(shape)


def make_views(n_views, H, W, seed=42, theta_range=(20.0, 75.0)):
    """-> poses [V,4,4], images [V,H,W,3] (white background blended), focal"""
    rng = np.random.default_rng(seed)
    focal = focal_from_fov(W)
    poses, imgs = [], []
    for v in range(n_views):
        theta = rng.uniform(*theta_range)
        phi = 360.0 * v / n_views + rng.uniform(-10, 10)
        pose = pose_from_spherical(RADIUS, theta, phi)
        ro, rd = camera_rays(pose, H, W, focal)
        rgb, _ = trace(ro, rd, True)
        poses.append(pose)
        imgs.append(rgb)
    return np.stack(poses).astype(f32), np.stack(imgs).astype(f32), focal


def orbit_poses(n_frames, theta=50.0):
    """reference: blender.py:260-277 (__build_path)"""
    return np.stack([pose_from_spherical(RADIUS, theta, phi)
                     for phi in np.linspace(0, 360, n_frames, endpoint=False)]).astype(f32)


def write_blender_scene(root, n_views=8, H=100, W=100, seed=42, splits=("train", "val", "test")):
    """Write a scene directory the reference's BlenderDataset can read."""
    from PIL import Image
    os.makedirs(root, exist_ok=True)
    for si, split in enumerate(splits):
        rng_seed = seed + si
        poses, _, focal = make_views(n_views, H, W, rng_seed)
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for v, pose in enumerate(poses):
            ro, rd = camera_rays(pose, H, W, focal)
            rgb, alpha = trace(ro, rd, False)
            rgba = np.concatenate([rgb, alpha[..., None]], -1)
            Image.fromarray((255 * np.clip(rgba, 0, 1)).astype(np.uint8), "RGBA").save(
                os.path.join(root, split, f"r_{v}.png"))
            frames.append({"file_path": f"./{split}/r_{v}", "transform_matrix": pose.tolist()})
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as f:
            json.dump({"camera_angle_x": FOV_X, "frames": frames}, f)
    return root


def make_llff_views(n_views, H, W, seed=42):
    """Forward-facing cameras for the NDC configuration (C3): poses looking down
    -z from a small grid around (0,0,RADIUS)."""
    rng = np.random.default_rng(seed)
    focal = focal_from_fov(W)
    poses, imgs = [], []
    for v in range(n_views):
        pose = np.eye(4, dtype=f32)
        pose[:3, 3] = [rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), RADIUS]
        ro, rd = camera_rays(pose, H, W, focal)
        rgb, _ = trace(ro, rd, True)
        poses.append(pose)
        imgs.append(rgb)
    return np.stack(poses).astype(f32), np.stack(imgs).astype(f32), focal


def write_llff_scene(root, scene="synth", n_views=9, H=24, W=32, seed=42):
    """Write ``<root>/<scene>/poses_bounds.npy`` + ``images_8/*.png`` in the LLFF layout read by the
    reference's Splitter (/root/reference/src/nerfdata/utils/splitter.py:174-231): per view a 3x5
    camera-to-world in LLFF axes (down, right, back | position | H, W, focal at FULL resolution,
    i.e. 8x the images_8 frames) followed by the near/far depth bounds.  Forward-facing cameras
    with small random rotations around (0, 0, RADIUS)."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    focal = focal_from_fov(W)
    d = os.path.join(root, scene, "images_8")
    os.makedirs(d, exist_ok=True)
    rows = []
    for v in range(n_views):
        yaw, pitch = rng.uniform(-0.08, 0.08, 2)
        ry = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        rx = np.array([[1, 0, 0], [0, np.cos(pitch), -np.sin(pitch)], [0, np.sin(pitch), np.cos(pitch)]])
        pose = np.eye(4)
        pose[:3, :3] = ry @ rx
        pose[:3, 3] = [rng.uniform(-0.6, 0.6), rng.uniform(-0.4, 0.4), RADIUS + rng.uniform(-0.1, 0.1)]
        ro, rd = camera_rays(pose.astype(f32), H, W, focal)
        rgb, _ = trace(ro, rd, True)
        Image.fromarray((255 * np.clip(rgb, 0, 1)).astype(np.uint8), "RGB").save(os.path.join(d, f"img_{v:03d}.png"))
        m = np.concatenate([-pose[:3, 1:2], pose[:3, 0:1], pose[:3, 2:3], pose[:3, 3:4],
                            np.array([[8.0 * H], [8.0 * W], [8.0 * focal]])], 1)  # (down, right, back | t | hwf)
        rows.append(np.concatenate([m.reshape(-1), [RADIUS - 1.6 + 0.05 * v, RADIUS + 1.6]]))
    np.save(os.path.join(root, scene, "poses_bounds.npy"), np.stack(rows).astype(np.float64))
    return os.path.join(root, scene)
