"""The reference's run loop (src/run-nerf.py: init_models -> train -> evaluation -> checkpoint ->
render_path) restated on the drop-in modules only — what a user of a-lemus96/fs-nerf runs after
switching the imports.  Procedural Blender-format scene on disk, BlenderDataset
loaders, coarse+fine NeRF with the hierarchical sampler (or --sampler occgrid for the reference's
occupancy-grid path), exponential LR decay, FreeNeRF frequency mask, occlusion regulariser.

    python examples/train_dropin.py --iters 300 --size 64
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fsnerf_b200.core.loss as L  # noqa: E402
import fsnerf_b200.core.models as M  # noqa: E402
import fsnerf_b200.core.scheduler as S  # noqa: E402
import fsnerf_b200.render.rendering as R  # noqa: E402
from fsnerf_b200 import synthetic as syn  # noqa: E402
from fsnerf_b200.evaluation import evaluation, load_checkpoint, save_checkpoint  # noqa: E402
from fsnerf_b200.nerfdata.datasets.blender import BlenderDataset  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--sampler", default="hierarchical", choices=["hierarchical", "occgrid"])
    ap.add_argument("--out", default=None)
    args = ap.parse_args(argv)
    device = torch.device("cuda:0")
    torch.manual_seed(42)
    np.random.seed(42)
    out_dir = args.out or tempfile.mkdtemp(prefix="fsnerf_b200_")
    root = os.path.join(out_dir, "datasets", "synthetic")
    syn.write_blender_scene(os.path.join(root, "spheres"), n_views=args.views, H=args.size, W=args.size, seed=42,
                            splits=("train", "val"))
    train_set = BlenderDataset("spheres", "train", white_bkgd=True, root=root, device=device)
    val_set = BlenderDataset("spheres", "val", n_imgs=2, img_mode=True, white_bkgd=True, root=root, device=device)
    train_loader = train_set.device_loader(args.batch, seed=42)
    val_loader = DataLoader(val_set, batch_size=1, shuffle=False)

    kw = {"pos_fn": {"n_freqs": 10, "log_space": True}, "dir_fn": {"n_freqs": 4, "log_space": True}}
    model = M.NeRF(3, 3, 8, 256, [4], **kw).to(device)
    params = list(model.parameters())
    step_size = 5e-3
    if args.sampler == "hierarchical":
        coarse = M.NeRF(3, 3, 8, 256, [4], **kw).to(device)
        params += list(coarse.parameters())
        estimator = R.HierarchicalEstimator(train_set.aabb, 128, 1, near=train_set.near, far=train_set.far,
                                            n_coarse=64, n_fine=64, proposal_model=coarse)
    else:
        estimator = R.OccGridEstimator(train_set.aabb, resolution=64, levels=1).to(device)
        step_size = 2e-2
    optimizer = torch.optim.Adam(params, lr=5e-4)
    scheduler = S.ExponentialDecay(optimizer, args.iters, 5e-4, r=0.1)
    occ_reg = L.OcclusionRegularizer(0.5, 2.0, "linear")
    reg_steps = int(0.9 * args.iters)
    it = iter(train_loader)
    psnrs = []
    for k in range(args.iters):
        model.train()
        estimator.train()
        try:
            rays_o, rays_d, rgb_gt = next(it)
        except StopIteration:
            it = iter(train_loader)
            rays_o, rays_d, rgb_gt = next(it)
        model.set_freq_mask(k, reg_steps)
        if args.sampler == "hierarchical":
            estimator.proposal_model.set_freq_mask(k, reg_steps)
        (rgb, *_, extras), ray_indices, t_vals = R.render_rays(rays_o, rays_d, estimator, model, train=True,
                                                              white_bkgd=True, render_step_size=step_size,
                                                              device=device)
        if extras is None:
            # the AssertionError fallback of render_rays (src/render/rendering.py:97-103: exactly one surviving
            # sample): a constant background with no graph behind it — the reference's loss.backward() would
            # raise here; skip the batch
            estimator.update_every_n_steps(step=k, occ_eval_fn=lambda x: model(x) * step_size, occ_thre=1e-2)
            continue
        loss = F.mse_loss(rgb, rgb_gt)
        psnrs.append(-10.0 * torch.log10(loss.detach()).item())
        if "rgb_coarse" in extras:
            loss = loss + F.mse_loss(extras["rgb_coarse"], rgb_gt)
        if len(extras["sigmas"]) > 0 and k < 20:  # a few steps of the occlusion term (gated like args.beta)
            loss = loss + 1e-3 * occ_reg(extras["sigmas"], t_vals, ray_indices)
        loss.backward()
        optimizer.step()
        scheduler.step()
        optimizer.zero_grad()
        estimator.update_every_n_steps(step=k, occ_eval_fn=lambda x: model(x) * step_size, occ_thre=1e-2)
    model.eval()
    estimator.eval()
    val_psnr, val_ssim, _ = evaluation(train_set.hwf, model, estimator, None, val_loader, 4096, device,
                                       render_step_size=step_size, white_bkgd=True)
    ckpt = save_checkpoint(model, out_dir)
    load_checkpoint(model, out_dir)
    frames, d_frames = R.render_path(train_set.path_poses[:3], train_set.hwf, train_set.near, train_set.far, 4096,
                                     model, estimator, white_bkgd=True, render_step_size=step_size, device=device)
    res = {"train_psnr_first": float(np.mean(psnrs[:10])), "train_psnr_last": float(np.mean(psnrs[-10:])),
           "val_psnr": float(val_psnr), "val_ssim": float(val_ssim), "lr_final": scheduler.lr,
           "checkpoint": ckpt, "frames": tuple(frames.shape), "d_frames": tuple(d_frames.shape)}
    print(res)
    return res


if __name__ == "__main__":
    main()
