"""Evaluation + checkpoint mirror of the reference's run loop.

``evaluation(...)`` follows /root/reference/src/run-nerf.py:108-191: render every validation
frame with ``render_frame`` (the B200 path), PSNR = -10 log10(mse) over the stacked frames,
SSIM averaged over frames with scikit-image's ``structural_similarity(channel_axis=-1,
data_range=1.0, gaussian_weights=True)`` semantics, LPIPS returned as None exactly like the
reference (run-nerf.py:176 overwrites it; its VGG weights are not available offline either).
SSIM runs on the device with torch convolutions (library ops; not part of the timed hot path).

``save_checkpoint`` / ``load_checkpoint``: ``torch.save(model.state_dict(), <out>/model/nn.pt)``
and back (run-nerf.py:417,437) — the 24 reference state-dict keys, loadable by the reference's
``core.models.NeRF`` and vice versa.
"""
import os
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def psnr(mse) -> float:
    return -10.0 * torch.log10(torch.as_tensor(mse))


def _gauss_kernel(sigma=1.5, truncate=3.5, device="cpu", dtype=torch.float64):
    r = int(truncate * sigma + 0.5)
    x = torch.arange(-r, r + 1, device=device, dtype=dtype)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum(), r


def _gauss_filter(x, k, r):
    """separable Gaussian with scipy.ndimage's 'reflect' boundary (d c b a | a b c d | d c b a)
    on the last two dims of [N,C,H,W]"""
    C = x.shape[1]
    x = _pad_symmetric(x, r)
    x = F.conv2d(x, k.view(1, 1, -1, 1).expand(C, 1, -1, 1), groups=C)
    return F.conv2d(x, k.view(1, 1, 1, -1).expand(C, 1, 1, -1), groups=C)


def _pad_symmetric(x, r):
    # torch's 'reflect' excludes the edge sample; scipy's 'reflect' (numpy 'symmetric') repeats it
    top, bot = x[..., :r, :].flip(-2), x[..., -r:, :].flip(-2)
    x = torch.cat([top, x, bot], dim=-2)
    left, right = x[..., :, :r].flip(-1), x[..., :, -r:].flip(-1)
    return torch.cat([left, x, right], dim=-1)


def ssim(img: torch.Tensor, img_gt: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """mean SSIM of [N,H,W,C] image batches, one value per image (scikit-image semantics with
    gaussian_weights=True: sigma 1.5, 11x11 window, population covariance, K1=0.01, K2=0.03,
    borders of (win-1)/2 pixels cropped, mean over pixels and channels)."""
    x = img.permute(0, 3, 1, 2).to(torch.float64)
    y = img_gt.permute(0, 3, 1, 2).to(torch.float64)
    k, r = _gauss_kernel(device=x.device)
    ux, uy = _gauss_filter(x, k, r), _gauss_filter(y, k, r)
    uxx, uyy, uxy = _gauss_filter(x * x, k, r), _gauss_filter(y * y, k, r), _gauss_filter(x * y, k, r)
    vx, vy, vxy = uxx - ux * ux, uyy - uy * uy, uxy - ux * uy
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    return s[..., r:-r, r:-r].mean(dim=(1, 2, 3))


def evaluation(hwf: Tuple[int, int, float], model, estimator, lpips_net, data_loader, chunksize: int,
               device, render_step_size: float = 5e-3, white_bkgd: bool = False):
    """-> (val_psnr, val_ssim, val_lpips=None); same arguments as run-nerf.py:108-117 plus
    ``white_bkgd`` (a module-level ``args`` global in the reference)."""
    from .render.rendering import render_frame
    ds = data_loader.dataset
    rgbs, rgbs_gt = [], []
    with torch.no_grad():
        for rgb_gt, pose in data_loader:
            rgbs_gt.append(rgb_gt)
            rgb, _ = render_frame(hwf, ds.near, ds.far, pose[0], chunksize, estimator, model, train=False,
                                  ndc=ds.ndc, white_bkgd=white_bkgd, render_step_size=render_step_size,
                                  device=device)
            rgbs.append(rgb)
    rgbs = torch.stack(rgbs, dim=0)
    rgbs_gt = torch.cat(rgbs_gt, dim=0).to(rgbs.device)
    val_psnr = psnr(F.mse_loss(rgbs, rgbs_gt))
    val_ssim = ssim(rgbs, rgbs_gt).mean().item()
    return val_psnr, val_ssim, None


def save_checkpoint(model, out_dir: str) -> str:
    os.makedirs(os.path.join(out_dir, "model"), exist_ok=True)
    path = os.path.join(out_dir, "model", "nn.pt")
    torch.save(model.state_dict(), path)
    return path


def load_checkpoint(model, out_dir: str, map_location: Optional[str] = None):
    model.load_state_dict(torch.load(os.path.join(out_dir, "model", "nn.pt"), map_location=map_location))
    return model
