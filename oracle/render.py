"""Oracle: the whole ray-march path (render + one train step).  TEST INFRASTRUCTURE ONLY.

Follows the orchestration of /root/reference/src/render/rendering.py:25-107
(sample -> query MLP at interval midpoints with per-sample dirs -> composite
-> background) and the train-step arithmetic of
/root/reference/src/run-nerf.py:216-223,255-258,282-285 (Adam defaults, mean
MSE, psnr = -10 log10 loss), with the occupancy-grid sampler replaced by
stratified + sample_pdf as BASELINE.json's north_star prescribes
(SURVEY.md §8 row a7).  torch-CPU fp32.
"""
import numpy as np
import torch

from . import sampling
from .compositing import composite_dense
from .mlp import nerf_forward


def query_mlp(sd, rays_o, rays_d, t_starts, t_ends, mask_pos=None, mask_dir=None, **kw):
    """reference: src/render/rendering.py:76-84 — x = o + d*(ts+te)/2, dirs = d
    gathered per sample -> raw [R,S,4]."""
    R, S = t_starts.shape
    tm = (t_starts + t_ends) / 2.0
    x = rays_o[:, None, :] + rays_d[:, None, :] * tm[..., None]
    d = rays_d[:, None, :].expand(R, S, 3)
    raw = nerf_forward(sd, x.reshape(-1, 3), d.reshape(-1, 3),
                       mask_pos=mask_pos, mask_dir=mask_dir, **kw)
    return raw.reshape(R, S, 4)


def render_rays_hier(sd_coarse, sd_fine, rays_o, rays_d, near, far, n_coarse,
                     n_fine, u_strat=None, u_pdf=None, white_bkgd=False,
                     mask_pos=None, mask_dir=None, **kw):
    """Coarse (stratified) -> sample_pdf -> fine.  All tensors torch f32.
    n_fine == 0 gives the coarse-only path (config C1).  Returns a dict."""
    rays_o = torch.as_tensor(rays_o, dtype=torch.float32)
    rays_d = torch.as_tensor(rays_d, dtype=torch.float32)
    R = rays_o.shape[0]
    bk = torch.ones(3) if white_bkgd else None
    mp = None if mask_pos is None else torch.as_tensor(mask_pos)
    md = None if mask_dir is None else torch.as_tensor(mask_dir)
    z_c = sampling.stratified(R, n_coarse, near, far,
                              None if u_strat is None else np.asarray(u_strat))
    ts_c, te_c = sampling.intervals_from_points(z_c, far)
    ts_c, te_c = torch.from_numpy(ts_c), torch.from_numpy(te_c)
    raw_c = query_mlp(sd_coarse, rays_o, rays_d, ts_c, te_c, mp, md, **kw)
    rgb_c, op_c, dp_c, w_c, _, _ = composite_dense(raw_c, ts_c, te_c, bk)
    out = dict(rgb_coarse=rgb_c, opacity_coarse=op_c, depth_coarse=dp_c,
               weights_coarse=w_c, raw_coarse=raw_c, t_starts_coarse=ts_c,
               t_ends_coarse=te_c, z_coarse=torch.from_numpy(z_c))
    if n_fine == 0:
        out.update(rgb=rgb_c, opacity=op_c, depth=dp_c, weights=w_c,
                   t_starts=ts_c, t_ends=te_c, raw=raw_c)
        return out
    sp = sampling.sample_pdf(z_c, w_c.detach().numpy(), n_fine, far,
                             None if u_pdf is None else np.asarray(u_pdf))
    ts_f, te_f = torch.from_numpy(sp["t_starts"]), torch.from_numpy(sp["t_ends"])
    raw_f = query_mlp(sd_fine, rays_o, rays_d, ts_f, te_f, mp, md, **kw)
    rgb_f, op_f, dp_f, w_f, _, _ = composite_dense(raw_f, ts_f, te_f, bk)
    out.update(rgb=rgb_f, opacity=op_f, depth=dp_f, weights=w_f, t_starts=ts_f,
               t_ends=te_f, raw=raw_f, pdf=sp)
    return out


def train_step(sd_coarse, sd_fine, opt_state, rays_o, rays_d, rgb_gt, near, far,
               n_coarse, n_fine, u_strat, u_pdf, lr, white_bkgd=False,
               mask_pos=None, mask_dir=None, betas=(0.9, 0.999), eps=1e-8,
               occ_reg=None, weight_reg=None, **kw):
    """One optimisation step in place on the state dicts.
    loss = mse(rgb_fine, gt) [+ mse(rgb_coarse, gt) when n_fine > 0];
    Adam exactly as torch.optim.Adam defaults (src/run-nerf.py:216-217).
    opt_state: dict(step=int, m={...}, v={...}) keyed 'c.<name>' / 'f.<name>'.
    occ_reg=(a,b,func): + OcclusionRegularizer of the output pass (src/run-nerf.py:260-264);
    weight_reg=(mode, alpha): + alpha * weight penalty over every network (run-nerf.py:266-279).
    -> (loss, psnr_fine, grads dict)"""
    from . import regularizers as oreg
    params = {}
    for k, v in sd_coarse.items():
        params["c." + k] = v.requires_grad_(True)
    if n_fine > 0:
        for k, v in sd_fine.items():
            params["f." + k] = v.requires_grad_(True)
    out = render_rays_hier(sd_coarse, sd_fine, rays_o, rays_d, near, far, n_coarse,
                           n_fine, u_strat, u_pdf, white_bkgd, mask_pos, mask_dir, **kw)
    gt = torch.as_tensor(rgb_gt, dtype=torch.float32)
    loss_f = torch.nn.functional.mse_loss(out["rgb"], gt)
    loss = loss_f
    if n_fine > 0:
        loss = loss + torch.nn.functional.mse_loss(out["rgb_coarse"], gt)
    if occ_reg is not None:
        loss = loss + oreg.occlusion_reg_dense(out["raw"][..., 3], out["t_starts"], out["t_ends"], *occ_reg)
    if weight_reg is not None:
        nets = [sd_coarse] + ([sd_fine] if n_fine > 0 else [])
        loss = loss + weight_reg[1] * sum(oreg.weight_reg(sd, weight_reg[0]) for sd in nets)
    grads = torch.autograd.grad(loss, list(params.values()))
    grads = dict(zip(params.keys(), grads))
    opt_state["step"] += 1
    t = opt_state["step"]
    b1, b2 = betas
    with torch.no_grad():
        for k, p in params.items():
            g = grads[k]
            m = opt_state["m"].setdefault(k, torch.zeros_like(p))
            v = opt_state["v"].setdefault(k, torch.zeros_like(p))
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1 = 1 - b1 ** t
            bc2 = 1 - b2 ** t
            denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
            p.addcdiv_(m, denom, value=-lr / bc1)
    for v in params.values():
        v.requires_grad_(False)
    psnr = -10.0 * torch.log10(loss_f.detach()).item()
    return loss.item(), psnr, grads
