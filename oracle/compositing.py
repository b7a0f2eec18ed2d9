"""Oracle: alpha compositing (volume rendering).  TEST INFRASTRUCTURE ONLY.

``render_packed`` restates nerfacc.volrend.rendering v0.5.3 (third-party,
pinned at /root/reference/environment.yaml:341; source NOT on the box — restated
from its published algorithm) as called at
/root/reference/src/render/rendering.py:89-96.  **parity unpinned** against
nerfacc itself; pinned by closed-form known answers in tests/test_oracle.py.

``composite_dense`` is the same arithmetic on a dense [R,S] layout (fixed
samples per ray) plus the canonical raw2outputs switches of Appendix B3.
torch-CPU fp32; autograd supplies the backward.
"""
import torch

EPS = torch.finfo(torch.float32).eps


def exclusive_sum_packed(x, ray_indices, n_rays):
    """Per-ray exclusive prefix sum over a packed sample list."""
    out = torch.zeros_like(x)
    if x.numel() == 0:
        return out
    csum = torch.cumsum(x, 0)
    # offset of each ray's first sample
    first = torch.ones_like(ray_indices, dtype=torch.bool)
    first[1:] = ray_indices[1:] != ray_indices[:-1]
    start_idx = torch.nonzero(first).squeeze(-1)
    base = torch.zeros(x.shape[0], dtype=x.dtype)
    seg_base = torch.cat([torch.zeros(1, dtype=x.dtype), csum[start_idx[1:] - 1]])
    seg_id = torch.cumsum(first.to(torch.int64), 0) - 1
    base = seg_base[seg_id]
    return csum - x - base


def render_packed(t_starts, t_ends, ray_indices, n_rays, rgbs, sigmas, render_bkgd=None):
    """nerfacc.volrend.rendering restated (Appendix A5).

    -> (colors[R,3], opacities[R,1], depths[R,1], extras{weights,alphas,trans,sigmas,rgbs})
    sigma is used raw (no activation, reference src/core/models.py:127).
    """
    sd = sigmas * (t_ends - t_starts)
    alphas = 1.0 - torch.exp(-sd)
    trans = torch.exp(-exclusive_sum_packed(sd, ray_indices, n_rays))
    weights = trans * alphas
    colors = torch.zeros(n_rays, 3).index_add_(0, ray_indices, weights[:, None] * rgbs)
    opac = torch.zeros(n_rays, 1).index_add_(0, ray_indices, weights[:, None])
    depth = torch.zeros(n_rays, 1).index_add_(
        0, ray_indices, weights[:, None] * ((t_starts + t_ends)[:, None] / 2.0))
    depth = depth / opac.clamp_min(EPS)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opac)
    return colors, opac, depth, dict(weights=weights, alphas=alphas, trans=trans,
                                     sigmas=sigmas, rgbs=rgbs)


def composite_dense(raw, t_starts, t_ends, bkgd=None, sigma_relu=False,
                    delta_scale=None, normalize_depth=True, product_trans=False):
    """raw [R,S,4]=(rgb,sigma), t_starts/t_ends [R,S] ->
    (rgb[R,3], opacity[R,1], depth[R,1], weights[R,S], alphas, trans).

    Defaults == render_packed on ray_indices = repeat_interleave(arange(R), S).
    Switches (Appendix B3, canonical raw2outputs): sigma_relu, delta_scale
    ([R] ||d||), normalize_depth=False (depth = sum w t), product_trans
    (T_i = prod_{j<i}(1-alpha_j+1e-10) instead of exp(-sum)).
    """
    rgbs, sig = raw[..., :3], raw[..., 3]
    delta = t_ends - t_starts
    if delta_scale is not None:
        delta = delta * delta_scale[:, None]
    if sigma_relu:
        sig = torch.relu(sig)
    sd = sig * delta
    alphas = 1.0 - torch.exp(-sd)
    if product_trans:
        one_m = 1.0 - alphas + 1e-10
        trans = torch.cumprod(torch.cat([torch.ones_like(one_m[:, :1]), one_m[:, :-1]], -1), -1)
    else:
        excl = torch.cumsum(sd, -1) - sd
        trans = torch.exp(-excl)
    weights = trans * alphas
    rgb = (weights[..., None] * rgbs).sum(1)
    opac = weights.sum(1, keepdim=True)
    tmid = (t_starts + t_ends) / 2.0
    depth = (weights * tmid).sum(1, keepdim=True)
    if normalize_depth:
        depth = depth / opac.clamp_min(EPS)
    if bkgd is not None:
        rgb = rgb + bkgd * (1.0 - opac)
    return rgb, opac, depth, weights, alphas, trans
