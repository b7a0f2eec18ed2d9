"""Oracle: the in-step regularisers of the reference's train loop.  TEST INFRASTRUCTURE ONLY.

* occlusion regulariser — /root/reference/src/core/loss.py:26-60 as called at
  /root/reference/src/run-nerf.py:260-264 (``loss += occ_reg(sigmas, t_vals,
  ray_indices)``: NOT multiplied by ``args.beta``, which only gates the branch);
* weight-norm ("frequency") penalty — /root/reference/src/run-nerf.py:266-279:
  over ``model.named_parameters()`` with ``"weight" in name and param.shape[0] > 3``,
  ``'l1'``: sum|w|, otherwise the Frobenius norm per tensor; ``loss += alpha * freq_reg``,
  gated by ``k < int(reg_ratio * Td)``.

PINNED: ``oracle/gen_golden_reg.py`` runs the reference's own ``core.loss`` and a
reference ``core.models.NeRF`` in the build container and commits the outputs as
``tests/golden/reference_reg.npz``.
"""
import torch


def occlusion_weights(t_vals, a, b, func="linear"):
    """loss.py:48-60"""
    if func == "linear":
        return -a * t_vals + b
    if func == "exp":
        return a * torch.exp(-b * t_vals)
    raise ValueError(f"Unknown occlusion regularizer type: {func}")


def occlusion_reg(sigmas, t_vals, ray_idxs, a, b, func="linear"):
    """loss.py:26-46: per ray present in ``ray_idxs`` (consecutive-unique), the sum of
    w(t)*sigma over its samples; mean over those rays.  Segment sums instead of the
    reference's Python loop over boolean masks — same value."""
    assert a >= 0 and b >= 0
    sigmas, t_vals = torch.as_tensor(sigmas), torch.as_tensor(t_vals)
    ray_idxs = torch.as_tensor(ray_idxs)
    uniq, inv = torch.unique_consecutive(ray_idxs, return_inverse=True)
    per_ray = torch.zeros(len(uniq), dtype=sigmas.dtype).index_add(
        0, inv, occlusion_weights(t_vals, a, b, func) * sigmas)
    return per_ray.mean()


def occlusion_reg_dense(sigma, t_starts, t_ends, a, b, func="linear"):
    """dense [R,S] layout of the hot path: t_vals = (t_starts+t_ends)/2
    (src/render/rendering.py:105), every ray has S samples."""
    t = (t_starts + t_ends) / 2.0
    return (occlusion_weights(t, a, b, func) * sigma).sum(-1).mean()


def regularised_names(named_shapes):
    """run-nerf.py:272-273"""
    return [n for n, shp in named_shapes if "weight" in n and shp[0] > 3]


def weight_reg(state_dict, mode="l1"):
    """run-nerf.py:266-279 (without the alpha factor)"""
    total = torch.zeros(())
    for n in regularised_names([(k, tuple(v.shape)) for k, v in state_dict.items()]):
        p = state_dict[n]
        total = total + (p.abs().sum() if mode == "l1" else p.square().sum().sqrt())
    return total
