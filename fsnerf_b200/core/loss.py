"""Drop-in for the reference's ``core.loss`` (src/core/loss.py:7-60).

``OcclusionRegularizer(a, b, func)(sigmas, t_vals, ray_idxs)``: same constructor, call
signature, assertions and error text.  The reference loops over the unique rays in Python with
a boolean mask per ray (O(rays x samples)); this mirror computes the same per-ray sums with one
segmented ``index_add`` on whatever device the tensors live on, so it is differentiable through
autograd exactly like the original.  Inside the fused engine (``engine.HotPath.train_step(
occ_reg=...)``) the term is instead folded into the compositing-backward kernel
(``fsnerf_composite_backward_occ``)."""
import torch
from torch import Tensor


class OcclusionRegularizer():
    def __init__(self, a: float, b: float, func: str = 'linear'):
        assert a >= 0, 'a should be non-negative'
        self.a = a
        assert b >= 0, 'b should be non-negative'
        self.b = b
        self.func = func

    def __call__(self, sigmas: Tensor, t_vals: Tensor, ray_idxs: Tensor) -> Tensor:
        uniques, inv = torch.unique_consecutive(ray_idxs, return_inverse=True)
        per_ray = torch.zeros(len(uniques), dtype=sigmas.dtype, device=sigmas.device).index_add(
            0, inv, self._weights(t_vals) * sigmas)
        return torch.mean(per_ray)

    def _weights(self, t_vals: Tensor) -> Tensor:
        if self.func == 'linear':
            return -self.a * t_vals + self.b
        if self.func == 'exp':
            return self.a * torch.exp(-self.b * t_vals)
        raise ValueError(f'Unknown occlusion regularizer type: {self.func}')
