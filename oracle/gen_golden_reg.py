"""Generate tests/golden/reference_reg.npz from the REFERENCE's own Python (build container
only; needs /root/reference).  TEST INFRASTRUCTURE ONLY.

    python -m oracle.gen_golden_reg

* ``core.loss.OcclusionRegularizer`` (imported unmodified) on seeded packed inputs, value and
  d/d(sigma) by autograd, both ``func`` variants, with rays of unequal length and an empty ray;
* the weight-penalty loop of src/run-nerf.py:266-279 executed over ``named_parameters()`` of a
  reference ``core.models.NeRF`` (seed 42): which tensors it covers, its value and its
  gradient w.r.t. two of them, for 'l1' and the Frobenius branch.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    sys.path.insert(0, REF)
    from core import loss as L, models as M
    out = {}
    g = torch.Generator().manual_seed(7)
    counts = [5, 0, 3, 9, 1, 7]  # samples per ray (ray 1 has none)
    ri = torch.repeat_interleave(torch.arange(len(counts)), torch.tensor(counts))
    t = 2.0 + 4.0 * torch.rand(len(ri), generator=g)
    sig = torch.randn(len(ri), generator=g).requires_grad_(True)
    out["occ_ray_idx"], out["occ_t"], out["occ_sigma"] = ri.numpy(), t.numpy(), sig.detach().numpy()
    for func, a, b in (("linear", 0.5, 2.0), ("exp", 1.5, 0.7)):
        val = L.OcclusionRegularizer(a, b, func)(sig, t, ri)
        (gs,) = torch.autograd.grad(val, sig)
        out[f"occ_{func}_ab"] = np.array([a, b], np.float32)
        out[f"occ_{func}_value"] = val.detach().numpy()
        out[f"occ_{func}_dsigma"] = gs.numpy()
    # dense case shaped like the hot path
    R, S = 6, 11
    ts = torch.sort(2 + 4 * torch.rand(R, S + 1, generator=g), -1).values
    sg = torch.randn(R, S, generator=g).requires_grad_(True)
    tv = ((ts[:, :-1] + ts[:, 1:]) / 2).reshape(-1)
    val = L.OcclusionRegularizer(0.5, 2.0, "linear")(sg.reshape(-1), tv, torch.arange(R).repeat_interleave(S))
    (gs,) = torch.autograd.grad(val, sg)
    out["occ_dense_edges"], out["occ_dense_sigma"] = ts.numpy(), sg.detach().numpy()
    out["occ_dense_value"], out["occ_dense_dsigma"] = val.detach().numpy(), gs.numpy()

    torch.manual_seed(42)
    model = M.NeRF(3, 3, 8, 256, [4], pos_fn={"n_freqs": 10, "log_space": True},
                   dir_fn={"n_freqs": 4, "log_space": True})
    covered = []
    for mode in ("l1", "l2"):
        freq_reg = torch.tensor(0.0)
        for name, param in model.named_parameters():  # run-nerf.py:271-277
            if "weight" in name and param.shape[0] > 3:
                if mode == "l1":
                    covered.append(name)
                    freq_reg += torch.abs(param).sum()
                else:
                    freq_reg += torch.square(param).sum().sqrt()
        grads = torch.autograd.grad(freq_reg, [model.layers[5].weight, model.branch.weight])
        out[f"wreg_{mode}_value"] = freq_reg.detach().numpy()
        # slices + checksums keep the fixture small
        out[f"wreg_{mode}_grad_layers5"] = grads[0][:6, :40].numpy()
        out[f"wreg_{mode}_grad_branch"] = grads[1][:6, :40].numpy()
        out[f"wreg_{mode}_grad_abs_sums"] = np.array([grads[0].abs().sum().item(), grads[1].abs().sum().item()])
    out["wreg_covered"] = np.array(covered)
    out["wreg_all_names"] = np.array([n for n, _ in model.named_parameters()])
    np.savez_compressed(os.path.join(OUT, "reference_reg.npz"), **out)
    print("wrote reference_reg.npz:", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
