"""Generate tests/golden/reference_sinerf.npz from the REFERENCE's own SiNeRF (build container
only).  TEST INFRASTRUCTURE ONLY.      python -m oracle.gen_golden_sinerf
Seed-42 construction of core.models.SiNeRF(3, 3, 256, [30, 1, ...]) -> state-dict names/shapes/sums,
outputs on seeded points, gradient norms of a fixed scalar loss."""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    sys.path.insert(0, REF)
    from core import models as M
    torch.manual_seed(42)
    model = M.SiNeRF(3, 3, 256, [30.] + [1.] * 7)
    sd = model.state_dict()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(32, 3, generator=g) * 2 - 1
    d = torch.nn.functional.normalize(torch.randn(32, 3, generator=g), dim=-1)
    out = model(x, d)
    sig = model(x)
    loss = (out * torch.linspace(0.1, 1.0, 4)).sum()
    grads = torch.autograd.grad(loss, list(model.parameters()))
    np.savez_compressed(os.path.join(OUT, "reference_sinerf.npz"),
                        names=np.array(list(sd.keys())), shapes=np.array([str(tuple(v.shape)) for v in sd.values()]),
                        w_sum=np.array([v.double().sum().item() for v in sd.values()]),
                        w_abs=np.array([v.double().abs().sum().item() for v in sd.values()]),
                        x=x.numpy(), d=d.numpy(), out=out.detach().numpy(), sigma_only=sig.detach().numpy(),
                        g_norm=np.array([gr.double().norm().item() for gr in grads]))
    print("wrote reference_sinerf.npz:", len(sd), "tensors,", sum(v.numel() for v in sd.values()), "params")


if __name__ == "__main__":
    main()
