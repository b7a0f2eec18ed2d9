"""Oracle: NeRF MLP forward (autograd gives the backward).  TEST INFRASTRUCTURE ONLY.

Functional restatement of /root/reference/src/core/models.py:53-143 over a
state dict with the reference's 24 keys (layers.{0..7}.{weight,bias},
sigma.*, connection.*, branch.*, rgb.*).  fp32, torch-CPU.  With masks == None
it must reproduce reference NeRF.forward bit-for-bit on the same host
(tests/golden/reference_mlp.npz).
"""
import torch
import torch.nn.functional as F

from .encoding import positional_encoding


def init_state_dict(n_layers=8, d_hidden=256, skip=(4,), n_freqs=10,
                    n_freqs_dir=4, seed=42, d_pos=3, d_dir=3):
    """Default nn.Linear init in the reference's construction order
    (src/core/models.py:95-109) under torch.manual_seed(seed)
    (src/run-nerf.py:35-36)."""
    d_pe = d_pos * (1 + 2 * n_freqs)
    d_de = d_dir * (1 + 2 * n_freqs_dir)
    sd = {}
    with torch.random.fork_rng():
        torch.manual_seed(seed)

        def linear(name, n_in, n_out):
            m = torch.nn.Linear(n_in, n_out)
            sd[f"{name}.weight"] = m.weight.detach().clone()
            sd[f"{name}.bias"] = m.bias.detach().clone()

        # RNG order of the reference ctor: hidden layers 1..n-1 are built
        # first (:96-99), THEN layers.0 (:100-102), then the heads (:105-108).
        for i in range(n_layers - 1):
            linear(f"layers.{i + 1}", d_hidden + d_pe if i in skip else d_hidden, d_hidden)
        linear("layers.0", d_pe, d_hidden)
        linear("sigma", d_hidden, 1)
        linear("connection", d_hidden, d_hidden)
        linear("branch", d_hidden + d_de, d_hidden // 2)
        linear("rgb", d_hidden // 2, 3)
    # hand back in state_dict() key order (module registration order)
    names = [f"layers.{i}" for i in range(n_layers)] + ["sigma", "connection", "branch", "rgb"]
    return {f"{n}.{k}": sd[f"{n}.{k}"] for n in names for k in ("weight", "bias")}


def nerf_forward(sd, x, dirs=None, n_layers=8, skip=(4,), n_freqs=10,
                 n_freqs_dir=4, log_space=True, mask_pos=None, mask_dir=None):
    """x [N,3], dirs [N,3]|None -> [N,4]=(rgb,sigma) or [N,1]=sigma.

    reference: src/core/models.py:111-143.  sigma has NO activation (:127,141);
    skip concat order [h, PE(x)] after layer i in skip (:122-123); branch input
    order [connection(h), PE(dirs)] (:130-133); rgb sigmoid (:135-136).
    mask_pos/mask_dir: FreeNeRF multiplicative masks on the encodings
    (SURVEY.md Appendix B4); None == all ones == the reference.
    """
    x_in = positional_encoding(x, n_freqs, log_space)
    if mask_pos is not None:
        x_in = x_in * mask_pos
    h = x_in
    for i in range(n_layers):
        h = F.relu(F.linear(h, sd[f"layers.{i}.weight"], sd[f"layers.{i}.bias"]))
        if i in skip:
            h = torch.cat([h, x_in], -1)
    sigma = F.linear(h, sd["sigma.weight"], sd["sigma.bias"])
    if dirs is None:
        return sigma
    h = F.linear(h, sd["connection.weight"], sd["connection.bias"])
    d_in = positional_encoding(dirs, n_freqs_dir, log_space)
    if mask_dir is not None:
        d_in = d_in * mask_dir
    h = torch.cat([h, d_in], -1)
    h = F.relu(F.linear(h, sd["branch.weight"], sd["branch.bias"]))
    rgb = torch.sigmoid(F.linear(h, sd["rgb.weight"], sd["rgb.bias"]))
    return torch.cat([rgb, sigma], -1)
