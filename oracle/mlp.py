"""Oracle: NeRF MLP forward (autograd gives the backward).  TEST INFRASTRUCTURE ONLY.

Functional restatement of /root/reference/src/core/models.py:53-143 over a
state dict with the reference's 24 keys (layers.{0..7}.{weight,bias},
sigma.*, connection.*, branch.*, rgb.*).  fp32, torch-CPU.  With masks == None
it must reproduce reference NeRF.forward bit-for-bit on the same host
(tests/golden/reference_mlp.npz).
"""
import torch
import torch.nn.functional as F

from .encoding import positional_encoding, positional_encoding_doubleangle


def init_state_dict(n_layers=8, d_hidden=256, skip=(4,), n_freqs=10,
                    n_freqs_dir=4, seed=42, d_pos=3, d_dir=3):
    """Default nn.Linear init in the reference's construction order
    (src/core/models.py:95-109) under torch.manual_seed(seed)
    (src/run-nerf.py:35-36)."""
    d_pe = d_pos * (1 + 2 * n_freqs)
    d_de = d_dir * (1 + 2 * n_freqs_dir)
    sd = {}
    with torch.random.fork_rng(devices=[]):
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


# --------------------------------------------------------------------------
# bf16-operand emulation of the CUDA path (tests only).  Same math as
# nerf_forward but rounded to bfloat16 exactly where the kernels round:
# encodings, every GEMM input activation (the stash images), the weights fed to
# the tensor cores, and d(pre-activation) in the backward.  Accumulation stays
# fp32.  This isolates kernel bugs from the (expected) bf16-vs-fp32 gap.
class _RoundGradBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _r(x):
    """round to bf16, straight-through gradient"""
    return x + (x.bfloat16().float() - x).detach()


def nerf_forward_bf16emu(sd, x, dirs, n_layers=8, skip=(4,), n_freqs=10, n_freqs_dir=4,
                         log_space=True, mask_pos=None, mask_dir=None):
    def lin(h_r, name):
        pre = F.linear(h_r, _r(sd[f"{name}.weight"]), sd[f"{name}.bias"])
        return _RoundGradBf16.apply(pre)

    enc = (lambda v, L: positional_encoding_doubleangle(v, L)) if log_space else (
        lambda v, L: positional_encoding(v, L, log_space))
    x_in = enc(x, n_freqs)
    if mask_pos is not None:
        x_in = x_in * mask_pos
    x_in = x_in.bfloat16().float()
    h_r = x_in
    h = None
    for i in range(n_layers):
        h = F.relu(lin(h_r, f"layers.{i}"))
        h_r = _r(h)
        # kernel's ReLU mask is (bf16(h) != 0): kill the gradient where rounding flushed to 0
        h_r = h_r * (h_r.detach() != 0)
        if i in skip:
            h_r = torch.cat([h_r, x_in], -1)
    sigma = F.linear(h, sd["sigma.weight"], sd["sigma.bias"])  # fp32 head on CUDA cores
    c_r = _r(lin(h_r, "connection"))
    d_in = enc(dirs, n_freqs_dir)
    if mask_dir is not None:
        d_in = d_in * mask_dir
    d_in = d_in.bfloat16().float()
    hb = F.relu(lin(torch.cat([c_r, d_in], -1), "branch"))
    rgb = torch.sigmoid(F.linear(hb, sd["rgb.weight"], sd["rgb.bias"]))
    return torch.cat([rgb, sigma], -1)
