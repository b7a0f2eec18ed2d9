"""Oracle: positional encoding + FreeNeRF frequency mask.  TEST INFRASTRUCTURE ONLY.

positional_encoding restates /root/reference/src/core/models.py:10-50
(PositionalEncoder): cat([x, sin(f0 x), cos(f0 x), ..., sin(f_{L-1} x),
cos(f_{L-1} x)]) with f_k = 2**k (log_space, default) or
linspace(1, 2**(L-1), L); no pi factor, no mask.

freq_mask has NO reference counterpart: canonical
FreeNeRF (Yang et al. 2023) schedule laid out over groups of 3 channels.
"""
import numpy as np
import torch

f32 = np.float32


def frequencies(n_freqs, log_space=True):
    """reference: src/core/models.py:30-34."""
    if log_space:
        return (2.0 ** torch.linspace(0.0, n_freqs - 1, n_freqs)).float()
    return torch.linspace(2.0 ** 0.0, 2.0 ** (n_freqs - 1), n_freqs).float()


def positional_encoding(x, n_freqs, log_space=True):
    """x [N,3] torch f32 -> [N, 3*(1+2L)]; reference: src/core/models.py:43-50."""
    out = [x]
    for f in frequencies(n_freqs, log_space):
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def positional_encoding_doubleangle(x, n_freqs):
    """The CUDA encoder's evaluation order for f_k = 2^k (log_space): an accurate
    sin/cos every 4th octave, double-angle steps sin 2a = 2 s c, cos 2a = 1 - 2 s^2 in
    between.  Equal to positional_encoding(x, n_freqs, True) to ~1e-6 absolute; used by
    the bf16 emulation so that both sides round (almost always) the same values to bf16."""
    out = [x]
    sn = cs = None
    for k in range(n_freqs):
        if k % 4 == 0:
            sn, cs = torch.sin(x * (2.0 ** k)), torch.cos(x * (2.0 ** k))
        else:
            sn, cs = 2.0 * sn * cs, 1.0 - 2.0 * sn * sn
        out.append(sn)
        out.append(cs)
    return torch.cat(out, -1)


def freq_mask(d_out, step, reg_steps, clip=False):
    """FreeNeRF mask over an encoding of d_out channels in groups of 3
    (SURVEY.md Appendix B4).  step >= reg_steps (or reg_steps <= 0) -> ones.

    ptr = min(G*step/reg_steps + 1, G), G = d_out/3, k = floor(ptr):
    mask[:3k] = 1, mask[3k:3k+3] = ptr-k, rest 0.
    """
    G = d_out // 3
    m = np.zeros(d_out, f32)
    if reg_steps <= 0 or step >= reg_steps:
        m[:] = 1.0
    else:
        ptr = min(G * float(step) / float(reg_steps) + 1.0, float(G))
        k = int(np.floor(ptr))
        m[: 3 * k] = 1.0
        if k < G:
            m[3 * k: 3 * k + 3] = f32(ptr - k)
    if clip:
        m = np.clip(m, f32(1e-8), f32(1.0 - 1e-8))
    return m
