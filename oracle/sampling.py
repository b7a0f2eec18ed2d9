"""Oracle: stratified sampling and hierarchical sample_pdf.  TEST INFRASTRUCTURE ONLY.

There is NO reference code for these (the reference
delegates sampling to nerfacc's occupancy grid, src/render/rendering.py:66-74).
The spec is the canonical NeRF definition —
**parity unpinned** — with two choices fixed here so that integer outputs are
bit-exact between this file and the CUDA kernels:

* every product/sum below is an individually rounded fp32 op (no FMA);
* the CDF is built in "warp-tree order": 32 lanes each own E consecutive
  weights, lane-local sums are sequential, the total is an xor-butterfly
  reduction and the cross-lane prefix is a Kogge-Stone scan.  sample_pdf_canonical
  (sequential torch.cumsum, the textbook form) is kept to show the two agree
  except for ulp-level ties.
"""
import numpy as np

f32 = np.float32
EPS_W = f32(1e-5)


def linspace01(n):
    """t_i = i/(n-1) as a correctly rounded fp32 division (n>1)."""
    if n == 1:
        return np.zeros(1, f32)
    return (np.arange(n, dtype=f32) / f32(n - 1)).astype(f32)


def rng_uniform(seed, n):
    """The stateless uniform stream of the seeded samplers (include/fsnerf_b200.h:
    fsnerf_rng_uniform; csrc/rays.cu:rng_uniform).  Element i: the low and high index words mixed
    with the low seed word, murmur3's fmix32, plus the high seed word, the lowbias32 finaliser;
    the top 24 bits scaled by 2^-24.  It stands in for the reference's torch.rand (src/render/
    rendering.py), whose actual stream no test of the reference pins either."""
    seed = int(seed) & ((1 << 64) - 1)
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        lo = (i & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        hi = (i >> np.uint64(32)).astype(np.uint32)
        h = lo * np.uint32(0x9E3779B1) + hi * np.uint32(0x85EBCA77) + np.uint32(seed & 0xFFFFFFFF)
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x85EBCA6B)
        h ^= h >> np.uint32(13)
        h *= np.uint32(0xC2B2AE35)
        h ^= h >> np.uint32(16)
        h += np.uint32(seed >> 32)
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x7FEB352D)
        h ^= h >> np.uint32(15)
        h *= np.uint32(0x846CA68B)
        h ^= h >> np.uint32(16)
    return ((h >> np.uint32(8)).astype(f32) * f32(2.0 ** -24)).astype(f32)


def stratified(n_rays, n_samples, near, far, u=None):
    """Appendix B1.  -> z [R,S] f32 sorted points.

    z = near*(1-t) + far*t; if u is given: mid=(z[1:]+z[:-1])/2,
    upper=cat(mid,z[-1]), lower=cat(z[0],mid), z = lower + (upper-lower)*u.
    """
    t = linspace01(n_samples)
    near, far = f32(near), f32(far)
    z = (near * (f32(1.0) - t)).astype(f32) + (far * t).astype(f32)
    z = np.broadcast_to(z.astype(f32), (n_rays, n_samples)).copy()
    if u is not None:
        u = np.asarray(u, f32)
        mid = (f32(0.5) * (z[:, 1:] + z[:, :-1])).astype(f32)
        upper = np.concatenate([mid, z[:, -1:]], -1)
        lower = np.concatenate([z[:, :1], mid], -1)
        z = (lower + ((upper - lower).astype(f32) * u).astype(f32)).astype(f32)
    return z


def intervals_from_points(z, far):
    """Sorted points -> contiguous intervals fed to the compositor:
    t_starts = z, t_ends = [z[1:], far]; the MLP is queried at the midpoint
    (reference convention, src/render/rendering.py:61,79)."""
    z = np.asarray(z, f32)
    t_ends = np.concatenate([z[:, 1:], np.full((z.shape[0], 1), f32(far), f32)], -1)
    return z, t_ends.astype(f32)


def _warp_tree_cdf(wp):
    """wp [R,n] f32 (already +1e-5) -> cdf [R,n+1] in warp-tree order."""
    R, n = wp.shape
    E = (n + 31) // 32
    pad = np.zeros((R, 32 * E), f32)
    pad[:, :n] = wp
    lanes = pad.reshape(R, 32, E)
    # lane-local sequential sum
    s = lanes[:, :, 0].copy()
    for e in range(1, E):
        s = (s + lanes[:, :, e]).astype(f32)
    # xor-butterfly total (all lanes end with the same value)
    tot = s.copy()
    for m in (16, 8, 4, 2, 1):
        idx = np.arange(32) ^ m
        tot = (tot + tot[:, idx]).astype(f32)
    total = tot[:, :1]  # lane 0 (identical in every lane)
    pdf = (lanes / total[:, :, None]).astype(f32)
    # lane-local inclusive prefix
    loc = np.empty_like(pdf)
    loc[:, :, 0] = pdf[:, :, 0]
    for e in range(1, E):
        loc[:, :, e] = (loc[:, :, e - 1] + pdf[:, :, e]).astype(f32)
    # Kogge-Stone inclusive scan over lane totals
    T = loc[:, :, E - 1].copy()
    for d in (1, 2, 4, 8, 16):
        sh = np.zeros_like(T)
        sh[:, d:] = T[:, :-d]
        Tn = (T + sh).astype(f32)
        Tn[:, :d] = T[:, :d]
        T = Tn
    excl = np.zeros_like(T)
    excl[:, 1:] = T[:, :-1]
    c = (excl[:, :, None] + loc).astype(f32)
    c[:, :, E - 1] = T  # last element of a lane takes the scan value itself
    cdf = np.concatenate([np.zeros((R, 1), f32), c.reshape(R, 32 * E)[:, :n]], -1)
    return cdf.astype(f32)


def sample_pdf(z_coarse, w_coarse, n_fine, far, u=None):
    """Appendix B2 in warp-tree order.

    z_coarse [R,Sc] sorted points, w_coarse [R,Sc] compositing weights,
    u [R,Sf] in [0,1) or None (deterministic linspace).
    -> dict(samples [R,Sf] f32, inds [R,Sf] i32, below, above,
            perm [R,Sc+Sf] i32 (stable sort permutation of cat(z_c, samples)),
            z [R,Sc+Sf] f32 sorted, t_starts, t_ends)
    """
    z_coarse = np.asarray(z_coarse, f32)
    w_coarse = np.asarray(w_coarse, f32)
    R, Sc = z_coarse.shape
    bins = (f32(0.5) * (z_coarse[:, 1:] + z_coarse[:, :-1])).astype(f32)  # [R,Sc-1]
    # weights are clamped at 0: the reference's raw (un-activated) sigma can make
    # compositing weights negative, which is not a density a CDF can be built from
    wp = (np.maximum(w_coarse[:, 1:-1], f32(0.0)) + EPS_W).astype(f32)   # [R,Sc-2]
    cdf = _warp_tree_cdf(wp)                                              # [R,Sc-1]
    nb = Sc - 1
    if u is None:
        u = np.broadcast_to(linspace01(n_fine), (R, n_fine)).copy()
    u = np.asarray(u, f32)
    # right=True: number of cdf entries <= u
    inds = (cdf[:, None, :] <= u[:, :, None]).sum(-1).astype(np.int32)
    below = np.maximum(inds - 1, 0).astype(np.int32)
    above = np.minimum(inds, nb - 1).astype(np.int32)
    ar = np.arange(R)[:, None]
    c0, c1 = cdf[ar, below], cdf[ar, above]
    b0, b1 = bins[ar, below], bins[ar, above]
    denom = (c1 - c0).astype(f32)
    denom = np.where(denom < EPS_W, f32(1.0), denom).astype(f32)
    t = ((u - c0).astype(f32) / denom).astype(f32)
    samples = (b0 + (t * (b1 - b0).astype(f32)).astype(f32)).astype(f32)
    cat = np.concatenate([z_coarse, samples], -1)
    perm = np.argsort(cat, axis=-1, kind="stable").astype(np.int32)
    z = np.take_along_axis(cat, perm, -1)
    t_starts, t_ends = intervals_from_points(z, far)
    return dict(samples=samples, inds=inds, below=below, above=above, cdf=cdf,
                perm=perm, z=z, t_starts=t_starts, t_ends=t_ends)


def sample_pdf_canonical(bins, weights, u):
    """Textbook sample_pdf (Mildenhall et al. 2020 reference implementation,
    SURVEY.md Appendix B2) with a sequential torch.cumsum — used only to show
    the warp-tree CDF selects the same bins.  -> (samples, inds)"""
    import torch
    bins = torch.as_tensor(bins)
    w = torch.as_tensor(weights) + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.as_tensor(u).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    b0, b1 = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = c1 - c0
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    samples = b0 + (u - c0) / denom * (b1 - b0)
    return samples.numpy(), inds.numpy().astype(np.int32)
