"""Oracle: ray generation, NDC warp, chunking.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/src/utils/utilities.py:36-134 in numpy fp32.
Every arithmetic step is kept in float32 in the same order as the reference so
results agree to the last bit on the same host (checked against
tests/golden/reference_rays.npz, which was produced by the reference itself).
"""
import numpy as np

f32 = np.float32


def get_rays(pose, hwf):
    """reference: src/utils/utilities.py:36-82.

    pose [3or4,4] f32, hwf=(H,W,focal) -> (origins[H,W,3], dirs[H,W,3]) f32.
    pixel (h,w): dir_c=((w-W/2)/f, -(h-H/2)/f, -1) normalised (:72), then
    dir_w[a] = sum_b dir_c[b]*pose[a,b] (:75-78); origin = pose[:3,-1] (:80).
    """
    H, W, focal = hwf
    pose = np.asarray(pose, dtype=f32)
    i = np.arange(W, dtype=f32)[None, :].repeat(H, 0)
    j = np.arange(H, dtype=f32)[:, None].repeat(W, 1)
    fx = f32(focal)
    x = (i - f32(W * 0.5)) / fx
    y = -(j - f32(H * 0.5)) / fx
    z = -np.ones_like(i)
    dirs = np.stack([x, y, z], -1)
    # torch.norm(dirs, dim=-1): sqrt(x*x + y*y + z*z) accumulated in order
    nrm = np.sqrt((dirs[..., 0] * dirs[..., 0] + dirs[..., 1] * dirs[..., 1])
                  + dirs[..., 2] * dirs[..., 2]).astype(f32)
    dirs = dirs / nrm[..., None]
    R = pose[:3, :3]
    prod = dirs[..., None, :] * R  # [H,W,3(a),3(b)]
    dirs_w = (prod[..., 0] + prod[..., 1]) + prod[..., 2]
    origins = np.broadcast_to(pose[:3, -1], dirs_w.shape).copy()
    return origins.astype(f32), dirs_w.astype(f32)


def to_ndc(rays_o, rays_d, hwf, near):
    """reference: src/utils/utilities.py:84-120 (called with near=1.0 at
    src/render/rendering.py:150).  NDC directions are NOT renormalised."""
    H, W, focal = hwf
    rays_o = np.asarray(rays_o, f32)
    rays_d = np.asarray(rays_d, f32)
    near = f32(near)
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    sx = f32(-1.0 / (W / (2.0 * focal)))
    sy = f32(-1.0 / (H / (2.0 * focal)))
    o0 = sx * rays_o[..., 0] / rays_o[..., 2]
    o1 = sy * rays_o[..., 1] / rays_o[..., 2]
    o2 = f32(1.0) + f32(2.0) * near / rays_o[..., 2]
    d0 = sx * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = sy * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = f32(-2.0) * near / rays_o[..., 2]
    return (np.stack([o0, o1, o2], -1).astype(f32),
            np.stack([d0, d1, d2], -1).astype(f32))


def get_chunks(n, chunksize):
    """reference: src/utils/utilities.py:122-134 — returns (start, stop) pairs
    of the slices ``inputs[i:i+chunksize]`` (last one ragged)."""
    return [(i, min(i + chunksize, n)) for i in range(0, n, chunksize)]


def rays_from_pixel_ids(poses, hwf, pixel_ids, ndc=False, ndc_near=1.0):
    """Batch form used by the CUDA ray generator: global pixel id
    p = view*H*W + h*W + w (the flattened ray-table order of
    src/nerfdata/datasets/llff.py:59-90) -> per-ray (o, d)."""
    H, W, _ = hwf
    poses = np.asarray(poses, f32)
    pixel_ids = np.asarray(pixel_ids, np.int64)
    view = pixel_ids // (H * W)
    rem = pixel_ids % (H * W)
    o_out = np.empty((len(pixel_ids), 3), f32)
    d_out = np.empty((len(pixel_ids), 3), f32)
    for v in np.unique(view):
        o, d = get_rays(poses[v], hwf)
        o = o.reshape(-1, 3)
        d = d.reshape(-1, 3)
        if ndc:
            o, d = to_ndc(o, d, hwf, ndc_near)
        m = view == v
        o_out[m] = o[rem[m]]
        d_out[m] = d[rem[m]]
    return o_out, d_out
