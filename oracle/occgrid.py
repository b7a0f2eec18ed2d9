"""Oracle: occupancy-grid sampler and grid update.  TEST INFRASTRUCTURE ONLY.

The reference samples with ``nerfacc.estimators.occ_grid.OccGridEstimator`` (nerfacc 0.5.3,
/root/reference/environment.yaml:341; call sites src/render/rendering.py:66-74 and
src/run-nerf.py:92-98,288-295).  Its source is NOT on the box and it has no tests or golden
vectors in the reference — **parity unpinned**.  This file states the semantics the CUDA kernels
(fsnerf_b200/csrc/occgrid.cu) implement, following nerfacc's published behaviour:

* ``march``: per ray, slab test against the outermost level's box; t_begin = the first point of
  the lattice near + j*step — near = the ray's near plane (+ U[0,1)*step when stratified) — at or
  after the entry, t_limit = min(far plane, exit); candidate interval
  k = [t_begin + k*step, +step) is emitted iff its midpoint is < t_limit and lies in an occupied
  cell of the finest level whose box contains it (cell = floor((p - min)/(max - min) * res),
  x-major / z-fastest).  Output is ray-major packed.  fp32 arithmetic in the kernel's order
  (fma for t_k and for o + d*t) so that the emitted sample SET is compared bit for bit.
* ``visibility``: trans = exp(-exclusive_sum(sigma*delta)) per ray; keep trans >= early_stop_eps
  (and alpha >= alpha_thre when alpha_thre > 0).
* ``update``: occs[c] = max(decay*occs[c], max of the candidates in c); threshold =
  min(mean(occs[occs >= 0]), occ_thre); binaries = occs > threshold.
"""
import numpy as np

f32 = np.float32


def level_aabbs(roi_aabb, levels):
    roi = np.asarray(roi_aabb, f32)
    c, h = (roi[:3] + roi[3:]) / f32(2), (roi[3:] - roi[:3]) / f32(2)
    return np.stack([np.concatenate([c - h * f32(2 ** l), c + h * f32(2 ** l)]) for l in range(levels)]).astype(f32)


def _fma(a, b, c):
    """fp32 fused multiply-add (exact product in fp64, one rounding)"""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def occupied(binaries, aabbs, p):
    """p [n,3] fp32 -> bool [n]"""
    levels, res = binaries.shape[0], binaries.shape[1]
    out = np.zeros(len(p), bool)
    done = np.zeros(len(p), bool)
    for l in range(levels):
        b = aabbs[l]
        inside = np.all((p >= b[:3]) & (p <= b[3:]), -1) & ~done
        f = (p - b[:3]) / (b[3:] - b[:3]) * f32(res)
        idx = np.clip(np.floor(f).astype(np.int64), 0, res - 1)
        occ = binaries[l, idx[:, 0], idx[:, 1], idx[:, 2]]
        out[inside] = occ[inside]
        done |= inside
    return out


def march(rays_o, rays_d, binaries, aabbs, step, near=0.0, far=1e10, near_planes=None):
    rays_o, rays_d = np.asarray(rays_o, f32), np.asarray(rays_d, f32)
    step = f32(step)
    b = aabbs[-1]
    ri, ts_out, te_out = [], [], []
    for r in range(len(rays_o)):
        o, d = rays_o[r], rays_d[r]
        t0, t1, miss = f32(-np.inf), f32(np.inf), False
        for k in range(3):
            if d[k] != 0:
                inv = f32(1) / d[k]
                ta, tb = (b[k] - o[k]) * inv, (b[3 + k] - o[k]) * inv
                t0, t1 = max(t0, min(ta, tb)), min(t1, max(ta, tb))
            elif o[k] < b[k] or o[k] > b[3 + k]:
                miss = True
        nr = f32(near if near_planes is None else near_planes[r])
        # lattice nr + k * step anchored at the ray's own (jittered) near plane: first point at or
        # after the box entry (csrc/occgrid.cu; nerfacc 0.5.3's traversal keeps its running t on
        # that lattice too — parity unpinned, see the module header)
        t_begin = _fma(f32(np.ceil(max(t0 - nr, f32(0)) / step)), step, nr)
        t_limit = min(f32(far), t1)
        if miss or not t_limit > t_begin:
            continue
        n_cand = int(min(np.ceil((t_limit - t_begin) / step), 1.0e7))
        k = np.arange(n_cand, dtype=f32)
        ts = _fma(k, step, t_begin)
        tm = ts + f32(0.5) * step
        p = np.stack([_fma(d[0], tm, o[0]), _fma(d[1], tm, o[1]), _fma(d[2], tm, o[2])], -1)
        keep = (tm < t_limit) & occupied(binaries, aabbs, p)
        ri.append(np.full(int(keep.sum()), r, np.int64))
        ts_out.append(ts[keep])
        te_out.append(ts[keep] + step)
    if not ri:
        return np.zeros(0, np.int64), np.zeros(0, f32), np.zeros(0, f32)
    return np.concatenate(ri), np.concatenate(ts_out), np.concatenate(te_out)


def visibility(sigmas, t_starts, t_ends, ray_indices, early_stop_eps=1e-4, alpha_thre=0.0):
    import torch
    from .compositing import exclusive_sum_packed
    sd = torch.as_tensor(sigmas) * (torch.as_tensor(t_ends) - torch.as_tensor(t_starts))
    ri = torch.as_tensor(ray_indices)
    trans = torch.exp(-exclusive_sum_packed(sd, ri, int(ri.max()) + 1 if len(ri) else 0))
    keep = trans >= early_stop_eps
    if alpha_thre > 0:
        keep &= (1 - torch.exp(-sd)) >= alpha_thre
    return keep.numpy()


def update(occs, occ, cell_ids=None, decay=0.95):
    """-> new occs (one level)"""
    occs = np.asarray(occs, f32).copy()
    occ = np.asarray(occ, f32)
    ids = np.arange(len(occs)) if cell_ids is None else np.asarray(cell_ids)
    best = np.full(len(occs), -np.inf, f32)
    np.maximum.at(best, ids, occ)
    touched = np.zeros(len(occs), bool)
    touched[ids] = True
    occs[touched] = np.maximum(occs[touched] * f32(decay), best[touched])
    return occs


def binarize(occs, occ_thre=1e-2):
    thre = min(float(np.asarray(occs, f32)[occs >= 0].mean(dtype=f32)), occ_thre)
    return occs > f32(thre), thre
