"""BASELINE.json full sizes (C2: 4096 rays x 64+128 samples; C4: 800x800 frame chunks) checked
through size-independent properties of the domain, where the CPU oracle would take minutes:
sortedness and coverage of the samples, conservation and linearity of compositing, a checksum
of checksums for the ray table, chunk-size invariance of rendering, determinism."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def dev():
    from fsnerf_b200 import ops
    ops.require_device(0)
    return torch.device("cuda:0")


def test_c2_sampling_properties(dev):
    from fsnerf_b200 import ops
    R, Sc, Sf, near, far = 4096, 64, 128, 2.0, 6.0
    g = torch.Generator(device=dev).manual_seed(0)
    u = torch.rand(R, Sc, device=dev, generator=g)
    ts, te = ops.sample_stratified(R, Sc, near, far, u)
    # bins: contiguous, ordered, inside [near, far]; sample k lies in stratum k
    assert torch.equal(te[:, :-1], ts[:, 1:]) and bool((te >= ts).all())
    assert float(ts.min()) >= near and float(te.max()) == far
    # canonical NeRF strata: S points on linspace(near, far), stratum k
    # = [mid_{k-1}, mid_k] with the two end strata half as wide
    z = torch.linspace(near, far, Sc, device=dev)
    mids = 0.5 * (z[1:] + z[:-1])
    lower, upper = torch.cat([z[:1], mids]), torch.cat([mids, z[-1:]])
    assert bool((ts >= lower - 1e-6).all()) and bool((ts <= upper + 1e-6).all())
    w = torch.rand(R, Sc, device=dev, generator=g) ** 8
    w[:7] = 0  # rays with no mass fall back to uniform resampling
    up = torch.rand(R, Sf, device=dev, generator=g)
    ts_f, te_f, smp, inds, perm = ops.sample_pdf(ts, w, Sf, far, up)
    S = Sc + Sf
    assert ts_f.shape == (R, S) and torch.equal(te_f[:, :-1], ts_f[:, 1:])
    assert bool((ts_f[:, 1:] >= ts_f[:, :-1]).all())                     # merged samples sorted
    assert float(ts_f.min()) >= near and float(te_f.max()) == far
    assert bool((inds >= 0).all()) and bool((inds <= Sc - 2).all())       # searchsorted bins
    # perm is a permutation of the merged index range per ray; the merged set = coarse U fine
    assert torch.equal(torch.sort(perm.long(), -1).values, torch.arange(S, device=dev).expand(R, S))
    merged = torch.sort(torch.cat([ts, smp], -1), -1).values
    assert torch.equal(merged, ts_f)
    # importance: fine samples concentrate where the weights are (mass-weighted mean matches)
    mid = 0.5 * (ts[:, 1:] + ts[:, :-1])
    wm = w[:, 1:-1] + 1e-5
    centre = 0.5 * (mid[:, :-1] + mid[:, 1:])  # pdf = piecewise uniform over [mid_j, mid_j+1], mass w_j+1
    mean_pdf = (wm * centre).sum(-1) / wm.sum(-1)
    err = (smp.mean(-1) - mean_pdf)[7:]  # 128 draws per ray: sampling noise ~0.1, no bias over 4089 rays
    assert float(err.abs().mean()) < 0.15 and abs(float(err.mean())) < 0.01
    # idempotence / determinism
    again = ops.sample_pdf(ts, w, Sf, far, up)
    assert all(torch.equal(a, b) for a, b in zip((ts_f, te_f, smp, inds, perm), again))


def test_c2_compositing_properties(dev):
    from fsnerf_b200 import ops
    R, S = 4096, 192
    g = torch.Generator(device=dev).manual_seed(1)
    e = torch.sort(2 + 4 * torch.rand(R, S + 1, device=dev, generator=g), -1).values
    ts, te = e[:, :-1].contiguous(), e[:, 1:].contiguous()
    raw = torch.rand(R, S, 4, device=dev, generator=g)
    raw[..., 3] = torch.rand(R, S, device=dev, generator=g) * 8
    bk = torch.tensor([1.0, 0.5, 0.25], device=dev)
    rgb, op, dp, w, al, tr = ops.composite_forward(raw, ts, te, bkgd=bk, extras=True)
    # conservation: weights are a sub-probability, opacity = sum w = 1 - final transmittance
    assert bool((w >= 0).all()) and float(op.max()) <= 1 + 1e-5
    np.testing.assert_allclose(op[:, 0].cpu().numpy(), w.sum(-1).cpu().numpy(), atol=2e-5)
    t_end = tr[:, -1] * (1 - al[:, -1])
    np.testing.assert_allclose((1 - op[:, 0]).cpu().numpy(), t_end.cpu().numpy(), atol=2e-5)
    assert bool((tr[:, 1:] <= tr[:, :-1] + 1e-7).all())                   # transmittance decreases
    assert float(dp.min()) >= 2.0 - 1e-4 and float(dp.max()) <= 6.0 + 1e-4  # depth inside [near, far]
    # linearity in colour: C(a c1 + b c2) + bkgd term = a C(c1) + b C(c2) (same sigma)
    raw2 = raw.clone()
    raw2[..., :3] = torch.rand(R, S, 3, device=dev, generator=g)
    mix = raw.clone()
    mix[..., :3] = 0.3 * raw[..., :3] + 0.7 * raw2[..., :3]
    rgb2 = ops.composite_forward(raw2, ts, te, bkgd=bk)[0]
    rgbm = ops.composite_forward(mix, ts, te, bkgd=bk)[0]
    np.testing.assert_allclose(rgbm.cpu().numpy(), (0.3 * rgb + 0.7 * rgb2).cpu().numpy(), atol=3e-6)
    # constant colour: rgb = c * opacity + bkgd * (1 - opacity)
    const = raw.clone()
    const[..., :3] = torch.tensor([0.2, 0.4, 0.6], device=dev)
    rgbc = ops.composite_forward(const, ts, te, bkgd=bk)[0]
    ref = torch.tensor([0.2, 0.4, 0.6], device=dev) * op + bk * (1 - op)
    np.testing.assert_allclose(rgbc.cpu().numpy(), ref.cpu().numpy(), atol=3e-6)
    # backward: d(sum rgb)/d(rgb_s) = w_s (the adjoint of the forward weights), bit-for-bit determinism
    ones = torch.ones(R, 3, device=dev)
    d_raw, _ = ops.composite_backward(raw, ts, te, ones, bkgd=bk)
    np.testing.assert_allclose(d_raw[..., 0].cpu().numpy(), w.cpu().numpy(), atol=2e-6)
    assert torch.equal(d_raw, ops.composite_backward(raw, ts, te, ones, bkgd=bk)[0])
    # empty / single-sample edge cases
    assert ops.composite_forward(raw[:0], ts[:0], te[:0])[0].shape == (0, 3)
    r1 = ops.composite_forward(raw[:5, :1].contiguous(), ts[:5, :1].contiguous(), te[:5, :1].contiguous())
    a1 = 1 - torch.exp(-raw[:5, 0, 3] * (te[:5, 0] - ts[:5, 0]))
    np.testing.assert_allclose(r1[1][:, 0].cpu().numpy(), a1.cpu().numpy(), atol=1e-6)


def test_c4_ray_table_checksums_and_chunk_invariance(dev):
    """800x800 frame: the ray table's checksum-of-checksums is independent of how the pixel range
    is chunked / partitioned across ranks (pixel bookkeeping bit-exact), and a rendered frame is
    bit-identical for any chunk size (ragged last chunk included)."""
    from fsnerf_b200 import ops, synthetic as syn, parallel
    from fsnerf_b200.engine import HotPath
    H = W = 800
    focal = syn.focal_from_fov(W)
    pose = torch.from_numpy(syn.orbit_poses(4)[1]).to(dev)[None].contiguous()
    o_all, d_all, _ = ops.gen_rays(pose, H, W, focal, first_id=0, n_rays=H * W)
    np.testing.assert_allclose(d_all.norm(dim=-1).cpu().numpy(), 1.0, atol=2e-7)
    assert torch.equal(o_all, pose[0, :3, 3].expand(H * W, 3))
    whole = d_all.view(torch.int32).sum(dtype=torch.int64)
    for world in (1, 2, 4, 8):
        parts = []
        for rank in range(world):
            for (f, a, b) in parallel.pixel_partition(1, H, W, rank, world):
                parts.append(ops.gen_rays(pose, H, W, focal, first_id=a, n_rays=b - a)[1])
        assert torch.equal(torch.cat(parts), d_all)
        assert sum(p.view(torch.int32).sum(dtype=torch.int64) for p in parts) == whole
    ids = torch.randperm(H * W, generator=torch.Generator().manual_seed(0))[:5000].to(dev)
    assert torch.equal(ops.gen_rays(pose, H, W, focal, pixel_ids=ids)[1], d_all[ids])
    # chunk invariance of the render (deterministic eval sampling): 3 chunkings of 20 000 rays
    hp = HotPath(n_coarse=64, n_fine=128, device=dev)
    n = 20000
    ref = hp.render(o_all[:n], d_all[:n])
    ref = [t.clone() for t in ref]
    for chunk in (4096, 7777):
        outs = [hp.render(o_all[i:i + chunk][: max(0, n - i)].contiguous(), d_all[i:i + chunk][: max(0, n - i)].contiguous())
                for i in range(0, n, chunk)]
        for k in range(3):
            assert torch.equal(torch.cat([o[k] for o in outs]), ref[k]), (chunk, k)
    assert float(ref[1].min()) >= 0 and float(ref[1].max()) <= 1 + 1e-5


def test_c2_train_step_determinism_and_data_parallel_sum(dev):
    """Full C2 batch (4096 rays, 64+128): the gradient of the global batch equals the sum of the
    gradients of its two halves seeded with the GLOBAL 1/(3R) scale (what the NCCL all-reduce
    sums), and two runs of the same step agree to fp32 summation order."""
    from fsnerf_b200 import synthetic as syn, ops
    from fsnerf_b200.engine import HotPath
    R = 4096
    poses, imgs, focal = syn.make_views(4, 64, 64, seed=42)
    pd, im = torch.from_numpy(poses).to(dev), torch.from_numpy(imgs).to(dev)
    ids = torch.randperm(4 * 64 * 64, generator=torch.Generator().manual_seed(3))[:R].to(dev)
    o, d, gt = ops.gen_rays(pd, 64, 64, focal, pixel_ids=ids, images=im)
    g = torch.Generator(device=dev).manual_seed(5)
    us, up = torch.rand(R, 64, device=dev, generator=g), torch.rand(R, 128, device=dev, generator=g)
    hp = HotPath(n_coarse=64, n_fine=128, device=dev, world_size=1)
    ls = hp.train_step(o, d, gt, us, up, apply_update=False).clone()
    full = hp.grads.clone()
    hp.train_step(o, d, gt, us, up, apply_update=False)
    rel = ((hp.grads - full).norm() / full.norm()).item()
    assert rel < 1e-5, rel
    acc = torch.zeros_like(full)
    ls_sum = torch.zeros_like(ls)
    for a, b in ((0, R // 2), (R // 2, R)):
        ls_sum += hp.train_step(o[a:b], d[a:b], gt[a:b], us[a:b], up[a:b], global_rays=R, apply_update=False)
        acc += hp.grads
    # fp32 summation order only: a weight-stationary wgrad CTA accumulates >100 k samples in tensor
    # memory, and the halves group them differently (measured ~1e-5)
    rel_halves = ((acc - full).norm() / full.norm()).item()
    assert rel_halves < 5e-5, rel_halves
    np.testing.assert_allclose(ls_sum.cpu().numpy(), ls.cpu().numpy(), rtol=1e-5)
    assert torch.isfinite(full).all() and float(full.abs().max()) > 0
