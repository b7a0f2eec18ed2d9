"""Occupancy-grid sampler, packed compositing and grid update (the
reference's real sampling path through nerfacc's OccGridEstimator) vs oracle/occgrid.py and
oracle/compositing.render_packed."""
import numpy as np
import pytest
import torch

from oracle import occgrid as oocc, compositing as ocomp, mlp as omlp

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def dev():
    from fsnerf_b200 import ops
    ops.require_device(0)
    return torch.device("cuda:0")


def _rays(n, seed, inside_frac=0.2):
    rng = np.random.default_rng(seed)
    o = rng.standard_normal((n, 3)).astype(f32)
    o = 4.0 * o / np.linalg.norm(o, axis=-1, keepdims=True)
    o[: int(n * inside_frac)] *= 0.2  # some cameras inside the box
    tgt = rng.uniform(-1, 1, (n, 3)).astype(f32)
    d = tgt - o
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d[-1] = [0, 0, 1]  # axis-aligned ray (zero components in the slab test)
    o[-1] = [0.3, -0.2, -4]
    d[-2] = [0, 1, 0]  # misses the box entirely
    o[-2] = [5, -4, 0]
    return o, d


@pytest.mark.parametrize("levels,res,step,strat", [(1, 16, 0.05, False), (2, 8, 0.031, True), (1, 128, 5e-3, True)])
def test_march_matches_oracle(dev, levels, res, step, strat):
    from fsnerf_b200 import ops
    n = 150 if res < 128 else 40
    o, d = _rays(n, seed=levels * 7 + res)
    rng = np.random.default_rng(3)
    binaries = rng.random((levels, res, res, res)) < (0.3 if res < 128 else 0.02)
    aabbs = oocc.level_aabbs([-1.5] * 3 + [1.5] * 3, levels)
    near_planes = (rng.random(n).astype(f32) * f32(step)) if strat else None
    ri_ref, ts_ref, te_ref = oocc.march(o, d, binaries, aabbs, step, near_planes=near_planes)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    ri, ts, te, offsets = ops.occgrid_march(cu(o), cu(d), cu(binaries), cu(aabbs), step,
                                            near_planes=None if near_planes is None else cu(near_planes))
    assert len(ri_ref) > 100
    np.testing.assert_array_equal(ri.cpu().numpy(), ri_ref)          # sample set + ray bookkeeping: bit-exact
    np.testing.assert_array_equal(ts.cpu().numpy(), ts_ref)
    np.testing.assert_array_equal(te.cpu().numpy(), te_ref)
    counts = np.bincount(ri_ref, minlength=n)
    np.testing.assert_array_equal(offsets.cpu().numpy(), np.concatenate([[0], np.cumsum(counts)]))
    assert counts[-2] == 0  # the ray that misses the box
    # empty grid / no rays
    z = ops.occgrid_march(cu(o), cu(d), cu(np.zeros_like(binaries)), cu(aabbs), step)
    assert z[0].numel() == 0 and int(z[3][-1]) == 0
    assert ops.occgrid_march(cu(o[:0]), cu(d[:0]), cu(binaries), cu(aabbs), step)[0].numel() == 0


def test_packed_compositing_matches_oracle(dev):
    from fsnerf_b200 import ops
    g = torch.Generator().manual_seed(0)
    counts = torch.tensor([5, 0, 33, 64, 1, 0, 97, 250, 32, 0])
    R, N = len(counts), int(counts.sum())
    ri = torch.repeat_interleave(torch.arange(R), counts)
    ts = torch.cat([torch.sort(2 + 4 * torch.rand(int(c), generator=g)).values for c in counts])
    te = ts + 0.02
    rgbs = torch.rand(N, 3, generator=g).requires_grad_(True)
    sig = (torch.randn(N, generator=g) * 6).requires_grad_(True)
    bk = torch.tensor([1.0, 0.5, 0.25], requires_grad=True)
    rgb, op, dp, ex = ocomp.render_packed(ts, te, ri, R, rgbs, sig, bk)
    raw = torch.cat([rgbs, sig[:, None]], -1).detach().to(dev)
    offsets = ops.offsets_from_ray_indices(ri.to(dev), R)
    out = ops.composite_packed_forward(raw, ts.to(dev), te.to(dev), offsets, bkgd=bk.detach().to(dev))
    for name, a, b in zip(("rgb", "opacity", "depth", "weights", "trans", "alphas"), out,
                          (rgb, op, dp, ex["weights"], ex["trans"], ex["alphas"])):
        assert (a.cpu().reshape(-1) - b.detach().reshape(-1)).abs().max().item() <= 2e-5 * max(
            1.0, float(b.detach().abs().max())), name
    d_rgb, d_op, d_dp = torch.randn(R, 3, generator=g), torch.randn(R, 1, generator=g), torch.randn(R, 1, generator=g)
    d_dp[op.detach() < 1e-3] = 0
    d_w = torch.randn(N, generator=g)
    loss = (rgb * d_rgb).sum() + (op * d_op).sum() + (dp * d_dp).sum() + (ex["weights"] * d_w).sum()
    g_rgbs, g_sig, g_bk = torch.autograd.grad(loss, [rgbs, sig, bk])
    d_raw, d_bk = ops.composite_packed_backward(raw, ts.to(dev), te.to(dev), offsets, out[4], d_rgb.to(dev),
                                                d_op.to(dev), d_dp.to(dev), d_w.to(dev), bkgd=bk.detach().to(dev),
                                                want_d_bkgd=True)
    ref = torch.cat([g_rgbs, g_sig[:, None]], -1)
    assert (d_raw.cpu() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    assert (d_bk.cpu() - g_bk).abs().max().item() <= 1e-4 * max(1.0, g_bk.abs().max().item())
    # rays without samples: background colour, zero opacity / depth
    np.testing.assert_allclose(out[0][1].cpu().numpy(), bk.detach().numpy())
    assert float(out[1][1]) == 0.0 and float(out[2][5]) == 0.0


def test_grid_update_matches_oracle(dev):
    from fsnerf_b200 import ops
    rng = np.random.default_rng(0)
    n_cells = 16 ** 3
    occs = rng.random(n_cells).astype(f32) * 0.02
    occs_d = torch.from_numpy(occs).to(dev)
    occ = (rng.standard_normal(n_cells) * 0.02).astype(f32)  # warm-up: every cell once
    ops.occgrid_update(occs_d, torch.from_numpy(occ).to(dev))
    ref = oocc.update(occs, occ)
    np.testing.assert_array_equal(occs_d.cpu().numpy(), ref)
    ids = rng.integers(0, n_cells, 3000)                       # later: random subset with collisions
    occ2 = (rng.standard_normal(3000) * 0.02).astype(f32)
    ops.occgrid_update(occs_d, torch.from_numpy(occ2).to(dev), cell_ids=torch.from_numpy(ids).to(dev))
    ref2 = oocc.update(ref, occ2, ids)
    np.testing.assert_array_equal(occs_d.cpu().numpy(), ref2)
    bin_ref, thre = oocc.binarize(ref2, 1e-2)
    binaries = torch.zeros(n_cells, dtype=torch.bool, device=dev)
    ops.occgrid_binarize(occs_d, thre, binaries.view(torch.uint8))
    np.testing.assert_array_equal(binaries.cpu().numpy(), bin_ref)


def test_occgrid_estimator_render_rays_and_train_loop(dev):
    """The reference's own call pattern (src/run-nerf.py:232-295) on the drop-in estimator: first
    step with an empty grid renders the background, update_every_n_steps fills the grid from the
    model's density, later steps march / filter / composite packed samples and backpropagate."""
    from fsnerf_b200.core.models import NeRF
    from fsnerf_b200.render.rendering import OccGridEstimator, render_rays
    from oracle import render as orender  # noqa: F401
    kw = {"pos_fn": {"n_freqs": 10, "log_space": True}, "dir_fn": {"n_freqs": 4, "log_space": True}}
    torch.manual_seed(42)
    model = NeRF(3, 3, 8, 256, [4], **kw).to(dev)
    with torch.no_grad():
        model.sigma.bias += 3.0  # dense enough for the transmittance filter to matter
    est = OccGridEstimator([-1.5] * 3 + [1.5] * 3, resolution=32, levels=1).to(dev)
    o, d = _rays(300, seed=1, inside_frac=0.0)
    ro, rd = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    gt = torch.rand(300, 3, device=dev)
    step_size = 0.02
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    model.train(); est.train()
    (rgb, op, dp, extras), ri, tv = render_rays(ro, rd, est, model, train=True, white_bkgd=True,
                                                render_step_size=step_size, device=dev)
    assert ri.numel() == 0 and len(extras["sigmas"]) == 0  # run-nerf.py:262 guards on this
    assert torch.equal(rgb.detach(), torch.ones(300, 3, device=dev)) and float(op.detach().abs().max()) == 0
    torch.nn.functional.mse_loss(rgb, gt).backward()  # reaches only the background colour
    est.update_every_n_steps(step=0, occ_eval_fn=lambda x: model(x) * step_size, occ_thre=1e-2)
    n_occ = int(est.binaries.sum())
    assert 0 < n_occ <= 32 ** 3
    # grid == oracle update of the same densities (warm-up evaluates every cell once; rand differs,
    # so compare through the invariant: binaries = occs > min(mean, thre))
    occs = est.occs.cpu().numpy()
    np.testing.assert_array_equal(est.binaries.flatten().cpu().numpy(), oocc.binarize(occs, 1e-2)[0])
    est.set_uniforms(torch.full((300,), 0.25, device=dev))
    (rgb, op, dp, extras), ri, tv = render_rays(ro, rd, est, model, train=True, white_bkgd=True,
                                                render_step_size=step_size, device=dev)
    assert ri.numel() > 1000 and bool((ri[1:] >= ri[:-1]).all())
    # the surviving samples = oracle march + oracle visibility filter on the kernel's own sigmas
    aabbs = oocc.level_aabbs([-1.5] * 3 + [1.5] * 3, 1)
    nearp = np.full(300, 0.25 * step_size, f32)
    ri0, ts0, te0 = oocc.march(o, d, est.binaries.cpu().numpy(), aabbs, step_size, near_planes=nearp)
    with torch.no_grad():
        x0 = ro[torch.from_numpy(ri0).to(dev)] + rd[torch.from_numpy(ri0).to(dev)] * torch.from_numpy(
            (ts0 + te0)).to(dev)[:, None] / 2.0
        sig0 = model(x0).squeeze(-1).cpu().numpy()
    keep_ref = oocc.visibility(sig0, ts0, te0, ri0, 1e-4, 0.0)
    assert 0 < keep_ref.sum() < len(keep_ref)
    # the kernel's kept set is a subsequence of the marched list; it may differ from the oracle's only
    # for samples whose transmittance sits within rounding of the 1e-4 threshold (scan order differs)
    key0 = ri0.astype(np.float64) * 1e3 + ((ts0 + te0) / f32(2.0)).astype(np.float64)
    keyk = ri.cpu().numpy().astype(np.float64) * 1e3 + tv.cpu().numpy().astype(np.float64)
    keep = np.isin(key0, keyk)
    assert keep.sum() == len(keyk) and (keep != keep_ref).sum() <= 3, (keep != keep_ref).sum()
    # composite of the kept samples vs the oracle's packed renderer (fp32 reference MLP on the same points)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    xk = torch.from_numpy(o)[ri0[keep]] + torch.from_numpy(d)[ri0[keep]] * torch.from_numpy((ts0 + te0)[keep])[:, None] / 2
    raw_ref = omlp.nerf_forward(sd, xk, torch.from_numpy(d)[ri0[keep]])
    rgb_ref, op_ref, dp_ref, _ = ocomp.render_packed(torch.from_numpy(ts0[keep]), torch.from_numpy(te0[keep]),
                                                     torch.from_numpy(ri0[keep]), 300, raw_ref[:, :3], raw_ref[:, 3],
                                                     torch.ones(3))
    assert (rgb.detach().cpu() - rgb_ref).abs().max().item() < 1e-3
    assert (op.detach().cpu() - op_ref).abs().max().item() < 1e-3
    # backward reaches every parameter; a few optimisation steps reduce the loss
    losses = []
    for k in range(1, 9):
        opt.zero_grad()
        (rgb, *_), ri, tv = render_rays(ro, rd, est, model, train=True, white_bkgd=True, render_step_size=step_size,
                                        device=dev)
        loss = torch.nn.functional.mse_loss(rgb, gt)
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
        opt.step()
        est.update_every_n_steps(step=k, occ_eval_fn=lambda x: model(x) * step_size, occ_thre=1e-2)
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    est.eval()
    with pytest.raises(RuntimeError):
        est.update_every_n_steps(step=16, occ_eval_fn=lambda x: model(x) * step_size)


def test_sinerf_through_packed_path(dev):
    """--model sinerf (src/run-nerf.py:81-88) on the drop-in estimator: the module mirror's own
    autograd composes with the packed compositing kernels; gradients match a pure-torch packed
    renderer on the same samples and a few steps reduce the loss."""
    from fsnerf_b200.core.models import SiNeRF
    from fsnerf_b200.render.rendering import OccGridEstimator, render_rays
    torch.manual_seed(42)
    model = SiNeRF(3, 3, 256, [30.] + [1.] * 7).to(dev)
    est = OccGridEstimator([-1.5] * 3 + [1.5] * 3, resolution=16, levels=1).to(dev)
    est.binaries[:] = True  # fully occupied: every ray marches through the box
    o, d = _rays(64, seed=2, inside_frac=0.0)
    ro, rd = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    gt = torch.rand(64, 3, device=dev)
    model.train(); est.train()
    (rgb, op, dp, ex), ri, tv = render_rays(ro, rd, est, model, train=False, white_bkgd=True, render_step_size=0.1,
                                            device=dev)
    assert ri.numel() > 500 and rgb.requires_grad
    loss = torch.nn.functional.mse_loss(rgb, gt)
    grads = torch.autograd.grad(loss, list(model.parameters()))
    # reference: the same samples rendered by the oracle's packed renderer in torch (CPU, fp32)
    cpu = SiNeRF(3, 3, 256, [30.] + [1.] * 7)
    cpu.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    ri_c, t_mid = ri.cpu(), tv.cpu()
    x = torch.from_numpy(o)[ri_c] + torch.from_numpy(d)[ri_c] * t_mid[:, None]
    raw = cpu(x, torch.from_numpy(d)[ri_c])
    rgb_ref, *_ = ocomp.render_packed(t_mid - 0.05, t_mid + 0.05, ri_c, 64, raw[:, :3], raw[:, 3], torch.ones(3))
    assert (rgb.detach().cpu() - rgb_ref.detach()).abs().max().item() < 2e-4
    g_ref = torch.autograd.grad(torch.nn.functional.mse_loss(rgb_ref, gt.cpu()), list(cpu.parameters()))
    for a, b in zip(grads, g_ref):
        assert ((a.cpu() - b).norm() / b.norm().clamp_min(1e-12)).item() < 2e-3
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        (rgb, *_), _, _ = render_rays(ro, rd, est, model, train=True, white_bkgd=True, render_step_size=0.1, device=dev)
        loss = torch.nn.functional.mse_loss(rgb, gt)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


def test_single_sample_assertion_fallback(dev):
    """src/render/rendering.py:88-103: with exactly ONE surviving sample nerfacc's shape assertion fires
    (`out[..., -1].squeeze(-1)` is 0-d) and render_rays answers with a constant background, opacity and
    extras None, depth zeros — reproduced on the packed path."""
    from fsnerf_b200.core.models import NeRF
    from fsnerf_b200.render.rendering import OccGridEstimator, render_rays
    kw = {"pos_fn": {"n_freqs": 10, "log_space": True}, "dir_fn": {"n_freqs": 4, "log_space": True}}
    model = NeRF(3, 3, 8, 256, [4], **kw).to(dev)
    est = OccGridEstimator([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], resolution=4, levels=1).to(dev)
    est.binaries.zero_()
    est.binaries[0, 2, 2, 0] = True  # one occupied cell: z in [-1, -0.5]
    # a ray along -z through that cell with a step as long as the cell: one interval survives
    o = torch.tensor([[0.25, 0.25, 3.0], [0.9, 0.9, 3.0]])
    d = torch.tensor([[0.0, 0.0, -1.0], [0.0, 0.0, -1.0]])
    est.eval()
    with torch.no_grad():
        ri, ts, te = est.sampling(o.to(dev), d.to(dev), render_step_size=0.5)
        assert ts.numel() == 1, ts
        (rgb, opac, depth, extras), ri2, tv = render_rays(o, d, est, model, train=False, white_bkgd=True,
                                                          render_step_size=0.5, device=dev)
    assert opac is None and extras is None
    assert torch.equal(rgb.cpu(), torch.ones(2, 3)) and torch.equal(depth.cpu(), torch.zeros(2, 1))
    assert ri2.numel() == 1 and tv.numel() == 1
