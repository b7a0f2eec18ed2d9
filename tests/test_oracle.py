"""CPU tests: the oracle against the reference-generated golden fixtures and
closed-form known answers."""
import os
import pytest
import numpy as np
import torch

from oracle import rays, encoding, mlp, sampling, compositing, render

f32 = np.float32


# ---------------------------------------------------------------- rays
def test_get_rays_matches_reference(golden):
    g = golden("reference_rays.npz")
    H, W, f = g["hwf2"]
    o, d = rays.get_rays(g["pose2"], (int(H), int(W), float(f)))
    assert o.shape == (12, 20, 3)
    np.testing.assert_array_equal(o, g["ro2"])
    np.testing.assert_allclose(d, g["rd2"], rtol=0, atol=1.2e-7)
    no, nd = rays.to_ndc(o.reshape(-1, 3), d.reshape(-1, 3), (int(H), int(W), float(f)), 1.0)
    np.testing.assert_allclose(no, g["ndc_o2"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(nd, g["ndc_d2"], rtol=2e-6, atol=2e-6)


def test_get_rays_pinned_values(golden):
    g = golden("reference_rays.npz")
    focal = float(g["focal"])
    o, d = rays.get_rays(g["pose"], (100, 100, focal))
    np.testing.assert_allclose(d[50, 50], [0, 0, -1], atol=1e-7)
    np.testing.assert_allclose(d[0, 0], [-0.3208, 0.3208, -0.8912], atol=1e-4)
    np.testing.assert_allclose(d[0, 0], g["rd_00"], atol=1.2e-7)
    np.testing.assert_allclose(d[99, 99], g["rd_last"], atol=1.2e-7)
    np.testing.assert_allclose(d[7], g["rd_row7"], atol=1.2e-7)
    np.testing.assert_array_equal(o[0, 0], g["ro_00"])
    n = np.linalg.norm(d.astype(np.float64), axis=-1)
    assert abs(n.min() - 1) < 2e-7 and abs(n.max() - 1) < 2e-7
    no, nd = rays.to_ndc(o.reshape(-1, 3), d.reshape(-1, 3), (100, 100, focal), 1.0)
    np.testing.assert_allclose(no[0], [-5, 5, -1], atol=1e-5)
    np.testing.assert_allclose(nd[0], [4, -4, 2], atol=1e-5)
    np.testing.assert_allclose(no[700:720], g["ndc_o_row"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(nd[700:720], g["ndc_d_row"], rtol=2e-6, atol=2e-6)


def test_chunks(golden):
    g = golden("reference_rays.npz")
    ch = rays.get_chunks(10000, 4096)
    assert [b - a for a, b in ch] == list(g["chunks"]) == [4096, 4096, 1808]
    assert rays.get_chunks(0, 16) == []


def test_rays_from_pixel_ids(golden):
    g = golden("reference_rays.npz")
    H, W, f = int(g["hwf2"][0]), int(g["hwf2"][1]), float(g["hwf2"][2])
    poses = np.stack([g["pose2"], g["pose"]])
    ids = np.array([0, 5, H * W - 1, H * W, H * W + 37])
    o, d = rays.rays_from_pixel_ids(poses, (H, W, f), ids)
    np.testing.assert_array_equal(d[1], rays.get_rays(poses[0], (H, W, f))[1].reshape(-1, 3)[5])
    np.testing.assert_array_equal(d[4], rays.get_rays(poses[1], (H, W, f))[1].reshape(-1, 3)[37])


# ------------------------------------------------------------ encoding
def test_positional_encoding_matches_reference(golden):
    g = golden("reference_mlp.npz")
    x, d = torch.from_numpy(g["x"]), torch.from_numpy(g["d"])
    np.testing.assert_array_equal(encoding.positional_encoding(x, 10).numpy(), g["pe_pos"])
    np.testing.assert_array_equal(encoding.positional_encoding(d, 4).numpy(), g["pe_dir"])
    np.testing.assert_array_equal(encoding.positional_encoding(d, 4, False).numpy(), g["pe_lin"])
    pin = encoding.positional_encoding(torch.tensor([[0.1, -0.2, 0.3]]), 10).numpy()
    np.testing.assert_array_equal(pin, g["pe_pin"])
    np.testing.assert_allclose(pin[0, :12], [0.1, -0.2, 0.3, 0.0998, -0.1987, 0.2955, 0.9950,
                                             0.9801, 0.9553, 0.1987, -0.3894, 0.5646], atol=1e-4)
    assert pin.shape[1] == 63 and g["pe_dir"].shape[1] == 27


def test_freq_mask():
    assert encoding.freq_mask(63, 100, 100).tolist() == [1.0] * 63
    assert encoding.freq_mask(63, 5, 0).tolist() == [1.0] * 63
    m0 = encoding.freq_mask(63, 0, 100)
    assert m0[:3].tolist() == [1, 1, 1] and m0[3:].sum() == 0
    m = encoding.freq_mask(63, 50, 100)  # ptr = 21*0.5+1 = 11.5
    assert m[:33].tolist() == [1.0] * 33 and m[33:36].tolist() == [0.5] * 3 and m[36:].sum() == 0
    m = encoding.freq_mask(27, 99, 100)  # ptr = min(9*.99+1, 9) = 9
    assert m.tolist() == [1.0] * 27
    # monotone in step
    prev = np.zeros(63)
    for s in range(0, 101, 7):
        cur = encoding.freq_mask(63, s, 100)
        assert (cur >= prev - 1e-7).all()
        prev = cur


# ----------------------------------------------------------------- mlp
def test_init_matches_reference_seed42(golden):
    g = golden("reference_mlp.npz")
    sd = mlp.init_state_dict()
    assert sum(v.numel() for v in sd.values()) == int(g["n_params"]) == 595844
    assert len(sd) == 24
    for i, name in enumerate(g["names"]):
        v = sd[str(name)]
        assert str(tuple(v.shape)) == str(g["shapes"][i])
        assert abs(v.double().sum().item() - g["w_sum"][i]) < 1e-9
        assert abs(v.double().abs().sum().item() - g["w_abs"][i]) < 1e-9


def test_mlp_forward_backward_matches_reference(golden):
    g = golden("reference_mlp.npz")
    sd = {k: v.requires_grad_(True) for k, v in mlp.init_state_dict().items()}
    x, d = torch.from_numpy(g["x"]), torch.from_numpy(g["d"])
    out = mlp.nerf_forward(sd, x, d)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=0, atol=1e-6)
    sig = mlp.nerf_forward(sd, x)
    np.testing.assert_allclose(sig.detach().numpy(), g["sigma_only"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out[:, 3:].detach().numpy(), g["sigma_only"], atol=1e-6)
    ones = mlp.nerf_forward(sd, x, d, mask_pos=torch.ones(63), mask_dir=torch.ones(27))
    np.testing.assert_array_equal(ones.detach().numpy(), out.detach().numpy())
    loss = (out * torch.from_numpy(g["cvec"])).sum()
    names = [str(n) for n in g["gnames"]]
    grads = torch.autograd.grad(loss, [sd[n] for n in names])
    for i, n in enumerate(names):
        assert abs(grads[i].double().norm().item() - g["g_norm"][i]) <= 1e-5 * max(1, g["g_norm"][i])
    np.testing.assert_allclose(grads[names.index("layers.3.weight")][:8, :8].numpy(),
                               g["g_layers3"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(grads[names.index("rgb.weight")].numpy(), g["g_rgb_w"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(grads[names.index("sigma.weight")].numpy(), g["g_sigma_w"], rtol=1e-4, atol=1e-5)


# ---------------------------------------------------------- compositing
def test_render_rays_callsite_matches_reference(golden):
    """Reference render_rays (its own code, stub estimator) vs oracle
    query_mlp + composite_dense: pins midpoint positions, dir gather, rgb/sigma
    split, background and t_vals (src/render/rendering.py:58-107)."""
    g = golden("reference_render.npz")
    sd = mlp.init_state_dict()
    sd["sigma.weight"] = sd["sigma.weight"] * float(g["sigma_w_scale"])
    sd["sigma.bias"] = sd["sigma.bias"] + float(g["sigma_b_add"])
    ro, rd = torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])
    ts, te = torch.from_numpy(g["t_starts"]), torch.from_numpy(g["t_ends"])
    raw = render.query_mlp(sd, ro, rd, ts, te)
    for tag, bk in (("b", None), ("w", torch.ones(3))):
        rgb, op, dp, w, _, _ = compositing.composite_dense(raw, ts, te, bk)
        np.testing.assert_allclose(rgb.numpy(), g[f"rgb_{tag}"], atol=2e-6)
        np.testing.assert_allclose(op.numpy(), g[f"opacity_{tag}"], atol=2e-6)
        np.testing.assert_allclose(dp.numpy(), g[f"depth_{tag}"], rtol=5e-6, atol=1e-5)
        np.testing.assert_allclose(w.reshape(-1).numpy(), g[f"weights_{tag}"], atol=2e-6)
        np.testing.assert_allclose(raw[..., 3].reshape(-1).numpy(), g[f"sigmas_{tag}"], atol=2e-5)
    R, S = ts.shape
    np.testing.assert_array_equal(g["ray_indices"], np.repeat(np.arange(R), S))
    np.testing.assert_allclose(g["t_vals"], ((ts + te) / 2).reshape(-1).numpy(), atol=0)
    assert bool(g["seen_stratified"]) is True and float(g["seen_far"]) == 1e10
    assert np.abs(g["opacity_b"]).min() > 0.01  # not trivially transparent (sigma is raw: may be < 0)


def test_composite_known_answers():
    # one ray, constant sigma s over [0,L] in S equal steps: opacity = 1-exp(-sL)
    S, s, L = 32, 0.7, 3.0
    e = torch.linspace(0, L, S + 1)[None]
    ts, te = e[:, :-1], e[:, 1:]
    raw = torch.zeros(1, S, 4)
    raw[..., 0] = 0.25
    raw[..., 3] = s
    rgb, op, dp, w, a, T = compositing.composite_dense(raw, ts, te, torch.ones(3))
    acc = 1 - np.exp(-s * L)
    assert abs(op.item() - acc) < 1e-6
    np.testing.assert_allclose(rgb[0].numpy(), [0.25 * acc + (1 - acc), 1 - acc, 1 - acc], atol=1e-6)
    assert abs(T[0, -1].item() - np.exp(-s * (L - L / S))) < 1e-6
    # packed == dense, including an empty ray and ragged counts
    g = torch.Generator().manual_seed(0)
    counts = [3, 0, 5, 1]
    ri = torch.repeat_interleave(torch.arange(4), torch.tensor(counts))
    N = len(ri)
    ts_p = torch.rand(N, generator=g)
    te_p = ts_p + 0.1
    rgbs, sig = torch.rand(N, 3, generator=g), torch.randn(N, generator=g) * 5
    c, o, d, ex = compositing.render_packed(ts_p, te_p, ri, 4, rgbs, sig, torch.ones(3))
    np.testing.assert_allclose(c[1].numpy(), [1, 1, 1])
    assert o[1].item() == 0 and d[1].item() == 0
    off = 0
    for r, n in enumerate(counts):
        if n == 0:
            continue
        raw = torch.cat([rgbs[off:off + n], sig[off:off + n, None]], -1)[None]
        rr, oo, dd, ww, _, _ = compositing.composite_dense(raw, ts_p[None, off:off + n],
                                                          te_p[None, off:off + n], torch.ones(3))
        np.testing.assert_allclose(rr[0].numpy(), c[r].numpy(), atol=1e-6)
        np.testing.assert_allclose(ww[0].numpy(), ex["weights"][off:off + n].numpy(), atol=1e-6)
        off += n


def test_composite_canonical_flags():
    """Appendix B3 switches reproduce textbook raw2outputs."""
    g = torch.Generator().manual_seed(3)
    R, S = 5, 12
    z = torch.sort(2 + 4 * torch.rand(R, S, generator=g), -1).values
    raw = torch.randn(R, S, 4, generator=g)
    raw[..., :3] = torch.sigmoid(raw[..., :3])
    dn = 1 + torch.rand(R, generator=g)
    te = torch.cat([z[:, 1:], z[:, -1:] + 1e10], -1)
    rgb, acc, depth, w, _, _ = compositing.composite_dense(
        raw, z, te, torch.ones(3), sigma_relu=True, delta_scale=dn,
        normalize_depth=False, product_trans=True)
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((R, 1), 1e10)], -1) * dn[:, None]
    alpha = 1 - torch.exp(-torch.relu(raw[..., 3]) * dists)
    T = torch.cumprod(torch.cat([torch.ones(R, 1), 1 - alpha + 1e-10], -1), -1)[:, :-1]
    w_ref = alpha * T
    np.testing.assert_allclose(w.numpy(), w_ref.numpy(), atol=1e-6)
    np.testing.assert_allclose(rgb.numpy(), ((w_ref[..., None] * raw[..., :3]).sum(1)
                                             + (1 - w_ref.sum(1, keepdim=True))).numpy(), atol=1e-6)


# ------------------------------------------------------------- sampling
def test_stratified():
    z = sampling.stratified(3, 64, 2.0, 6.0)
    assert z.dtype == f32 and z.shape == (3, 64)
    assert z[0, 0] == 2.0 and z[0, -1] == 6.0 and (np.diff(z, axis=-1) > 0).all()
    u = np.random.default_rng(0).random((3, 64), dtype=f32)
    zp = sampling.stratified(3, 64, 2.0, 6.0, u)
    mid = 0.5 * (z[:, 1:] + z[:, :-1])
    lower = np.concatenate([z[:, :1], mid], -1)
    upper = np.concatenate([mid, z[:, -1:]], -1)
    assert (zp >= lower).all() and (zp <= upper).all() and (np.diff(zp, axis=-1) >= 0).all()
    ts, te = sampling.intervals_from_points(zp, 6.0)
    np.testing.assert_array_equal(te[:, :-1], ts[:, 1:])
    assert (te[:, -1] == 6.0).all()


def test_sample_pdf_properties_and_canonical_agreement():
    rng = np.random.default_rng(1)
    R, Sc, Sf = 257, 64, 128
    u0 = rng.random((R, Sc), dtype=f32)
    z = sampling.stratified(R, Sc, 2.0, 6.0, u0)
    w = rng.random((R, Sc), dtype=f32) ** 8  # peaky
    w[0] = 0  # all-zero weights -> uniform pdf
    w[1, 1:-1] = 0
    w[1, 30] = 1.0  # a single spike
    u = rng.random((R, Sf), dtype=f32)
    sp = sampling.sample_pdf(z, w, Sf, 6.0, u)
    assert sp["inds"].dtype == np.int32 and sp["inds"].min() >= 1 and sp["inds"].max() <= Sc - 1
    cdf = sp["cdf"]
    assert (np.diff(cdf, axis=-1) > 0).all() and cdf[:, 0].max() == 0
    assert np.abs(cdf[:, -1] - 1).max() < 1e-5
    bins = 0.5 * (z[:, 1:] + z[:, :-1])
    assert (sp["samples"] >= bins[:, :1] - 1e-6).all() and (sp["samples"] <= bins[:, -1:] + 1e-6).all()
    # spike ray: every fine sample falls in the spike's bin [bins[29], bins[30]]
    # except the 61 * 1e-5 residual mass
    inside = (sp["samples"][1] >= bins[1, 29]) & (sp["samples"][1] <= bins[1, 30])
    assert inside.mean() > 0.97
    # sorted merge, stable permutation of cat(z, samples)
    assert (np.diff(sp["z"], axis=-1) >= 0).all() and sp["z"].shape == (R, Sc + Sf)
    cat = np.concatenate([z, sp["samples"]], -1)
    np.testing.assert_array_equal(np.sort(sp["perm"], -1), np.broadcast_to(np.arange(Sc + Sf), (R, Sc + Sf)))
    np.testing.assert_array_equal(np.take_along_axis(cat, sp["perm"], -1), sp["z"])
    # agreement with the textbook sequential-cumsum form
    s_ref, i_ref = sampling.sample_pdf_canonical(bins, w[:, 1:-1], u)
    assert (i_ref != sp["inds"]).mean() < 1e-3
    same = i_ref == sp["inds"]
    assert np.abs(s_ref - sp["samples"])[same].max() < 2e-4
    # deterministic u
    spd = sampling.sample_pdf(z, w, Sf, 6.0, None)
    assert (np.diff(spd["samples"], axis=-1) >= -1e-6).all()


def test_sample_pdf_general_sizes():
    rng = np.random.default_rng(2)
    for Sc, Sf in ((8, 16), (33, 7), (128, 64)):
        z = sampling.stratified(5, Sc, 0.0, 1.0, rng.random((5, Sc), dtype=f32))
        w = rng.random((5, Sc), dtype=f32)
        sp = sampling.sample_pdf(z, w, Sf, 1.0, rng.random((5, Sf), dtype=f32))
        assert sp["z"].shape == (5, Sc + Sf) and (np.diff(sp["cdf"], axis=-1) > 0).all()


# -------------------------------------------------------- whole pipeline
def test_render_and_train_step_runs():
    torch.manual_seed(0)
    sdc, sdf = mlp.init_state_dict(seed=42), mlp.init_state_dict(seed=43)
    R = 16
    rng = np.random.default_rng(0)
    o = np.tile(np.array([[0, 0, 4.0]], f32), (R, 1))
    d = np.array([0, 0, -1], f32) + 0.1 * rng.standard_normal((R, 3)).astype(f32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    us, up = rng.random((R, 16), dtype=f32), rng.random((R, 32), dtype=f32)
    out = render.render_rays_hier(sdc, sdf, o, d, 2.0, 6.0, 16, 32, us, up, white_bkgd=True)
    assert out["rgb"].shape == (R, 3) and out["t_starts"].shape == (R, 48)
    gt = rng.random((R, 3), dtype=f32)
    st = dict(step=0, m={}, v={})
    before = sdf["layers.0.weight"].clone()
    loss, psnr, grads = render.train_step(sdc, sdf, st, o, d, gt, 2.0, 6.0, 16, 32, us, up, 5e-4, True)
    assert np.isfinite(loss) and len(grads) == 48 and st["step"] == 1
    delta = (sdf["layers.0.weight"] - before).abs().max().item()
    assert 0 < delta <= 5e-4 * 1.001  # first Adam step moves each weight by <= lr


def test_misc_reference_known_answers(golden):
    g = golden("reference_misc.npz")
    lro, r, T = 5e-4, 0.1, 8000
    for t, lr in zip(g["lr_steps"], g["lrs"]):
        ours = lro * r ** (t / T) if t < T else lro * r
        assert abs(ours - lr) < 1e-12
    assert abs(g["lrs"][0] - 4.99856e-4) < 1e-9 and abs(g["lrs"][1] - 1.5811e-4) < 1e-8
    assert float(g["occ_reg"]) == 4.0


def test_doubleangle_encoding_close_to_reference_form():
    x = (torch.rand(4096, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1) * 4
    a = encoding.positional_encoding(x, 10)
    b = encoding.positional_encoding_doubleangle(x, 10)
    assert (a - b).abs().max().item() < 4e-6  # << bf16 rounding (4e-3) applied right after in the kernel
    assert torch.equal(a[:, :9], b[:, :9]) and torch.equal(a[:, 27:33], b[:, 27:33])


def test_scheduler_mirror_matches_reference_values(golden):
    """fsnerf_b200.core.scheduler (row a10) against values produced by the reference's
    core.scheduler.ExponentialDecay (tests/golden/reference_misc.npz, oracle/gen_golden.py)"""
    from fsnerf_b200.core.scheduler import Constant, ExponentialDecay, lr_at
    g = golden("reference_misc.npz")
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=5e-4)
    sch = ExponentialDecay(opt, 8000, 5e-4, r=0.1)
    seen = {}
    for t in range(1, 8100):
        sch.step()
        seen[t] = opt.param_groups[0]["lr"]
    for t, lr in zip(g["lr_steps"], g["lrs"]):
        assert abs(seen[int(t)] - lr) < 1e-12 and abs(lr_at(int(t), 8000, 5e-4, 0.1) - lr) < 1e-12
    assert seen[8099] == 5e-4 * 0.1 and sch.lrf == 5e-4 * 0.1
    c = Constant(opt, 10, 3e-4)
    c.step()
    assert opt.param_groups[0]["lr"] == 3e-4
    with pytest.raises(ValueError):
        Constant(opt, 10, -1.0)


# ------------------------------------------------------------------ in-step regularisers (f2)
def test_regularizers_match_reference(golden):
    """oracle/regularizers.py vs the reference's own core.loss.OcclusionRegularizer and the
    weight-penalty loop of src/run-nerf.py:266-279 (fixture: oracle/gen_golden_reg.py)"""
    from oracle import regularizers as oreg, mlp as omlp
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_reg.npz"))
    ri, t = torch.from_numpy(g["occ_ray_idx"]), torch.from_numpy(g["occ_t"])
    for func in ("linear", "exp"):
        a, b = (float(x) for x in g[f"occ_{func}_ab"])
        sig = torch.from_numpy(g["occ_sigma"]).requires_grad_(True)
        val = oreg.occlusion_reg(sig, t, ri, a, b, func)
        (gs,) = torch.autograd.grad(val, sig)
        np.testing.assert_allclose(val.item(), g[f"occ_{func}_value"], rtol=1e-6)
        np.testing.assert_allclose(gs.numpy(), g[f"occ_{func}_dsigma"], rtol=1e-6, atol=1e-8)
    e = torch.from_numpy(g["occ_dense_edges"])
    sg = torch.from_numpy(g["occ_dense_sigma"]).requires_grad_(True)
    val = oreg.occlusion_reg_dense(sg, e[:, :-1], e[:, 1:], 0.5, 2.0, "linear")
    (gs,) = torch.autograd.grad(val, sg)
    np.testing.assert_allclose(val.item(), g["occ_dense_value"], rtol=1e-6)
    np.testing.assert_allclose(gs.numpy(), g["occ_dense_dsigma"], rtol=1e-6, atol=1e-8)
    # the survey's pinned known answer
    v = oreg.occlusion_reg(torch.tensor([1., 1, 1, 2, 2]), torch.tensor([1., 2, 3, 1, 2]),
                           torch.tensor([0, 0, 0, 2, 2]), 0.5, 2.0, "linear")
    assert abs(v.item() - 4.0) < 1e-6
    with pytest.raises(ValueError):
        oreg.occlusion_weights(t, 1.0, 1.0, "cubic")

    sd = omlp.init_state_dict(seed=42)
    names = oreg.regularised_names([(k, tuple(v.shape)) for k, v in sd.items()])
    assert names == list(g["wreg_covered"]) and len(names) == 10
    assert "sigma.weight" not in names and "rgb.weight" not in names
    for mode in ("l1", "l2"):
        ps = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        val = oreg.weight_reg(ps, mode)
        gl, gb = torch.autograd.grad(val, [ps["layers.5.weight"], ps["branch.weight"]])
        np.testing.assert_allclose(val.item(), g[f"wreg_{mode}_value"], rtol=1e-6)
        np.testing.assert_allclose(gl[:6, :40].numpy(), g[f"wreg_{mode}_grad_layers5"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(gb[:6, :40].numpy(), g[f"wreg_{mode}_grad_branch"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose([gl.abs().sum().item(), gb.abs().sum().item()],
                                   g[f"wreg_{mode}_grad_abs_sums"], rtol=1e-5)


def test_dropin_occlusion_regularizer_matches_reference():
    """fsnerf_b200.core.loss (host-side mirror, torch ops only) vs the reference fixture"""
    from fsnerf_b200.core.loss import OcclusionRegularizer
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_reg.npz"))
    ri, t = torch.from_numpy(g["occ_ray_idx"]), torch.from_numpy(g["occ_t"])
    for func in ("linear", "exp"):
        a, b = (float(x) for x in g[f"occ_{func}_ab"])
        sig = torch.from_numpy(g["occ_sigma"]).requires_grad_(True)
        val = OcclusionRegularizer(a, b, func)(sig, t, ri)
        (gs,) = torch.autograd.grad(val, sig)
        np.testing.assert_allclose(val.item(), g[f"occ_{func}_value"], rtol=1e-6)
        np.testing.assert_allclose(gs.numpy(), g[f"occ_{func}_dsigma"], rtol=1e-6, atol=1e-8)
    with pytest.raises(AssertionError):
        OcclusionRegularizer(-1.0, 1.0)
    with pytest.raises(ValueError):
        OcclusionRegularizer(1.0, 1.0, "cubic")(sig, t, ri)


# ------------------------------------------------------------------ occupancy grid (f1)
def test_occgrid_oracle_known_answers():
    """closed-form checks of oracle/occgrid.py (the statement of semantics for csrc/occgrid.cu)"""
    from oracle import occgrid as oocc
    aabbs = oocc.level_aabbs([-1, -1, -1, 1, 1, 1], 2)
    np.testing.assert_array_equal(aabbs, np.array([[-1, -1, -1, 1, 1, 1], [-2, -2, -2, 2, 2, 2]], np.float32))
    binaries = np.zeros((1, 4, 4, 4), bool)
    binaries[0, 2, 2, 1] = True                      # cell x,y in [0,.5), z in [-.5,0)
    o = np.array([[0.25, 0.25, -3.0], [0.25, 0.25, -3.0], [5.0, 5.0, -3.0]], np.float32)
    d = np.array([[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0]], np.float32)
    ri, ts, te = oocc.march(o, d, binaries, aabbs[:1], 0.125, near_planes=np.array([0, 0.0625, 0], np.float32))
    # ray 0 enters at t=2: lattice 2 + k/8, midpoints inside z in [-.5,0) <=> t_mid in [2.5,3): k = 4..7
    np.testing.assert_array_equal(ri, [0, 0, 0, 0, 1, 1, 1, 1])
    np.testing.assert_allclose(ts[:4], [2.5, 2.625, 2.75, 2.875])
    np.testing.assert_allclose(te - ts, 0.125)
    # the lattice is anchored at the ray's own near plane (0 + u*step), not at the box entry, so the
    # stratified jitter survives for a camera outside the box: 0.0625 + k/8 with midpoints in [2.5,3)
    np.testing.assert_allclose(ts[4:], [2.4375, 2.5625, 2.6875, 2.8125])
    # camera inside the box: the lattice starts at the (jittered) near plane
    ri2, ts2, _ = oocc.march(np.array([[0.25, 0.25, -0.75]], np.float32), d[:1], binaries, aabbs[:1], 0.125,
                             near_planes=np.array([0.0625], np.float32))
    np.testing.assert_allclose(ts2, [0.1875, 0.3125, 0.4375, 0.5625])
    # two levels: the point is looked up in the finest level that contains it
    b2 = np.zeros((2, 4, 4, 4), bool)
    b2[1] = True
    ri3, ts3, _ = oocc.march(o[:1], d[:1], b2, aabbs, 0.5)
    np.testing.assert_allclose(ts3, [1.0, 1.5, 4.0, 4.5])  # outer shell only: level 0 is empty
    # visibility: trans = exp(-exclusive sum sigma*delta) >= eps
    keep = oocc.visibility(np.full(6, 40.0, np.float32), np.arange(6, dtype=np.float32) * 0.1,
                           np.arange(1, 7, dtype=np.float32) * 0.1, np.array([0, 0, 0, 0, 1, 1]), 1e-4)
    np.testing.assert_array_equal(keep, [True, True, True, False, True, True])
    # update + binarize
    occs = oocc.update(np.array([0.0, 0.2, 0.0, 0.4], np.float32), np.array([0.1, 0.05, 0.3], np.float32),
                       np.array([0, 1, 1]))
    np.testing.assert_allclose(occs, [0.1, 0.3, 0.0, 0.4], rtol=1e-6)
    b, thre = oocc.binarize(occs, 1e-2)
    assert thre == 1e-2 and list(b) == [True, True, False, True]
    b, thre = oocc.binarize(np.array([0.001, 0.003, 0.0, 0.0], np.float32), 1e-2)
    assert abs(thre - 0.001) < 1e-9 and list(b) == [False, True, False, False]


def test_sinerf_mirror_matches_reference(golden):
    """fsnerf_b200.core.models.SiNeRF (torch-op module mirror) vs the reference's SiNeRF: seeded
    construction, state dict, forward values and gradient norms (fixture: oracle/gen_golden_sinerf.py)"""
    from fsnerf_b200.core.models import SiNeRF
    g = golden("reference_sinerf.npz")
    torch.manual_seed(42)
    model = SiNeRF(3, 3, 256, [30.] + [1.] * 7)
    sd = model.state_dict()
    assert list(sd.keys()) == [str(n) for n in g["names"]]
    for i, (k, v) in enumerate(sd.items()):
        assert str(tuple(v.shape)) == str(g["shapes"][i]), k
        assert abs(v.double().sum().item() - g["w_sum"][i]) < 1e-9 and abs(v.double().abs().sum().item() - g["w_abs"][i]) < 1e-9, k
    x, d = torch.from_numpy(g["x"]), torch.from_numpy(g["d"])
    out = model(x, d)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], atol=1e-6)
    np.testing.assert_allclose(model(x).detach().numpy(), g["sigma_only"], atol=1e-6)
    loss = (out * torch.linspace(0.1, 1.0, 4)).sum()
    grads = torch.autograd.grad(loss, list(model.parameters()))
    np.testing.assert_allclose([gr.double().norm().item() for gr in grads], g["g_norm"], rtol=1e-5, atol=1e-9)


def test_rng_uniform_stream():
    """The seeded samplers' uniform stream: the vectorised oracle against a scalar restatement
    with Python integers (masking by hand), plus the basic statistics of a U[0,1) sample."""
    from oracle.sampling import rng_uniform
    M = 0xFFFFFFFF

    def scalar(seed, i):
        h = ((i & M) * 0x9E3779B1 + (i >> 32) * 0x85EBCA77 + (seed & M)) & M
        h ^= h >> 16; h = h * 0x85EBCA6B & M; h ^= h >> 13; h = h * 0xC2B2AE35 & M; h ^= h >> 16
        h = (h + (seed >> 32)) & M
        h ^= h >> 16; h = h * 0x7FEB352D & M; h ^= h >> 15; h = h * 0x846CA68B & M; h ^= h >> 16
        return (h >> 8) / float(1 << 24)

    for seed in (0, 1, 42, 0x9E3779B97F4A7C15, (1 << 64) - 1):
        u = rng_uniform(seed, 4096)
        assert u.dtype == np.float32
        assert [float(x) for x in u[:64]] == [scalar(seed, i) for i in range(64)]
        assert float(u[4095]) == scalar(seed, 4095)
    u = rng_uniform(7, 1 << 18)
    assert 0.0 <= u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 3e-3 and abs(u.var() * 12 - 1) < 1e-2
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 1e-2
    assert abs(np.corrcoef(u, rng_uniform(8, 1 << 18))[0, 1]) < 1e-2
