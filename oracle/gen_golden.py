"""Generate tests/golden/reference_*.npz by running the REFERENCE's own Python.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.gen_golden

Imports /root/reference/src/{core.models, core.scheduler, core.loss,
utils.utilities, render.rendering} unmodified.  The last two import
third-party modules that are absent here (nerfacc, imageio, matplotlib); those
are replaced by empty ``sys.modules`` stubs, and ``nerfacc.volrend.rendering``
by oracle.compositing.render_packed (the restated algorithm) — so the
render_rays fixture pins the reference's *call-site plumbing* (midpoint
positions, per-sample dir gather, rgb/sigma split, background, t_vals), not
nerfacc's arithmetic.  Nothing under /root/reference is copied into the repo.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _install_stubs():
    from oracle.compositing import render_packed

    def rendering(t_starts, t_ends, ray_indices, n_rays, rgb_sigma_fn, render_bkgd):
        assert t_starts.shape == t_ends.shape == ray_indices.shape
        rgbs, sigmas = rgb_sigma_fn(t_starts, t_ends, ray_indices)
        assert rgbs.shape[-1] == 3 and sigmas.shape == t_starts.shape
        return render_packed(t_starts, t_ends, ray_indices, n_rays, rgbs, sigmas, render_bkgd)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("nerfacc")
    mod("nerfacc.volrend", rendering=rendering)
    mod("nerfacc.estimators")
    mod("nerfacc.estimators.occ_grid", OccGridEstimator=object)
    mod("imageio")
    mod("matplotlib", cm=mod("matplotlib.cm"), pyplot=mod("matplotlib.pyplot"))
    mod("mpl_toolkits")
    mod("mpl_toolkits.mplot3d", axes3d=None)


class StubEstimator:
    """Estimator-shaped object handing fixed intervals to the reference's
    render_rays (stands in for OccGridEstimator.sampling)."""

    def __init__(self, t_starts, t_ends):
        self.ts, self.te = t_starts, t_ends
        self.seen = {}

    def sampling(self, rays_o, rays_d, sigma_fn=None, render_step_size=None,
                 stratified=None, near_plane=None, far_plane=None):
        self.seen = dict(stratified=stratified, near_plane=near_plane, far_plane=far_plane)
        R, S = self.ts.shape
        ri = torch.arange(R).repeat_interleave(S)
        return ri, self.ts.reshape(-1), self.te.reshape(-1)


def main():
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    _install_stubs()
    import core.models as M
    import core.scheduler as S
    import core.loss as L
    import utils.utilities as U
    import render.rendering as R

    # ---- rays -----------------------------------------------------------
    fov = 0.6911112
    H = W = 100
    focal = 0.5 * W / np.tan(0.5 * fov)
    pose = torch.eye(4)
    pose[2, 3] = 4.0
    ro, rd = U.get_rays(pose, (H, W, focal))
    no, nd = U.to_ndc(ro.reshape(-1, 3), rd.reshape(-1, 3), (H, W, focal), 1.0)
    g = torch.Generator().manual_seed(0)
    # a generic rotated pose on a small non-square image (full arrays kept)
    A = torch.randn(3, 3, generator=g)
    Q, _ = torch.linalg.qr(A)
    pose2 = torch.eye(4)
    pose2[:3, :3] = Q
    pose2[:3, 3] = torch.tensor([0.3, -1.2, 3.5])
    H2, W2, f2 = 12, 20, 17.25
    ro2, rd2 = U.get_rays(pose2, (H2, W2, f2))
    no2, nd2 = U.to_ndc(ro2.reshape(-1, 3), rd2.reshape(-1, 3), (H2, W2, f2), 1.0)
    chunks = [len(c) for c in U.get_chunks(torch.zeros(10000, 3), 4096)]
    np.savez(os.path.join(OUT, "reference_rays.npz"),
             focal=np.float64(focal), pose=pose.numpy(),
             rd_00=rd[0, 0].numpy(), rd_center=rd[50, 50].numpy(), rd_last=rd[99, 99].numpy(),
             ro_00=ro[0, 0].numpy(), rd_row7=rd[7].numpy(),
             ndc_o_0=no[0].numpy(), ndc_d_0=nd[0].numpy(),
             ndc_o_row=no[700:720].numpy(), ndc_d_row=nd[700:720].numpy(),
             norm_min=np.float32(rd.norm(dim=-1).min()), norm_max=np.float32(rd.norm(dim=-1).max()),
             pose2=pose2.numpy(), hwf2=np.array([H2, W2, f2]), ro2=ro2.numpy(), rd2=rd2.numpy(),
             ndc_o2=no2.numpy(), ndc_d2=nd2.numpy(), chunks=np.array(chunks),
             origin_stride=np.array(ro.stride()))

    # ---- encoder + MLP --------------------------------------------------
    torch.manual_seed(42)
    kw = {"pos_fn": {"n_freqs": 10, "log_space": True}, "dir_fn": {"n_freqs": 4, "log_space": True}}
    model = M.NeRF(3, 3, 8, 256, [4], **kw)
    sd = model.state_dict()
    names = list(sd.keys())
    shapes = [tuple(v.shape) for v in sd.values()]
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(64, 3, generator=g) * 2 - 1) * 3.0
    d = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=-1)
    pe_pos = M.PositionalEncoder(3, 10, True)(x)
    pe_dir = M.PositionalEncoder(3, 4, True)(d)
    pe_lin = M.PositionalEncoder(3, 4, False)(d)
    out = model(x, d)
    sig = model(x)
    cvec = torch.linspace(-1.0, 1.0, 4)
    loss = (out * cvec).sum()
    grads = torch.autograd.grad(loss, list(model.parameters()))
    gnames = [n for n, _ in model.named_parameters()]
    np.savez(os.path.join(OUT, "reference_mlp.npz"),
             names=np.array(names), shapes=np.array([str(s) for s in shapes]),
             n_params=np.int64(sum(v.numel() for v in sd.values())),
             w_sum=np.array([v.double().sum().item() for v in sd.values()]),
             w_abs=np.array([v.double().abs().sum().item() for v in sd.values()]),
             w_head=np.stack([v.flatten()[:4].numpy() if v.numel() >= 4 else
                              np.pad(v.flatten().numpy(), (0, 4 - v.numel())) for v in sd.values()]),
             x=x.numpy(), d=d.numpy(), pe_pos=pe_pos.numpy(), pe_dir=pe_dir.numpy(),
             pe_lin=pe_lin.numpy(),
             pe_pin=M.PositionalEncoder(3, 10, True)(torch.tensor([[0.1, -0.2, 0.3]])).numpy(),
             out=out.detach().numpy(), sigma_only=sig.detach().numpy(), cvec=cvec.numpy(),
             gnames=np.array(gnames),
             g_norm=np.array([gr.double().norm().item() for gr in grads]),
             g_sum=np.array([gr.double().sum().item() for gr in grads]),
             g_layers3=grads[gnames.index("layers.3.weight")][:8, :8].numpy(),
             g_branch_b=grads[gnames.index("branch.bias")].numpy(),
             g_rgb_w=grads[gnames.index("rgb.weight")].numpy(),
             g_sigma_w=grads[gnames.index("sigma.weight")].numpy())

    # ---- render_rays call-site plumbing ---------------------------------
    Rr, Ss = 24, 16
    g = torch.Generator().manual_seed(2)
    rays_o = torch.tensor([0.0, 0.0, 4.0]).expand(Rr, 3) + 0.05 * torch.randn(Rr, 3, generator=g)
    rays_d = torch.nn.functional.normalize(
        torch.tensor([0.0, 0.0, -1.0]) + 0.2 * torch.randn(Rr, 3, generator=g), dim=-1)
    edges = torch.sort(2.0 + 4.0 * torch.rand(Rr, Ss + 1, generator=g), -1).values
    ts, te = edges[:, :-1].contiguous(), edges[:, 1:].contiguous()
    # scale sigma so the render is not trivially transparent
    with torch.no_grad():
        model.sigma.weight.mul_(40.0)
        model.sigma.bias.add_(1.0)
    res = {}
    for white in (False, True):
        est = StubEstimator(ts, te)
        (rgb, opac, depth, extras), ri, tv = R.render_rays(
            rays_o, rays_d, est, model, train=True, white_bkgd=white)
        tag = "w" if white else "b"
        res.update({f"rgb_{tag}": rgb.detach().numpy(), f"opacity_{tag}": opac.detach().numpy(),
                    f"depth_{tag}": depth.detach().numpy(),
                    f"weights_{tag}": extras["weights"].detach().numpy(),
                    f"sigmas_{tag}": extras["sigmas"].detach().numpy()})
        res["ray_indices"] = ri.numpy()
        res["t_vals"] = tv.numpy()
        res["seen_stratified"] = np.array(est.seen["stratified"])
        res["seen_far"] = np.float64(est.seen["far_plane"])
    np.savez(os.path.join(OUT, "reference_render.npz"), rays_o=rays_o.numpy(),
             rays_d=rays_d.numpy(), t_starts=ts.numpy(), t_ends=te.numpy(),
             sigma_w_scale=np.float32(40.0), sigma_b_add=np.float32(1.0), **res)

    # ---- scheduler / regulariser known answers --------------------------
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=5e-4)
    sch = S.ExponentialDecay(opt, 8000, 5e-4, r=0.1)
    lrs = []
    for t in range(1, 8003):
        sch.step()
        if t in (1, 4000, 7999, 8000, 8002):
            lrs.append(sch.lr)
    occ = L.OcclusionRegularizer(0.5, 2, "linear")(
        torch.tensor([1.0, 1, 1, 2, 2]), torch.tensor([1.0, 2, 3, 1, 2]), torch.tensor([0, 0, 0, 2, 2]))
    np.savez(os.path.join(OUT, "reference_misc.npz"), lr_steps=np.array([1, 4000, 7999, 8000, 8002]),
             lrs=np.array(lrs), occ_reg=np.float32(occ.item()))
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
