"""Error behaviour at the C ABI (include/fsnerf_b200.h: 0 / negative code + fsnerf_last_error();
the Python layer raises FsnerfError) and at the drop-in boundary, plus empty / ragged inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from fsnerf_b200 import ops
    ops.require_device(0)
    return torch.device("cuda:0")


def test_cabi_argument_errors(dev):
    import ctypes as C
    from fsnerf_b200 import ops, _lib
    from fsnerf_b200._lib import FsnerfError
    lib = _lib.load()
    # direct C call: negative code + message, nothing launched
    rc = lib.fsnerf_composite_forward(4, 8, None, None, None, None, None, 0, None, None, None, None, None, None, None)
    assert rc == -1 and b"null pointer" in lib.fsnerf_last_error()
    z = torch.sort(torch.rand(4, 2, device=dev), -1).values
    with pytest.raises(FsnerfError, match=r"n_coarse must be in \[3,258\]"):
        ops.sample_pdf(z, torch.rand(4, 2, device=dev), 8, 1.0, None)
    z = torch.sort(torch.rand(4, 16, device=dev), -1).values
    with pytest.raises(FsnerfError, match=r"n_fine must be in \[1,1024\]"):
        ops.sample_pdf(z, torch.rand(4, 16, device=dev), 2000, 1.0, None)
    raw = torch.rand(2, 600, 4, device=dev)
    e = torch.sort(torch.rand(2, 601, device=dev), -1).values
    with pytest.raises(FsnerfError, match=r"n_samples must be in \[1,512\]"):
        ops.composite_backward(raw, e[:, :-1].contiguous(), e[:, 1:].contiguous(), torch.rand(2, 3, device=dev))
    assert ops.composite_forward(raw, e[:, :-1].contiguous(), e[:, 1:].contiguous())[0].shape == (2, 3)  # fwd: any S
    with pytest.raises(FsnerfError, match="CUDA tensor"):
        ops.composite_forward(raw.cpu(), e[:, :-1].cpu(), e[:, 1:].cpu())
    with pytest.raises(FsnerfError, match="d_hidden must be 256"):
        ops.mlp_pack(ops.make_cfg(d_hidden=128), torch.zeros(10, device=dev))
    with pytest.raises(FsnerfError, match="skip"):
        ops.mlp_param_count(ops.make_cfg(n_layers=4, skip=(5,)))
    cfg = ops.make_cfg()
    params = torch.zeros(ops.mlp_param_count(cfg), device=dev)
    packed = ops.mlp_pack(cfg, params)
    x = torch.rand(10, 3, device=dev)
    with pytest.raises(FsnerfError, match="dirs required unless density_only"):
        ops.mlp_forward(cfg, params, packed, x=x, dirs=None)
    stash = torch.empty(ops.mlp_stash_bytes(cfg, 10), dtype=torch.uint8, device=dev)
    with pytest.raises(FsnerfError, match="needs the full network"):
        ops.mlp_forward(cfg, params, packed, x=x, density_only=True, stash=stash)
    with pytest.raises(FsnerfError, match="step counts from 1"):
        ops.adam_step(params, params, params.clone(), params.clone(), 1e-3, 0)
    with pytest.raises(FsnerfError, match="pose_rows must be 3 or 4"):
        ops.gen_rays(torch.eye(4, device=dev)[None, :2].contiguous(), 4, 4, 5.0, n_rays=16)
    torch.cuda.synchronize()  # nothing above may have poisoned the context


def test_empty_and_ragged_inputs(dev):
    from fsnerf_b200 import ops
    from fsnerf_b200.utils import utilities as U
    cfg = ops.make_cfg()
    params = torch.randn(ops.mlp_param_count(cfg), device=dev) * 0.05
    packed = ops.mlp_pack(cfg, params)
    out = ops.mlp_forward(cfg, params, packed, x=torch.zeros(0, 3, device=dev), dirs=torch.zeros(0, 3, device=dev))
    assert out.shape == (0, 4)
    ts, te = ops.sample_stratified(0, 8, 0.0, 1.0, None, device=dev)
    assert ts.shape == (0, 8)
    o, d, _ = ops.gen_rays(torch.eye(4, device=dev)[None], 10, 10, 12.0, first_id=0, n_rays=0)
    assert o.shape == (0, 3)
    # reference: get_chunks(10^4 rays, 4096) -> 4096, 4096, 1808; every chunk renders
    chunks = U.get_chunks(torch.arange(10000, device=dev), 4096)
    assert [len(c) for c in chunks] == [4096, 4096, 1808]
    # one sample short / long of a 128-sample tile give the same values for the shared samples
    x = torch.rand(129, 3, device=dev) * 2 - 1
    dd = torch.nn.functional.normalize(torch.randn(129, 3, device=dev), dim=-1)
    full = ops.mlp_forward(cfg, params, packed, x=x, dirs=dd)
    for n in (1, 127, 128):
        part = ops.mlp_forward(cfg, params, packed, x=x[:n].contiguous(), dirs=dd[:n].contiguous())
        assert torch.equal(part, full[:n]), n


def test_dropin_boundary_errors(dev):
    from fsnerf_b200._lib import FsnerfError
    from fsnerf_b200.core.models import NeRF, PositionalEncoder
    from fsnerf_b200.render.rendering import HierarchicalEstimator, OccGridEstimator, render_rays
    kw = {"pos_fn": {"n_freqs": 10, "log_space": True}, "dir_fn": {"n_freqs": 4, "log_space": True}}
    model = NeRF(3, 3, 8, 256, [4], **kw)
    with pytest.raises(FsnerfError, match="CUDA device"):
        model(torch.zeros(4, 3), torch.zeros(4, 3))  # parameters still on the CPU: no CPU path
    model = model.to(dev)
    est = HierarchicalEstimator(near=2.0, far=6.0, n_coarse=8, n_fine=0)
    with pytest.raises(FsnerfError, match="no CPU path"):
        render_rays(torch.zeros(4, 3), torch.zeros(4, 3), est, model, device=torch.device("cpu"))
    with pytest.raises(FsnerfError, match="proposal"):
        HierarchicalEstimator(near=2.0, far=6.0, n_coarse=8, n_fine=8)
    with pytest.raises(FsnerfError, match="CUDA tensor required"):
        PositionalEncoder(3, 10)(torch.zeros(4, 3))
    with pytest.raises(AssertionError, match=r"Expected \[6\] aabb"):
        OccGridEstimator([0, 0, 0, 1, 1], 8, 1)
    # stride-0 origins (reference get_rays returns an expanded view, utilities.py:80) are accepted
    o = torch.tensor([0.0, 0.0, 4.0], device=dev).expand(16, 3)
    d = torch.nn.functional.normalize(torch.tensor([0.0, 0.0, -1.0], device=dev) + 0.1 * torch.randn(16, 3, device=dev), dim=-1)
    with torch.no_grad():
        (rgb, op, dp, ex), ri, tv = render_rays(o, d, est, model, device=dev)
    assert rgb.shape == (16, 3) and ri.dtype == torch.int64 and ri.numel() == 16 * 8 == tv.numel()
    assert set(ex) >= {"weights", "alphas", "trans", "sigmas", "rgbs"}


def test_no_out_of_bounds_writes_on_partial_tiles(dev):
    """compute-sanitizer is not available on the pool: guard bands instead.  Every buffer the MLP
    kernels write (out, stash, dstash workspace, grads) is followed by a sentinel region that must
    survive launches on sample counts that are not multiples of the 128-sample tile."""
    from fsnerf_b200 import ops
    from oracle import mlp as omlp
    cfg = ops.make_cfg()
    params = ops.flatten_state_dict(cfg, omlp.init_state_dict(seed=5), dev)
    packed = ops.mlp_pack(cfg, params)
    n_par = ops.mlp_param_count(cfg)
    G = 4096  # guard bytes / floats
    for P in (1, 129, 128 * 3 + 77):
        g = torch.Generator().manual_seed(P)
        x = (torch.rand(P, 3, generator=g) * 2 - 1).to(dev)
        d = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1).to(dev)
        out_big = torch.full((P * 4 + G,), 7.25, device=dev)
        nb_stash, nb_ws = ops.mlp_stash_bytes(cfg, P), ops.mlp_bwd_workspace_bytes(cfg, P)
        stash_big = torch.full((nb_stash + G,), 0x5A, dtype=torch.uint8, device=dev)
        ws_big = torch.full((nb_ws + G,), 0xA5, dtype=torch.uint8, device=dev)
        grads_big = torch.zeros(n_par + G, device=dev)
        grads_big[n_par:] = -3.5
        out = ops.mlp_forward(cfg, params, packed, x=x, dirs=d, stash=stash_big[:nb_stash],
                              out=out_big[:P * 4].view(P, 4))
        d_out = torch.randn(P, 4, generator=g).to(dev)
        ops.mlp_backward(cfg, params, packed, P, stash_big[:nb_stash], out, d_out, grads_big[:n_par], ws_big[:nb_ws])
        torch.cuda.synchronize()
        assert bool((out_big[P * 4:] == 7.25).all()), P
        assert bool((stash_big[nb_stash:] == 0x5A).all()), P
        assert bool((ws_big[nb_ws:] == 0xA5).all()), P
        assert bool((grads_big[n_par:] == -3.5).all()), P
        assert torch.isfinite(out).all() and torch.isfinite(grads_big[:n_par]).all()
        # density-only forms write exactly [P] / the sigma slots
        sig_big = torch.full((P + G,), 7.25, device=dev)
        ops.mlp_forward(cfg, params, packed, x=x, density_only=True, out=sig_big[:P])
        raw_big = torch.full((P * 4 + G,), 7.25, device=dev)
        ops.mlp_forward(cfg, params, packed, x=x, density_only=2, out=raw_big[:P * 4].view(P, 4))
        torch.cuda.synchronize()
        assert bool((sig_big[P:] == 7.25).all()) and torch.equal(sig_big[:P], out[:, 3])
        assert bool((raw_big[P * 4:] == 7.25).all()) and bool((raw_big[:P * 4].view(P, 4)[:, :3] == 7.25).all())
