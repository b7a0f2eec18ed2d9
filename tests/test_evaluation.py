"""Evaluation + checkpoint mirror."""
import os

import numpy as np
import pytest
import torch


def test_ssim_matches_oracle_restatement():
    from fsnerf_b200.evaluation import ssim, psnr
    from oracle import metrics
    rng = np.random.default_rng(0)
    a = rng.random((3, 40, 52, 3)).astype(np.float32)
    b = np.clip(a + 0.1 * rng.standard_normal(a.shape).astype(np.float32), 0, 1)
    ours = ssim(torch.from_numpy(a), torch.from_numpy(b))
    ref = [metrics.ssim(x, y) for x, y in zip(a, b)]
    np.testing.assert_allclose(ours.numpy(), ref, rtol=1e-9)
    assert abs(ssim(torch.from_numpy(a), torch.from_numpy(a)).mean().item() - 1.0) < 1e-12
    assert 0.0 < ours.mean().item() < 0.95
    assert abs(psnr(0.01).item() - 20.0) < 1e-5  # -10 log10(mse), run-nerf.py:160


def test_checkpoint_roundtrip_is_reference_format(tmp_path, golden):
    """nn.pt = torch.save(model.state_dict()): the reference's 24 keys/shapes (golden) in order"""
    from fsnerf_b200.core.models import NeRF
    from fsnerf_b200.evaluation import save_checkpoint, load_checkpoint
    kw = {"pos_fn": {"n_freqs": 10, "log_space": True}, "dir_fn": {"n_freqs": 4, "log_space": True}}
    torch.manual_seed(42)
    model = NeRF(3, 3, 8, 256, [4], **kw)
    path = save_checkpoint(model, str(tmp_path))
    assert path.endswith(os.path.join("model", "nn.pt"))
    sd = torch.load(path)
    g = golden("reference_mlp.npz")
    ref_names = [str(n) for n in g["names"]]
    assert list(sd.keys()) == ref_names and len(ref_names) == 24
    for i, k in enumerate(ref_names):  # seed-42 construction reproduces the reference's initial weights
        assert str(tuple(sd[k].shape)) == str(g["shapes"][i])
        assert abs(sd[k].double().sum().item() - g["w_sum"][i]) < 1e-9
        assert abs(sd[k].double().abs().sum().item() - g["w_abs"][i]) < 1e-9
    other = NeRF(3, 3, 8, 256, [4], **kw)
    load_checkpoint(other, str(tmp_path))
    assert all(torch.equal(a, b) for a, b in zip(other.state_dict().values(), sd.values()))


@pytest.mark.gpu
def test_evaluation_loop_on_device(tmp_path):
    """evaluation() over an image-mode loader: PSNR/SSIM of rendered frames vs the same metrics
    computed from render_frame outputs directly; checkpoint -> HotPath.load_state_dict."""
    from torch.utils.data import DataLoader
    from fsnerf_b200 import synthetic as syn
    from fsnerf_b200.core.models import NeRF
    from fsnerf_b200.engine import HotPath
    from fsnerf_b200.evaluation import evaluation, save_checkpoint, ssim
    from fsnerf_b200.nerfdata.datasets.blender import BlenderDataset
    from fsnerf_b200.render.rendering import HierarchicalEstimator, render_frame
    from oracle import metrics
    dev = torch.device("cuda:0")
    syn.write_blender_scene(str(tmp_path / "synthetic" / "s"), n_views=3, H=24, W=24, seed=1, splits=("val",))
    np.random.seed(0)
    ds = BlenderDataset("s", "val", img_mode=True, white_bkgd=True, root=str(tmp_path / "synthetic"), device=dev)
    kw = {"pos_fn": {"n_freqs": 10, "log_space": True}, "dir_fn": {"n_freqs": 4, "log_space": True}}
    torch.manual_seed(42)
    coarse, fine = NeRF(3, 3, 8, 256, [4], **kw).to(dev), NeRF(3, 3, 8, 256, [4], **kw).to(dev)
    est = HierarchicalEstimator(near=ds.near, far=ds.far, n_coarse=16, n_fine=16, proposal_model=coarse)
    loader = DataLoader(ds, batch_size=1, shuffle=False)
    val_psnr, val_ssim, val_lpips = evaluation(ds.hwf, fine, est, None, loader, 200, dev, white_bkgd=True)
    assert val_lpips is None
    with torch.no_grad():
        frames = torch.stack([render_frame(ds.hwf, ds.near, ds.far, p, 200, est, fine, white_bkgd=True,
                                           device=dev)[0] for p in ds.poses])
    mse = ((frames.cpu() - ds.imgs) ** 2).mean()
    assert abs(val_psnr.item() - (-10 * torch.log10(mse)).item()) < 1e-4
    ref_ssim = np.mean([metrics.ssim(a, b) for a, b in zip(frames.cpu().numpy(), ds.imgs.numpy())])
    assert abs(val_ssim - ref_ssim) < 1e-6
    assert abs(ssim(frames, ds.imgs.to(dev)).mean().item() - ref_ssim) < 1e-6
    # checkpoint written by the drop-in model loads into the fused engine (same 24 keys)
    sd = torch.load(save_checkpoint(fine, str(tmp_path)))
    hp = HotPath(n_coarse=16, n_fine=16, device=dev)
    hp.load_state_dict(1, sd)
    assert all(torch.equal(hp.state_dict(1)[k].cpu(), v.cpu()) for k, v in sd.items())
