"""On-disk formats + loaders: the nerfdata mirror against fixtures produced by
the REFERENCE's own loaders (oracle/gen_golden_data.py) on the same procedurally written
scenes.  Host logic on CPU; ray tables and the device-resident batch source on the GPU."""
import os

import numpy as np
import pytest
import torch

from fsnerf_b200 import synthetic as syn


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_data.npz"))


@pytest.fixture(scope="module")
def scenes(tmp_path_factory):
    root = tmp_path_factory.mktemp("datasets")
    syn.write_llff_scene(str(root / "llff"), "synth", n_views=9, H=24, W=32, seed=42)
    syn.write_blender_scene(str(root / "synthetic" / "synth"), n_views=5, H=20, W=20, seed=42, splits=("train",))
    return root


def _splitter(scenes, **kw):
    from fsnerf_b200.nerfdata.utils.splitter import Splitter
    return Splitter("llff", "synth", root=str(scenes / "llff"), **kw)


def test_llff_splitter_matches_reference(ref, scenes):
    sp = _splitter(scenes, n_training_views=3)
    np.testing.assert_allclose(sp.poses, ref["llff_poses"], rtol=0, atol=1e-6)
    assert sp.poses.dtype == np.float32 and sp.poses.shape == (9, 3, 4)
    assert sp.hwf[:2] == (24, 32) and abs(sp.hwf[2] - ref["llff_hwf"][2]) < 1e-4
    np.testing.assert_allclose([sp.min_bound, sp.max_bound], [ref["llff_min_bound"], ref["llff_max_bound"]], atol=1e-5)
    np.testing.assert_allclose(sp.path_poses, ref["llff_path_poses"], rtol=0, atol=1e-5)
    np.random.seed(0)
    sp.split()
    assert list(sp.test_ids) == list(ref["llff_test_ids"]) and list(sp.val_ids) == list(ref["llff_val_ids"])
    assert list(sp.train_ids) == list(ref["llff_train_ids"])
    ids = np.concatenate([sp.test_ids, sp.val_ids, sp.train_ids])
    assert len(set(ids.tolist())) == len(ids)
    train, val, test = sp.get_datasets(False, white_bkgd=False, ndc=True, device="cpu")
    assert [train.near, train.far] == list(ref["llff_train_near_far"])
    assert len(train) == int(ref["llff_train_len"]) and len(val) == int(ref["llff_val_len"])
    img0, pose0 = val[0]
    np.testing.assert_allclose(img0[::5, ::5].numpy(), ref["llff_val_img0"], atol=1e-7)
    np.testing.assert_allclose(pose0.numpy(), ref["llff_val_pose0"], atol=1e-6)
    world, _, _ = sp.get_datasets(False, ndc=False, device="cpu")
    np.testing.assert_allclose([world.near, world.far], ref["llff_world_near_far"], rtol=1e-6)


def test_llff_splitter_errors(scenes):
    from fsnerf_b200.nerfdata.utils.splitter import Splitter
    with pytest.raises(ValueError, match="not supported"):
        Splitter("synthetic", "synth", root=str(scenes / "llff"))
    with pytest.raises(AssertionError, match="not found"):
        Splitter("llff", "nope", root=str(scenes / "llff"))
    sp = _splitter(scenes)
    with pytest.raises(AssertionError, match="Split the source data"):
        sp.get_datasets()
    sp.n_training_views = -1
    np.random.seed(1)
    sp.split()
    assert len(sp.train_ids) == 9 - 2  # all remaining views


def test_blender_loader_matches_reference(ref, scenes):
    from fsnerf_b200.nerfdata.datasets.blender import BlenderDataset
    np.random.seed(0)
    ds = BlenderDataset("synth", "train", img_mode=True, white_bkgd=False, root=str(scenes / "synthetic"),
                        device="cpu")
    np.testing.assert_allclose(ds.poses.numpy(), ref["blender_poses"], atol=0)
    assert ds.hwf[:2] == (20, 20) and abs(ds.hwf[2] - ref["blender_hwf"][2]) < 1e-9
    np.testing.assert_allclose(ds.path_poses.numpy(), ref["blender_path_poses"], atol=1e-6)
    assert (ds.near, ds.far, ds.ndc) == (2.0, 6.0, False) and len(ds) == 5
    white = BlenderDataset("synth", "train", img_mode=True, white_bkgd=True, root=str(scenes / "synthetic"),
                           device="cpu")
    rgba = ref["blender_imgs_sub"]
    np.testing.assert_allclose(ds.imgs[:, ::4, ::4].numpy(), rgba[..., :3], atol=1e-7)
    np.testing.assert_allclose(white.imgs[:, ::4, ::4].numpy(), rgba[..., :3] * rgba[..., 3:] + (1 - rgba[..., 3:]),
                               atol=1e-6)
    # view selection (the reference line raises IndexError; intent documented in the module)
    np.random.seed(0)
    few = BlenderDataset("synth", "train", n_imgs=3, img_mode=True, root=str(scenes / "synthetic"), device="cpu")
    assert few.imgs.shape[0] == 3 and few.poses.shape == (3, 4, 4)
    assert all(any(torch.equal(p, q) for q in ds.poses) for p in few.poses)


@pytest.mark.gpu
def test_llff_ray_table_and_device_loader(ref, scenes):
    """NDC ray table built by the ray-generation kernel == the reference's host table (get_rays +
    to_ndc per pose, llff.py:59-90); DeviceRayLoader batches == rows of that table; every ray is
    visited exactly once per epoch."""
    dev = torch.device("cuda:0")
    sp = _splitter(scenes, n_training_views=3)
    np.random.seed(0)
    sp.split()
    train, _, _ = sp.get_datasets(False, white_bkgd=False, ndc=True, device=dev)
    np.testing.assert_allclose(train.rays_o[::97].cpu().numpy(), ref["llff_train_rays_o"], atol=2e-6)
    np.testing.assert_allclose(train.rays_d[::97].cpu().numpy(), ref["llff_train_rays_d"], atol=2e-6)
    np.testing.assert_allclose(train.rgb[::97].numpy(), ref["llff_train_rgb"], atol=1e-7)
    np.testing.assert_allclose(train.aabb.cpu().numpy(), ref["llff_train_aabb"], atol=2e-6)
    o, d, c = train[97]
    np.testing.assert_allclose(o.cpu().numpy(), ref["llff_train_rays_o"][1], atol=2e-6)
    world, _, _ = sp.get_datasets(False, ndc=False, device=dev)
    np.testing.assert_allclose(world.rays_d[::97].cpu().numpy(), ref["llff_world_rays_d"], atol=2e-6)
    np.testing.assert_allclose(world.aabb.cpu().numpy(), ref["llff_world_aabb"], atol=0)

    loader = train.device_loader(500, seed=3)
    assert len(loader) == -(-len(train) // 500)
    seen = torch.zeros(len(train), dtype=torch.int32, device=dev)
    perm = torch.randperm(len(train), generator=torch.Generator(device=dev).manual_seed(3), device=dev)
    n = 0
    for i, (ro, rd, gt) in enumerate(loader):
        ids = perm[i * 500:(i + 1) * 500]
        seen[ids] += 1
        assert torch.equal(ro, train.rays_o[ids]) and torch.equal(rd, train.rays_d[ids])
        assert torch.equal(gt.cpu(), train.rgb[ids.cpu()])
        n += ro.shape[0]
    assert n == len(train) and bool((seen == 1).all())
    loaders = sp.get_dataloaders(256, white_bkgd=False, ndc=True, device=dev)
    ro, rd, gt = next(iter(loaders[0]))
    assert ro.shape == (256, 3) and ro.is_cuda and loaders[0].dataset.hwf == sp.hwf
    img, pose = next(iter(loaders[1]))
    assert img.shape == (1, 24, 32, 3) and pose.shape == (1, 3, 4)


@pytest.mark.gpu
def test_blender_ray_table(scenes):
    from fsnerf_b200.nerfdata.datasets.blender import BlenderDataset
    from oracle import rays as orays
    dev = torch.device("cuda:0")
    np.random.seed(0)
    ds = BlenderDataset("synth", "train", white_bkgd=True, root=str(scenes / "synthetic"), device=dev)
    H, W, f = ds.hwf
    ids = np.arange(0, 5 * H * W, 37).astype(np.int64)
    o, d = orays.rays_from_pixel_ids(ds.poses.numpy(), (H, W, f), ids)
    np.testing.assert_allclose(ds.rays_o[ids].cpu().numpy(), o, atol=1e-6)
    np.testing.assert_allclose(ds.rays_d[ids].cpu().numpy(), d, atol=1e-6)
    ro, rd, gt = ds[123]
    assert torch.equal(gt, ds.imgs.reshape(-1, 3)[123])
