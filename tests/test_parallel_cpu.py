"""CPU tests of the N>1 path (world_size 2, gloo): ray sharding + gradient
all-reduce reproduce the single-process gradient of the global batch, and the
pixel partition of render_path covers every pixel exactly once."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fsnerf_b200 import parallel  # noqa: E402

KW = dict(n_layers=2, skip=(), n_freqs=2, n_freqs_dir=1)


def _tiny_sd(seed):
    from oracle import mlp as omlp
    return omlp.init_state_dict(n_layers=2, d_hidden=32, skip=(), n_freqs=2, n_freqs_dir=1, seed=seed)


def _local_grads(sd, o, d, gt, us, G):
    """gradient of sum-sq-err_local / (3G) through the oracle render (coarse only)"""
    from oracle import render as orender
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = orender.render_rays_hier(sdr, None, o, d, 2.0, 6.0, 8, 0, us, None, True, **KW)
    loss = ((out["rgb"] - torch.from_numpy(gt)) ** 2).sum() * parallel.loss_grad_scale(G)
    g = torch.autograd.grad(loss, list(sdr.values()))
    return torch.cat([x.reshape(-1) for x in g])


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    rng = np.random.default_rng(0)  # same global batch on every rank
    G = 10
    o = np.tile(np.array([[0, 0, 4.0]], np.float32), (G, 1))
    d = np.array([0, 0, -1], np.float32) + 0.1 * rng.standard_normal((G, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    gt = rng.random((G, 3), dtype=np.float32)
    us = rng.random((G, 8), dtype=np.float32)
    sd = _tiny_sd(42)
    a, b = parallel.shard_range(G, rank, world)
    flat = _local_grads(sd, o[a:b], d[a:b], gt[a:b], us[a:b], G)
    parallel.allreduce_gradients(flat)
    if rank == 0:
        torch.save(flat, os.path.join(tmp, "dp.pt"))
        torch.save(_local_grads(sd, o, d, gt, us, G), os.path.join(tmp, "single.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gradient_equals_single_process(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    dp, single = torch.load(tmp_path / "dp.pt"), torch.load(tmp_path / "single.pt")
    assert single.abs().max() > 0
    assert (dp - single).abs().max().item() <= 1e-6 * max(1.0, single.abs().max().item())


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 4096, 32768):
        for world in (1, 2, 3, 4, 8):
            r = [parallel.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    x = torch.arange(10)
    assert torch.equal(torch.cat([parallel.shard_batch(x, k, 3) for k in range(3)]), x)


def test_pixel_partition_covers_every_pixel_once():
    F, H, W = 5, 7, 9
    for world in (1, 2, 4, 8):
        seen = np.zeros(F * H * W, np.int32)
        for rank in range(world):
            for f, a, b in parallel.pixel_partition(F, H, W, rank, world):
                assert 0 <= a < b <= H * W
                seen[f * H * W + a: f * H * W + b] += 1
        assert (seen == 1).all()


class _FlatNet(torch.nn.Module):
    """stand-in for the drop-in NeRF: parameters are views of one flat buffer and the backward leaves the
    flat gradient in `_last_flat_grad` (what fsnerf_b200.core.models.NeRF does on the device)"""

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.zeros(3, 2))
        self.b = torch.nn.Parameter(torch.zeros(5))
        self._layout = [(0, 6), (8, 5)]  # 16-byte aligned offsets, like fsnerf_mlp_param_layout
        self._last_flat_grad = None

    def _param_list(self):
        return [self.a, self.b]

    def fake_backward(self, flat):
        self._last_flat_grad = flat
        for (o, n), p in zip(self._layout, self._param_list()):
            p.grad = flat[o:o + n].view(p.shape)


def _worker_modules(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat_net, plain = _FlatNet(), torch.nn.Linear(4, 3)
    flat_net.fake_backward(torch.arange(16, dtype=torch.float32) * (rank + 1))
    for p in plain.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    parallel.allreduce_module_gradients([flat_net, plain])
    if rank == 0:
        torch.save(dict(a=flat_net.a.grad.clone(), b=flat_net.b.grad.clone(), flat=flat_net._last_flat_grad.clone(),
                        w=plain.weight.grad.clone(), bias=plain.bias.grad.clone()), os.path.join(tmp, "mod.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_module_gradients(tmp_path):
    """the reference-style loop's exchange: one collective per flat-buffer network (aliasing kept), a
    coalesced copy for any other module"""
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_modules, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = torch.load(tmp_path / "mod.pt")
    want = torch.arange(16, dtype=torch.float32) * 3  # ranks contribute x1 and x2
    assert torch.equal(r["flat"], want)
    assert torch.equal(r["a"], want[0:6].view(3, 2)) and torch.equal(r["b"], want[8:13])
    assert torch.equal(r["w"], torch.full((3, 4), 3.0)) and torch.equal(r["bias"], torch.full((3,), 3.0))


def test_jitter_seed_streams_are_distinct():
    """engine.jitter_seeds: every (rank, step, sampler) gets its own 64-bit key, reproducibly"""
    from fsnerf_b200.engine import jitter_seeds
    keys = set()
    for rank in range(8):
        for draw in range(1, 200):
            a, b = jitter_seeds(42, rank, draw)
            assert 0 <= a < (1 << 64) and 0 <= b < (1 << 64)
            keys.update((a, b))
    assert len(keys) == 8 * 199 * 2
    assert jitter_seeds(42, 3, 17) == jitter_seeds(42, 3, 17)
    assert jitter_seeds(42, 3, 17) != jitter_seeds(43, 3, 17)
