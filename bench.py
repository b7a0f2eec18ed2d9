#!/usr/bin/env python
"""bench.py — train rays/s (fwd+bwd+Adam) of the fs-nerf ray-march hot path on N B200s,
with render Mrays/s, the live roofline of the dominant kernel and the CPU baseline.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --steps 3 --warmup 1        # CPU arm

Workload (BASELINE.json configs[1]): synthetic Blender-format scene, 8 views at
400x400, coarse+fine NeRF 8x256 (64 + 128 samples, hierarchical sample_pdf),
4096-ray batches per GPU (weak scaling: global batch 4096*N; N=8 is config[4]'s
32768-ray global batch).  One "step" = one optimisation step: ray generation
from pixel ids, stratified + sample_pdf sampling, coarse+fine MLP forward, compositing,
MSE, full backward, [NCCL all-reduce of the flat gradient], Adam.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 400
N_VIEWS = 8
R_PER_GPU = 4096
N_COARSE, N_FINE = 64, 128
NEAR, FAR = 2.0, 6.0
F_FWD = 1_186_816      # FLOP per sample evaluation, forward
F_TRAIN = 3_489_024    # forward + backward
F_DENSITY = 982_528    # density-only forward (the coarse pass of a hierarchical RENDER skips the view branch)
METRIC = "train_rays_per_s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled (NVML) every ~10 ms during the timed region"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.01)
        except Exception as e:  # NVML unavailable: fall back to one nvidia-smi sample
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                a, b = out.stdout.strip().split(",")
                self.sm.append(float(a))
                self.max_mhz = float(b)
            except Exception:
                self.reasons.add(f"unavailable: {e}")

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def make_scene():
    from fsnerf_b200 import synthetic as syn
    return syn.make_views(N_VIEWS, H, W, seed=42)


# ----------------------------------------------------------------------------- CPU arm
def cpu_step_fn(sample_rays):
    """One optimisation step of the ORACLE PORT (reference's Python NeRF math restated,
    oracle/render.py) on `sample_rays` rays of the same workload; returns a closure."""
    import torch
    from oracle import mlp as omlp, render as orender, rays as orays
    poses, imgs, focal = make_scene()
    torch.set_num_threads(os.cpu_count())
    rng = np.random.default_rng(0)
    sdc, sdf = omlp.init_state_dict(seed=42), omlp.init_state_dict(seed=43)
    st = dict(step=0, m={}, v={})

    def step():
        ids = rng.permutation(N_VIEWS * H * W)[:sample_rays].astype(np.int64)
        o, d = orays.rays_from_pixel_ids(poses, (H, W, focal), ids)
        gt = imgs.reshape(-1, 3)[ids]
        us = rng.random((sample_rays, N_COARSE), dtype=np.float32)
        up = rng.random((sample_rays, N_FINE), dtype=np.float32)
        orender.train_step(sdc, sdf, st, o, d, gt, NEAR, FAR, N_COARSE, N_FINE, us, up, 5e-4, True)
    return step


def cpu_baseline(sample_rays=512, steps=16, warmup=1):
    step = cpu_step_fn(sample_rays)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": sample_rays / dt, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{steps} optimisation steps of {sample_rays} rays x ({N_COARSE}+{N_COARSE + N_FINE}) "
                      f"sample evaluations of the same workload (oracle/render.py train_step, fp32 torch-CPU, "
                      f"{os.cpu_count()} threads); {dt:.2f} s/step"}, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 512
    base, dt = cpu_baseline(sample, steps=max(1, args.steps), warmup=max(1, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "rays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference arm = the reference's PyTorch NeRF math on the host CPU (oracle port: the "
                    "reference itself cannot run here, its sampler/compositor nerfacc 0.5.3 is not installable); "
                    "each step is a bounded 512-ray sample of the 4096-ray workload step"}
    print(json.dumps(line))


def workload_config(n):
    return {"workload": f"C2: synthetic Blender-format scene {N_VIEWS} views {H}x{W}, coarse+fine NeRF 8x256, "
                        f"{N_COARSE}+{N_FINE} samples/ray (hierarchical sample_pdf), {R_PER_GPU}-ray batch per GPU",
            "rays_per_gpu": R_PER_GPU, "global_rays": R_PER_GPU * n, "n_coarse": N_COARSE, "n_fine": N_FINE,
            "parallelism": f"dp{n} (ray-sharded, one NCCL all-reduce of the flat fp32 gradient)" if n > 1 else "single GPU",
            "l2": "per-step working set (>9 GB of activation stash) exceeds the 126 MB L2; no explicit flush"}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` (mean over the launches
    of one step) from the newest committed `ncu --set full` capture (profiles/*_traffic.json,
    written by tools/ncu_table.py); None when there is no capture of that kernel."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    key = {"mlp_fwd_train": "mlp_fwd"}.get(kernel, kernel)
    rows = [r for k, v in d.items() if k.startswith(key) for r in v]
    if not rows:
        return None, None
    return (sum(r["dram_read_bytes"] + r["dram_write_bytes"] for r in rows) / len(rows),
            os.path.relpath(files[-1], ROOT) + f" ({len(rows)} launches)")


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--no-micro", action="store_true", help="skip the render-scale compositing / sampling rooflines")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from fsnerf_b200 import ops
    from fsnerf_b200.engine import HotPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None)  # NCCL_DEBUG>=VERSION prints a banner on stdout; keep it to the ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    ops.require_device(local)
    W_ = max(3, args.warmup)
    K = args.steps

    poses, imgs, focal = make_scene()
    poses_d = torch.from_numpy(poses).to(dev)
    imgs_d = torch.from_numpy(imgs).to(dev)
    hp = HotPath(n_coarse=N_COARSE, n_fine=N_FINE, near=NEAR, far=FAR, white_bkgd=True, device=dev, lr=5e-4)
    G = R_PER_GPU * world
    g = torch.Generator().manual_seed(1234)
    n_pix = N_VIEWS * H * W
    total_steps = W_ + K
    # seeded pixel permutation, sliced per step and per rank
    perm = torch.stack([torch.randperm(n_pix, generator=g)[:G] for _ in range(total_steps)])  # [steps, G]
    ids_all = perm[:, rank * R_PER_GPU:(rank + 1) * R_PER_GPU].contiguous()
    ids_dev = ids_all.to(dev)

    def step_resident(i):
        ro, rd, gt = ops.gen_rays(poses_d, H, W, focal, pixel_ids=ids_dev[i], images=imgs_d)
        hp.launches += 1
        return hp.train_step(ro, rd, gt, global_rays=G)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    for i in range(W_):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.profile_enable(True)
    l0 = hp.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        ls = step_resident(W_ + i)
    ev1.record()
    barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    launches = hp.launches - l0
    prof = ops.profile_read()
    ops.profile_enable(False)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = G * K / (ms / 1e3)
    final_loss = (ls[0].item() + ls[1].item()) / (3 * R_PER_GPU)

    # ---- e2e: host buffers through the public API (H2D of the batch, D2H of the loss, every step)
    from fsnerf_b200 import synthetic as syn
    Ke = K
    host_batches = []
    ray_tab_o, ray_tab_d = [], []
    for v in range(N_VIEWS):  # the reference's dataset precomputes the ray table on the host (llff.py:59-90)
        o_, d_ = syn.camera_rays(poses[v], H, W, focal)
        ray_tab_o.append(o_.reshape(-1, 3))
        ray_tab_d.append(d_.reshape(-1, 3))
    ray_tab_o, ray_tab_d = np.concatenate(ray_tab_o), np.concatenate(ray_tab_d)
    rgb_tab = imgs.reshape(-1, 3)
    for i in range(W_ + Ke):
        idx = ids_all[i % total_steps].numpy()
        host_batches.append(tuple(torch.from_numpy(np.ascontiguousarray(a[idx])).pin_memory()
                                  for a in (ray_tab_o, ray_tab_d, rgb_tab)))
    loss_host = torch.zeros(2).pin_memory()

    def step_host(i):
        ro, rd, gt = (b.to(dev, non_blocking=True) for b in host_batches[i])
        ls_ = hp.train_step(ro, rd, gt, global_rays=G)
        loss_host.copy_(ls_, non_blocking=False)  # D2H read of the step's loss (synchronises)
        return loss_host

    for i in range(W_):
        step_host(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(Ke):
        step_host(W_ + i)
    e1.record()
    barrier()
    ms_e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e = t.item()
    e2e_value = G * Ke / (ms_e / 1e3)
    h2d = 3 * R_PER_GPU * 3 * 4
    d2h = 8

    # ---- the reference's own train loop (src/run-nerf.py:232-285) on the drop-in modules: render_rays
    #      through autograd, F.mse_loss, loss.backward(), torch.optim.Adam.  Reported beside the
    #      fused engine so the cost of staying inside the reference's loop structure is visible.
    dropin = None
    if rank == 0 and not args.no_dropin:
        import torch.nn.functional as Fnn
        from fsnerf_b200.core.models import NeRF
        from fsnerf_b200.render.rendering import HierarchicalEstimator, render_rays
        kw = {"pos_fn": {"n_freqs": 10, "log_space": True}, "dir_fn": {"n_freqs": 4, "log_space": True}}
        torch.manual_seed(42)
        coarse, fine = NeRF(3, 3, 8, 256, [4], **kw).to(dev), NeRF(3, 3, 8, 256, [4], **kw).to(dev)
        est = HierarchicalEstimator(near=NEAR, far=FAR, n_coarse=N_COARSE, n_fine=N_FINE, proposal_model=coarse)
        opt = torch.optim.Adam(list(fine.parameters()) + list(coarse.parameters()), lr=5e-4)
        fine.train(); coarse.train(); est.train()

        def step_dropin(i):
            ro, rd, gt = (b.to(dev, non_blocking=True) for b in host_batches[i])
            (rgb, *_, extras), _, _ = render_rays(ro, rd, est, fine, train=True, white_bkgd=True, device=dev)
            loss = Fnn.mse_loss(rgb, gt) + Fnn.mse_loss(extras["rgb_coarse"], gt)
            loss.backward()
            opt.step()
            opt.zero_grad()
            return loss
        Kd = min(K, 10)
        for i in range(W_):
            step_dropin(i)
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for i in range(Kd):
            last = step_dropin(W_ + i)
        float(last.detach())  # D2H read of the loss
        d1.record()
        torch.cuda.synchronize()
        ms_d = d0.elapsed_time(d1) / Kd
        dropin = {"value": R_PER_GPU / (ms_d / 1e3), "unit": "rays/s", "ms_per_step": ms_d, "steps": Kd,
                  "what": "render_rays (autograd) + F.mse_loss + loss.backward() + torch.optim.Adam, "
                          "host batches, one GPU"}
        del coarse, fine, est, opt

    # ---- render Mrays/s (rank-local pixel slice of one 800x800 frame; no collective)
    render = None
    if not args.no_render:
        RH = RW = 800
        chunk = 65536
        n_chunks = 4
        rfocal = syn.focal_from_fov(RW)
        pose_r = torch.from_numpy(syn.orbit_poses(8)[rank % 8]).to(dev)[None].contiguous()

        def render_chunk(c):
            ro, rd, _ = ops.gen_rays(pose_r, RH, RW, rfocal, first_id=(c * chunk) % (RH * RW - chunk), n_rays=chunk)
            return hp.render(ro, rd)
        render_chunk(0)
        barrier()
        ops.profile_enable(True)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for c in range(n_chunks):
            render_chunk(c + 1)
        r1.record()
        barrier()
        ms_r = r0.elapsed_time(r1)
        rprof = ops.profile_read()
        ops.profile_enable(False)
        t = torch.tensor([ms_r], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_r = t.item()
        pk = peaks()
        fwd_ms, fwd_n = rprof.get("mlp_fwd", (0.0, 1))
        render = {"value": world * n_chunks * chunk / (ms_r / 1e3) / 1e6, "unit": "Mrays/s",
                  "chunk_rays": chunk, "chunks": n_chunks, "frame": f"{RH}x{RW}",
                  "mlp_fwd_tflops": (n_chunks * chunk * (N_COARSE * F_DENSITY + (N_COARSE + N_FINE) * F_FWD))
                  / (fwd_ms / 1e3) / 1e12 if fwd_ms > 0 else None,
                  "mlp_fwd_frac_of_bf16_peak": ((n_chunks * chunk * (N_COARSE * F_DENSITY + (N_COARSE + N_FINE) * F_FWD))
                                                / (fwd_ms / 1e3) / 1e12 / pk["tf_sus"]) if fwd_ms > 0 else None,
                  "flops_per_ray": N_COARSE * F_DENSITY + (N_COARSE + N_FINE) * F_FWD}

    # ---- compositing kernel at render scale (HBM roofline of kernel (4); SURVEY.md §7 "hard parts":
    #      at training sizes it is launch-latency bound and L2 resident)
    comp = None
    if rank == 0 and not args.no_micro:
        Rc, Sc_ = 262144, N_COARSE + N_FINE
        gcomp = torch.Generator(device=dev).manual_seed(0)
        raw = torch.rand(Rc, Sc_, 4, device=dev, generator=gcomp)
        e_ = torch.sort(2 + 4 * torch.rand(Rc, Sc_ + 1, device=dev, generator=gcomp), -1).values
        ts_, te_ = e_[:, :-1].contiguous(), e_[:, 1:].contiguous()
        d_rgb_ = torch.rand(Rc, 3, device=dev, generator=gcomp)
        bk_ = torch.ones(3, device=dev)
        for _ in range(3):
            ops.composite_forward(raw, ts_, te_, bkgd=bk_)
            ops.composite_backward(raw, ts_, te_, d_rgb_, bkgd=bk_)
        torch.cuda.synchronize()
        ops.profile_enable(True)
        for _ in range(10):
            ops.composite_forward(raw, ts_, te_, bkgd=bk_)   # 1.4 GB in + out per launch >> L2
            ops.composite_backward(raw, ts_, te_, d_rgb_, bkgd=bk_)
        cprof = ops.profile_read()
        ops.profile_enable(False)
        pk_ = peaks()
        f_ms, b_ms = cprof["composite_fwd"][0] / 10, cprof["composite_bwd"][0] / 10
        f_bytes, b_bytes = Rc * (28 * Sc_ + 20), Rc * (40 * Sc_ + 40)  # bwd recomputes weights: 24 B in + 16 B out / sample
        comp = {"rays": Rc, "samples_per_ray": Sc_, "bound": "hbm", "peak": pk_["hbm"], "unit": "GB/s",
                "fwd": {"ms": f_ms, "achieved": f_bytes / f_ms / 1e6, "frac": f_bytes / f_ms / 1e6 / pk_["hbm"]},
                "bwd": {"ms": b_ms, "achieved": b_bytes / b_ms / 1e6, "frac": b_bytes / b_ms / 1e6 / pk_["hbm"]}}
        del raw, e_, ts_, te_
        # sampling kernels (1) at the same scale, against the same HBM roof (algorithmic bytes of SURVEY §8d)
        us_ = torch.rand(Rc, N_COARSE, device=dev, generator=gcomp)
        up_ = torch.rand(Rc, N_FINE, device=dev, generator=gcomp)
        w_ = torch.rand(Rc, N_COARSE, device=dev, generator=gcomp) ** 4
        for _ in range(2):
            tsc_, _ = ops.sample_stratified(Rc, N_COARSE, NEAR, FAR, us_)
            ops.sample_pdf(tsc_, w_, N_FINE, FAR, up_, want_aux=False)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        for _ in range(10):
            tsc_, _ = ops.sample_stratified(Rc, N_COARSE, NEAR, FAR, us_)
        ev[1].record()
        for _ in range(10):
            ops.sample_pdf(tsc_, w_, N_FINE, FAR, up_, want_aux=False)       # training: random u
        ev[2].record()
        for _ in range(10):
            ops.sample_pdf(tsc_, w_, N_FINE, FAR, None, want_aux=False)      # rendering: deterministic u
        ev[3].record()
        torch.cuda.synchronize()
        st_b = Rc * (24 + 12 * N_COARSE)
        pdf_b = Rc * (4 * (2 * N_COARSE - 3) + 4 * N_FINE + 8 * (N_COARSE + N_FINE))  # weights, bins, u in; merged intervals out
        def _roof(b, ms_):
            return {"ms": ms_, "achieved": b / ms_ / 1e6, "frac": b / ms_ / 1e6 / pk_["hbm"]}
        comp["sampling"] = {"stratified": _roof(st_b, ev[0].elapsed_time(ev[1]) / 10),
                            "sample_pdf_random_u": _roof(pdf_b, ev[1].elapsed_time(ev[2]) / 10),
                            "sample_pdf_deterministic_u": _roof(pdf_b - Rc * 4 * N_FINE, ev[2].elapsed_time(ev[3]) / 10)}
        del us_, up_, w_

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event timing inside the timed region)
    pk = peaks()
    P_c, P_f = R_PER_GPU * N_COARSE, R_PER_GPU * (N_COARSE + N_FINE)
    flops = {  # algorithmic FLOPs per step (both networks) by kernel
        "mlp_fwd_train": (P_c + P_f) * F_FWD,
        # backward = F_TRAIN - F_FWD, split dgrad (input grads of 9 GEMMs) / wgrad (weight grads):
        "mlp_dgrad": (P_c + P_f) * (F_TRAIN - F_FWD - F_FWD),
        "mlp_wgrad": (P_c + P_f) * F_FWD,
    }
    # algorithmic HBM bytes per step of the same kernels under the stash layout: per
    # 128-sample tile the forward writes every GEMM-input image once (640 KB) + the 1-bit masks (34 KB),
    # dgrad writes the d(pre-activation) images (608 KB) and reads the masks, wgrad reads each dpre image
    # and each input image once (608 KB + 608 KB; the re-reads of the skip / branch images are NOT counted)
    tiles = (P_c + P_f) / 128
    KB = 1024
    algo_bytes = {"mlp_fwd_train": tiles * (640 + 34) * KB, "mlp_dgrad": tiles * (608 + 34) * KB,
                  "mlp_wgrad": tiles * (608 + 608) * KB}
    share = {k: v[0] / ms for k, v in prof.items()}
    dom = max((k for k in prof if k in flops), key=lambda k: prof[k][0])
    dom_ms_per_step = prof[dom][0] / K
    roofs = {}
    for k in flops:
        if k not in prof:
            continue
        t = prof[k][0] / K / 1e3
        tf, gb = flops[k] / t / 1e12, algo_bytes[k] / t / 1e9
        roofs[k] = {"tensor": {"achieved": tf, "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": tf / pk["tf_sus"]},
                    "hbm": {"achieved": gb, "peak": pk["hbm"], "unit": "GB/s", "frac": gb / pk["hbm"],
                            "algorithmic_bytes_per_step": algo_bytes[k]}}
    # the roof that binds the dominant kernel = the one it is closer to
    bound = max(("tensor", "hbm"), key=lambda b: roofs[dom][b]["frac"])
    R_ = roofs[dom][bound]
    traffic, traffic_src = ncu_traffic(dom)
    # DRAM bytes of the committed ncu capture / the live kernel time (includes the re-reads)
    dram = {}
    for k in flops:
        tb, _ = ncu_traffic(k)
        if tb is not None and k in prof:
            gbs = tb * (prof[k][1] / K) / (prof[k][0] / K / 1e3) / 1e9
            dram[k] = {"gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm"]}
    roofline = {"kernel": dom, "bound": bound, "achieved": R_["achieved"], "peak": R_["peak"], "unit": R_["unit"],
                "frac": R_["frac"], "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": pk["src"] + (" (sustained bf16)" if bound == "tensor" else " (copy bandwidth)"),
                "per_kernel": roofs, "dram_from_traffic": dram,
                "launches_per_step": prof[dom][1] / K, "ms_per_step": dom_ms_per_step,
                "kernel_ms_per_step": {k: v[0] / K for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
                "kernel_share_of_step": {k: round(s, 4) for k, s in sorted(share.items(), key=lambda kv: -kv[1])},
                "mlp_all_tflops": sum(flops.values()) / (sum(prof[k][0] for k in flops if k in prof) / K / 1e3) / 1e12}

    base = None
    if not args.no_cpu_baseline:
        base, _ = cpu_baseline(512, steps=16, warmup=1)  # ~10-15 s of CPU work on the box's host cores

    line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e / Ke},
            "gpu_launches": launches, "roofline": roofline, "roofline_compositing": comp, "cpu_baseline": base, "render": render,
            "dropin_loop": dropin,
            "clocks": sampler.summary(), "final_loss": final_loss,
            "mlp_model_flops_per_ray": (N_COARSE + N_COARSE + N_FINE) * F_TRAIN}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
