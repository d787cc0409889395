#!/usr/bin/env python
"""bench.py -- headline benchmark: Instant-NGP training throughput (rays/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c4|c5]

Default workload = BASELINE.json configs[1] (c2): Lego-shaped 800x800x100 views, 8192-ray batch, scale 0.5, HashGrid
T=2^19.  c4 = configs[3] (ScanNet-shaped 624x468 sparse views + LeReS-style depth-prior loss), c5 = configs[4]
(unbounded scale 16, T=2^22, 1920x1080, exp_step_factor 1/256; the 8-GPU data-parallel configuration).

A "step" = one training step of ngp_pl/train.py:144-170: ray generation from (img_idxs, pix_idxs), AABB, marcher,
hash-grid encode + density/colour MLPs, compositing, NeRFLoss (+ depth-prior term), backward, Adam on all parameters,
and the density-grid update every 16 steps.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

CONFIGS = {
    "c2": dict(workload="NeRF-synthetic Lego-shaped 800x800x100 views, 8192-ray batch, scale 0.5, HashGrid L=16 F=2 T=2^19",
               scale=0.5, n_rays=8192, W=800, H=800, n_img=100, log2_T=19, esf=0.0, scene="object", spr=128,
               lambda_depth=0.0, fx=None),
    "c4": dict(workload="ScanNet-shaped 624x468 sparse views (20 cameras inside a room) + LeReS-style shift/scale-invariant "
                        "depth-prior loss, 8192-ray batch, scale 0.5, HashGrid L=16 F=2 T=2^19",
               scale=0.5, n_rays=8192, W=624, H=468, n_img=20, log2_T=19, esf=0.0, scene="room", spr=256,
               lambda_depth=0.1, fx=577.87 * 624 / 640),
    "c5": dict(workload="unbounded scale=16 (6 cascades), HashGrid L=16 F=2 T=2^22, 1920x1080x100 views, exp_step_factor "
                        "1/256, 8192-ray batch per GPU",
               scale=16.0, n_rays=8192, W=1920, H=1080, n_img=100, log2_T=22, esf=1.0 / 256, scene="unbounded", spr=256,
               lambda_depth=0.0, fx=1500.0),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--pretrain", type=int, default=512,
                    help="untimed training steps before warm-up so that occupancy/density reach steady state")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--comm", default=None, choices=["p2p", "nccl"],
                    help="N > 1: gradient/parameter exchange (default p2p = fused kernel over NVLink peer memory)")
    ap.add_argument("--cpu-rays", type=int, default=2048, help="rays per step of the bounded CPU-baseline sample")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip render / quality / API-path / probe extras")
    ap.add_argument("--profile", action="store_true", help="print the per-kernel event timing table to stderr")
    ap.add_argument("--ncu-steps", type=int, default=0,
                    help="profiling mode for `ncu --profile-from-start off`: after pretrain + warm-up run this many eager "
                         "steps and the memory probes inside cudaProfilerStart/Stop, write gpurun_out/ncu_units.json, exit")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- workloads
class Workload:
    """Synthetic cameras + analytic scene of one BASELINE configuration (datasets are unavailable offline)."""

    def __init__(self, key):
        from google_nerf_b200 import synthetic as syn
        self.key, self.cfg, self.syn = key, CONFIGS[key], syn
        c = self.cfg
        self.K = syn.intrinsics(c["W"], c["H"], fx=c["fx"])
        self.dirs = syn.directions(c["W"], c["H"], self.K)
        if c["scene"] == "object":
            self.poses = syn.hemisphere_poses(c["n_img"])
        elif c["scene"] == "room":
            self.poses = syn.room_poses(c["n_img"])
        else:
            self.poses = syn.ring_poses(c["n_img"])
        self.sc = {"room": syn.ROOM, "unbounded": syn.UNBOUNDED}.get(c["scene"])

    def shade(self, ro, rd, gen=None):
        """-> ground-truth colour, prior disparity (or None): the prior is the true disparity under an affine map plus
        noise (what an image-based monocular depth network delivers), missing on 5 % of the rays."""
        if self.sc is None:
            return self.syn.shade(ro, rd, self.cfg["scale"]), None
        col, t = self.syn.scene_shade(ro, rd, self.sc)
        if self.cfg["lambda_depth"] <= 0:
            return col, None
        disp = torch.where(torch.isfinite(t), 1.0 / t.clamp(min=1e-3), torch.zeros_like(t))
        noise = torch.randn(disp.shape, generator=gen, device="cpu").to(disp.device) if gen is not None else torch.randn_like(disp)
        drop = (torch.rand(disp.shape, generator=gen, device="cpu").to(disp.device) if gen is not None else torch.rand_like(disp)) < 0.05
        prior = (0.6 * disp + 0.35 + 0.02 * noise).clamp(min=1e-3)
        return col, torch.where((disp > 0) & ~drop, prior, torch.zeros_like(prior))


# ---------------------------------------------------------------------------------------------- CPU baseline
def cpu_baseline(key, n_rays, steps, warmup):
    """The same training step as torch CPU ops (oracle/ngp_ref.py) on a bounded sample of the workload."""
    from oracle import ngp_ref as O
    wl = Workload(key); c, syn = wl.cfg, wl.syn
    torch.set_num_threads(os.cpu_count())
    g = torch.Generator().manual_seed(0)
    ref = O.NGPRef(c["scale"], log2_T=c["log2_T"], seed=1337)
    if wl.sc is None:
        grid = syn.density_grid(c["scale"], ref.cascades)
    else:                                                     # analytic occupancy of the world-unit scenes
        coords = syn.grid_coords(128); idx = syn._morton(coords)
        grid = torch.zeros(ref.cascades, 128 ** 3)
        for cc in range(ref.cascades):
            s = min(2 ** (cc - 1), c["scale"])
            grid[cc, idx] = syn.scene_inside((coords.float() / 127 * 2 - 1) * (s - s / 128), wl.sc).float() * 10
    ref.density_bitfield = syn.bitfield_from_grid(grid)
    opt = O.AdamRef([ref.xyz_params, ref.rgb_params], lr=1e-2, eps=1e-15)
    n_img, n_pix = c["n_img"], c["W"] * c["H"]
    times = []
    for it in range(warmup + steps):
        ii = torch.randint(n_img, (n_rays,), generator=g); pi = torch.randint(n_pix, (n_rays,), generator=g)
        t0 = time.perf_counter()
        rays_o, rays_d = syn.get_rays(wl.dirs[pi], wl.poses[ii])
        target, prior = wl.shade(rays_o, rays_d, g)
        noise = torch.rand(n_rays, generator=g)
        if prior is None:
            O.train_step(ref, opt, rays_o, rays_d, target, noise, exp_step_factor=c["esf"])
        else:
            res = O.render(ref, rays_o, rays_d.clone(), noise=noise, exp_step_factor=c["esf"])
            (O.nerf_loss(res, target) + O.depth_prior_loss(res, prior, c["lambda_depth"])).backward()
            opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.median(times))
    return dict(value=n_rays / sec, unit="rays/s", cores=torch.get_num_threads(), kind="port",
                protocol="full training steps (forward, backward, Adam) over a bounded ray sample; BASELINE.md section 3 "
                         "describes a 4096-ray forward render instead -- rays/s of a training step is the metric here",
                sample=f"{steps} steps x {n_rays} rays of the same workload ({key}; oracle/ngp_ref.py train_step, torch CPU "
                       f"ops; median step {sec:.2f} s; analytic occupancy, random-init weights)"), sec


def run_reference(args, rank):
    if rank != 0:
        return
    # exactly K timed steps after W warm-up steps; each step is a bounded sample of the workload (a CPU step over 2048
    # rays takes ~0.65 s on 16 cores -- the same sample size as the main arm's cpu_baseline), shrunk further when K + W
    # is large so that the arm ends within a few minutes
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    n_rays = args.cpu_rays if steps + warmup <= 300 else max(128, int(args.cpu_rays * 300 / (steps + warmup)) // 128 * 128)
    cb, sec = cpu_baseline(args.config, n_rays, steps, warmup)
    line = dict(impl="reference", metric="train_rays_per_s", value=cb["value"], unit="rays/s", n_gpus=args.gpus,
                steps=steps, warmup=warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f16", data="synthetic",
                config=dict(workload=CONFIGS[args.config]["workload"], cpu_sample_rays_per_step=n_rays,
                            note="rays/s normalises the sample size: the CPU arm steps a bounded sample of the batch"),
                cpu_baseline=cb, e2e=dict(value=cb["value"], unit="rays/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="reference vren/tiny-cuda-nn kernels are CUDA-only and absent from the reference tree; this arm "
                     "times the same render/train math as torch CPU ops on the host cores (BASELINE.json north_star)")
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread every ~2 ms (the
    timed region is tens of milliseconds, too short for an `nvidia-smi -lms` child to start), nvidia-smi as fallback."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.sm, self.bits, self.max_mhz, self.err = index, [], 0, None, None
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons",
                                  getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons", None))

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        if get_reasons is not None:
                            self.bits |= int(get_reasons(h))
                    except Exception as e:                   # keep sampling; report the last error
                        self.err = repr(e)
                    time.sleep(0.002)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception as e:
            self.err = repr(e)

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if not self.sm:                                      # NVML unavailable: one nvidia-smi query after the fact
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=20).stdout
                a, b = [float(v) for v in out.strip().split(",")[:2]]
                return dict(sm_mhz=a, sm_max_mhz=b, reasons=[], samples=0, note="NVML unavailable (%s); sampled after the "
                            "timed region" % self.err)
            except Exception:
                return dict(sm_mhz=None, sm_max_mhz=None, reasons=["clock sampling unavailable: %s" % self.err], samples=0)
        reasons = [name for bit, name in self.REASONS if self.bits & bit]
        return dict(sm_mhz=float(np.median(self.sm)), sm_max_mhz=self.max_mhz, reasons=reasons, samples=len(self.sm))


# ---------------------------------------------------------------------------------------------- kernel table
# algorithmic bytes per unit and the roofline that bounds each kernel
def kernel_costs(n_rays, n_samples, n_alive, shard, table_in_l2, world, grad16):
    s, r, a = n_samples, n_rays, n_alive
    gather = "l2" if table_in_l2 else "hbm"                   # 21.8 MiB fp16 table is L2-resident, 185 MiB (T=2^22) is not
    wire = (world - 1) * ((2 if grad16 else 4) + 2)           # NVLink bytes per owned parameter: W-1 gradient reads + W-1 fp16 stores
    return {
        "b2n_rays_from_indices": ("hbm", 52 * r), "b2n_ray_aabb_intersect": ("hbm", 32 * r), "b2n_clamp_near": ("hbm", 8 * r),
        "b2n_raymarching_train_count": ("hbm", 36 * r + 24 * r),
        "b2n_raymarching_train_write": ("hbm", 60 * r + 32 * s),
        "b2n_hashgrid_fw": (gather, 588 * s),
        "b2n_hashgrid_bw": (gather, 1100 * a),
        "b2n_frequency_fw": ("hbm", 172 * s),
        # fused tcgen05 field MLPs: 20.5 / 61 kFLOP per sample (the backward pass recomputes the hidden layers) are ~1% of
        # the tensor roofline; their algorithmic traffic is enc 64 + dirs 12 + sigma 4 + rgb 12 + h 32 = 124 B/sample
        # forward, enc 64 + h 32 + dirs 12 + rgb 12 + dL 16 + index 4 in and dL/denc 64 out = 204 B/alive sample backward
        "b2n_field_mlp_fw": ("hbm", 124 * s),
        "b2n_field_mlp_bw": ("hbm", 204 * a),
        "b2n_field_pack_weights": ("hbm", 40960),
        "b2n_composite_loss_fwbw": ("hbm", 40 * s + 76 * r),
        "b2n_composite_train_fw": ("hbm", 24 * s + 48 * r), "b2n_composite_train_bw": ("hbm", 40 * s + 96 * r),
        "b2n_nerf_loss_fwbw": ("hbm", 56 * r), "b2n_ssi_depth_loss_fwbw": ("hbm", 12 * r),
        "b2n_adam_step": ("hbm", 34 * shard),                 # fp32 p,g,m,v read + p,g(zero),m,v written + fp16 copy
        "b2n_adam_step_peer": ("nvlink", wire * shard),
        "b2n_grad_pack_half": ("hbm", 10 * shard * world),
    }


def nvlink_counters(index):
    """Data bytes sent / received over all NVLinks of one GPU so far (`nvidia-smi nvlink -gt d`), or None."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True,
                             timeout=20).stdout
        tx = sum(int(l.split(":")[-1].split()[0]) for l in out.splitlines() if "Data Tx" in l)
        rx = sum(int(l.split(":")[-1].split()[0]) for l in out.splitlines() if "Data Rx" in l)
        return (tx * 1024, rx * 1024) if (tx or rx) else None
    except Exception:
        return None


def load_ncu_metrics():
    """Per-kernel counters of the committed ncu capture (profiles/summarise.py writes the file from the .ncu-rep of
    profiles/run_ncu.sh): DRAM bytes, L2 sectors / requests and L1 hit rate per processed unit.  bench.py only scales
    them to this run's unit counts; nothing here is a constant typed in by hand."""
    path = os.path.join(ROOT, "profiles", "ncu_metrics.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def profile_kernels(tr, reps=5):
    """CUDA-event duration of every libb2n launch inside real (eager) training steps.  world > 1: called by every rank
    together (the exchange has barriers); the table of the calling rank is returned."""
    from google_nerf_b200 import _lib as L
    orig = L.call
    rec = []

    def timed(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(name, *a); e1.record()
        rec.append((name, e0, e1))

    was = tr.use_graph
    tr.use_graph = False
    L.call = timed
    try:
        for _ in range(reps):
            sset = tr.sets[tr.cur]
            tr.step_count += 1; tr._set_hyper(); tr._march(sset)
            if tr.comm == "nccl" and not tr.comm_in_graph:
                tr._forward_backward(sset); tr._reduce_grads(); tr._optimizer(); tr._gather_params()
            else:
                tr._train(tr.cur)
        torch.cuda.synchronize()
    finally:
        L.call = orig
        tr.use_graph = was
    out = {}
    for name, e0, e1 in rec:
        d = out.setdefault(name, [0.0, 0])
        d[0] += e0.elapsed_time(e1); d[1] += 1
    # per step: total time of a kernel name (a name may launch more than once per step, e.g. the two peer barriers)
    return {k: v[0] / reps for k, v in out.items()}, len(rec) // reps


def api_path_rate(wl, dev, steps=24):
    """The reference's own call path (train.py:144-170) on this repo's kernels: render() -> NeRFLoss -> backward ->
    FusedAdam.step under autograd, eager launches, batches of the headline size.  rays/s over `steps` timed steps."""
    sys.path.insert(0, os.path.join(ROOT, "google-nerf_b200", "shims"))
    from apex.optimizers import FusedAdam
    from google_nerf_b200.losses import NeRFLoss
    from google_nerf_b200.models.networks import NGP
    from google_nerf_b200.models.rendering import render
    c, syn = wl.cfg, wl.syn
    model = NGP(c["scale"], log2_T=c["log2_T"]).to(dev).init_grid_buffers()
    model.mark_invisible_cells(wl.K.to(dev), wl.poses.to(dev), (c["W"], c["H"]))
    opt = FusedAdam(model.parameters(), 1e-2, eps=1e-15)
    loss_fn = NeRFLoss()
    dd, pp = wl.dirs.to(dev), wl.poses.to(dev)
    g = torch.Generator(device=dev).manual_seed(5)
    kw = {"exp_step_factor": c["esf"]} if c["esf"] else {}
    # the dataloader's part, done ahead of the timed loop: batch indices and their ground-truth colours on the GPU
    n_total = 40 + steps
    img = torch.randint(c["n_img"], (n_total, c["n_rays"]), device=dev, generator=g)
    pix = torch.randint(c["W"] * c["H"], (n_total, c["n_rays"]), device=dev, generator=g)
    tgt_all = torch.stack([wl.shade(*syn.get_rays(dd[pix[k]], pp[img[k]]))[0] for k in range(n_total)])

    def step(k):
        if k % 16 == 0:
            model.update_density_grid(0.01 * 1024 / 3 ** 0.5, warmup=k < 256)
        ro, rd = syn.get_rays(dd[pix[k]], pp[img[k]])           # train.py:150-157
        res = render(model, ro, rd, **kw)
        loss = sum(v.mean() for v in loss_fn(res, {"rgb": tgt_all[k]}).values())
        opt.zero_grad(); loss.backward(); opt.step()
        return loss

    for k in range(40):                                       # untimed: occupancy warm-up, allocator steady state
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(40, 40 + steps):
        last = step(k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return dict(rays_per_s=c["n_rays"] / (ms * 1e-3), ms_per_step=ms, last_loss=float(last.item()),
                note="render() + NeRFLoss + loss.backward() + FusedAdam.step() (train.py:144-170) through the drop-in "
                     "modules: one fused autograd node (hash gather + tcgen05 field kernels) behind NGP.forward; eager "
                     "launches, incl. get_rays and the grid update every 16 steps; batches (indices + colours) prepared "
                     "ahead like a DataLoader would")


# ---------------------------------------------------------------------------------------------- main arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch.distributed as dist
    from google_nerf_b200 import _lib as LL
    from google_nerf_b200.models.networks import NGP
    from google_nerf_b200.trainer import NGPTrainer

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = Workload(args.config); cfg, syn = wl.cfg, wl.syn
    SCALE, N_RAYS, W_IMG, H_IMG, N_IMG = cfg["scale"], cfg["n_rays"], cfg["W"], cfg["H"], cfg["n_img"]
    # ---- synthetic dataset (same on every rank), model replicated with identical seeds
    torch.manual_seed(1337)
    K, dirs, poses = wl.K, wl.dirs, wl.poses
    model = NGP(SCALE, encoding="HashGrid", log2_T=cfg["log2_T"]).to(dev)
    tr = NGPTrainer(model, n_rays=N_RAYS, use_graph=not args.no_graph, seed=1234, samples_per_ray=cfg["spr"],
                    comm=args.comm, exp_step_factor=cfg["esf"], lambda_depth=cfg["lambda_depth"])
    tr.set_dataset(dirs, poses)
    model.mark_invisible_cells(K.to(dev), poses.to(dev), (W_IMG, H_IMG))

    total_steps = args.pretrain + args.warmup + 2 * args.steps + 8
    g = torch.Generator().manual_seed(100 + rank)
    img_all = torch.randint(N_IMG, (total_steps, N_RAYS), generator=g)
    pix_all = torch.randint(W_IMG * H_IMG, (total_steps, N_RAYS), generator=g)
    # ground truth of every batch, shaded on the GPU in chunks, kept in pinned host memory (the dataloader's
    # `rays[img_idxs, pix_idxs]`, datasets/base.py:31); the depth prior travels the same way
    rgb_all = torch.empty(total_steps, N_RAYS, 3).pin_memory()
    disp_all = torch.empty(total_steps, N_RAYS).pin_memory() if cfg["lambda_depth"] > 0 else None
    dd, pp = dirs.to(dev), poses.to(dev)
    for s0 in range(0, total_steps, 64):
        ii = img_all[s0:s0 + 64].reshape(-1).to(dev); pi = pix_all[s0:s0 + 64].reshape(-1).to(dev)
        ro, rd = syn.get_rays(dd[pi], pp[ii])
        col, prior = wl.shade(ro, rd)
        rgb_all[s0:s0 + 64] = col.view(-1, N_RAYS, 3).cpu()
        if disp_all is not None:
            disp_all[s0:s0 + 64] = prior.view(-1, N_RAYS).cpu()
    img_all, pix_all = img_all.pin_memory(), pix_all.pin_memory()
    img_dev, pix_dev, rgb_dev = img_all.to(dev), pix_all.to(dev), rgb_all.to(dev)
    disp_dev = disp_all.to(dev) if disp_all is not None else None

    def dev_batch(i):
        b = {"img_idxs": img_dev[i], "pix_idxs": pix_dev[i], "rgb": rgb_dev[i]}
        if disp_dev is not None:
            b["disp"] = disp_dev[i]
        return b

    def host_batch(i):
        b = {"img_idxs": img_all[i], "pix_idxs": pix_all[i], "rgb": rgb_all[i]}
        if disp_all is not None:
            b["disp"] = disp_all[i]
        return b
    it = [0]
    primed = [False]

    def _step(make):
        # like a prefetching DataLoader (train.py:126-131) the next batch is handed over with the current step, so
        # its rays are generated and marched while this step's backward runs
        i = it[0]; it[0] += 1
        cur = None if primed[0] else make(i)
        primed[0] = True
        return tr.step_batch(cur, next_batch=make(i + 1))

    def dev_step():                       # inputs already resident in HBM
        return _step(dev_batch)

    # the call a user makes: host (pinned) batch in, loss out.  The loss of every step is copied to pinned host memory
    # and read by the host, one step behind the GPU (the read of step k happens after step k+1 has been queued, like a
    # logger that prints the previous step's loss), so that the D2H read does not drain the GPU between steps.
    loss_host = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def host_step(k):
        loss_dev = _step(host_batch)
        loss_host[k & 1].copy_(loss_dev, non_blocking=True)
        loss_ev[k & 1].record()
        if k == 0:
            return None
        loss_ev[(k - 1) & 1].synchronize()
        return float(loss_host[(k - 1) & 1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- pretrain (untimed) to steady-state occupancy; grow the sample capacity if it overflows
    def overflowed():                     # collective: every rank re-captures its graphs together
        f = torch.tensor([int(tr.overflowed())], device=dev)
        if world > 1:
            dist.all_reduce(f, op=dist.ReduceOp.MAX)
        return bool(f.item())

    for i in range(args.pretrain):
        dev_step()
        if i % 32 == 31 and overflowed():
            tr.grow(1.5); primed[0] = False
    for _ in range(max(args.warmup, 3)):
        dev_step()
    barrier()
    if overflowed():
        tr.grow(1.5); primed[0] = False
        for _ in range(3):
            dev_step()
        barrier()

    # ---- profiling mode (run under `ncu --profile-from-start off`): eager steps + the memory probes, nothing timed
    if args.ncu_steps > 0:
        assert world == 1
        tr.use_graph = False
        sset = tr.sets[tr.cur]
        probe = torch.zeros(32 << 20, dtype=torch.uint8, device=dev); sink = torch.zeros(4, dtype=torch.int32, device=dev)
        import ctypes
        nl = ctypes.c_int64(0)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(args.ncu_steps):
            tr.step_count += 1; tr._set_hyper(); tr._march(sset); tr._train(tr.cur)
        LL.call("b2n_membench_gather", LL.ptr(probe), 16 << 20, 16, LL.ptr(sink), ctypes.byref(nl))
        LL.call("b2n_membench_read", LL.ptr(probe), 32 << 20, 20, LL.ptr(sink))
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        units = dict(config=args.config, rays=N_RAYS, samples=int(sset.counter[0].item()), alive=int(tr.alive_cnt.item()),
                     params=tr.shard, gather_probe_loads=nl.value, read_probe_bytes=(32 << 20) * 20, steps=args.ncu_steps)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "ncu_units.json"), "w") as f:
            json.dump(units, f)
        print(json.dumps(units), flush=True)
        return

    # ---- timed: device-resident inputs
    clocks = ClockSampler(local); clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nv0 = nvlink_counters(local) if (world > 1 and rank == 0) else None
    barrier(); e0.record()
    for _ in range(args.steps):
        dev_step()
    e1.record(); barrier()
    nv1 = nvlink_counters(local) if nv0 is not None else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    # ---- timed: end to end through the public API with host batches (H2D + loss D2H inside)
    barrier(); t0 = time.perf_counter(); e0.record()
    for k in range(args.steps):
        last_loss = host_step(k)
    loss_ev[(args.steps - 1) & 1].synchronize()
    last_loss = float(loss_host[(args.steps - 1) & 1])        # the last step's loss, still inside the timed region
    e1.record(); barrier()
    ms_e2e = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)], device=dev)
    clk = clocks.stop()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(ms.item()), float(ms_e2e.item())
    if tr.peer is not None:
        tr.peer.check()                   # a timed-out peer barrier would have invalidated the run
    skipped, loss_scale_now = tr.skipped_steps()

    # ---- every rank: per-kernel event table of eager steps (the exchange kernels have barriers: collective)
    table, launches_per_step = profile_kernels(tr)
    samples = int(tr.sets[tr.cur].counter[0].item())          # sample / alive-sample counts of the profiled batch
    alive = int(tr.alive_cnt.item())
    # ---- every rank: gather the model (sharded optimiser) and time the occupancy-grid update (it max-reduces)
    tr.sync_model()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        tr.update_density_grid(warmup=False)
    torch.cuda.synchronize(); grid_update_ms = (time.perf_counter() - t0) / 3 * 1e3
    extras = not args.no_extras
    # ---- every rank: test-time render of full frames (BASELINE.json configs[2]), sharded over the ranks in
    # round-robin row tiles (dist_utils.render_sharded); the time is the max over ranks and includes the all-gather
    from google_nerf_b200.models.rendering import render
    from google_nerf_b200.dist_utils import render_sharded
    render_info = None
    if extras:
        def frame_rate(**more):
            # 2 warm-up + 5 timed frames of the SAME view (frames of different views differ in cost); the rays exist before
            # the clock starts; median of the timed frames, max over ranks
            frames, samples = [], 0
            rkw = dict(test_time=True, T_threshold=1e-2, exp_step_factor=cfg["esf"], **more)
            with torch.no_grad():
                ro, rd = syn.get_rays(dd, pp[0])
                for f in range(7):
                    torch.cuda.synchronize(); barrier(); t0 = time.perf_counter()
                    res = render_sharded(lambda o, d, **kw: render(model, o, d, **kw), ro, rd, tile=W_IMG, **rkw)
                    torch.cuda.synchronize(); dt = torch.tensor([time.perf_counter() - t0], device=dev)
                    if world > 1:
                        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                    frames.append(float(dt.item())); samples = int(res["total_samples"])
            t = sorted(frames[2:])[len(frames[2:]) // 2]
            return dict(mrays_per_s=W_IMG * H_IMG / t / 1e6, ms_per_frame=t * 1e3, ms_per_frame_min=min(frames[2:]) * 1e3,
                        samples_per_ray=samples / (W_IMG * H_IMG))
        fr = frame_rate(); rl = frame_rate(whole_rays=False)
        render_info = dict(**fr, rays=W_IMG * H_IMG, n_gpus=world, round_loop=rl,
                           note="render(test_time=True), T_threshold 1e-2 as in test.ipynb.  Default path: whole rays in "
                                "one persistent kernel (csrc/render_tc.cu) while the fp16 hash table fits the L2 (c2, "
                                "c4), else the round loop (c5: both entries are the round loop); round_loop = the loop of "
                                "rendering.py:42-114 driven from the device (whole_rays=False).  N > 1: row tiles dealt "
                                "round-robin to the ranks, max over ranks, all-gather of rgb/depth/opacity included")
        whole = getattr(model, "_whole_rays", None)
        render_info["path"] = "whole_rays" if whole is not None else "round_loop"
        if whole is not None and world == 1:
            render_info["fell_back_to_round_loop"] = bool(int(whole.ctl_host[1]) > 0)
    # ---- image quality of what was just trained (sanity of the whole path, not a timed number): PSNR of a training view
    # and of held-out views against the analytic ground truth (T_threshold 1e-4 like validation, train.py:178-183)
    quality = None
    if rank == 0 and extras:
        from google_nerf_b200.metrics import psnr
        with torch.no_grad():
            held = {"object": lambda: syn.hemisphere_poses(4, seed=12345), "room": lambda: syn.room_poses(4, seed=12345),
                    "unbounded": lambda: syn.ring_poses(4, seed=12345)}[cfg["scene"]]().to(dev)
            vals = []
            for pose in (pp[0], held[0], held[1]):
                ro, rd = syn.get_rays(dd, pose)
                img = render(model, ro, rd, test_time=True, exp_step_factor=cfg["esf"])["rgb"]
                vals.append(float(psnr(img, wl.shade(ro, rd)[0])))
        quality = dict(psnr_train_view=vals[0], psnr_heldout_views=vals[1:], train_steps=tr.step_count,
                       note=f"{W_IMG}x{H_IMG} frames against the analytic scene's closed-form shading")
    api = None
    if rank == 0 and extras and args.config == "c2":
        api = api_path_rate(wl, dev)
    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
        hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
        tf_peak = peaks["bf16_tflops_sustained"] if peaks else 1400.0
        peak_src = "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6.65 TB/s / 1.4 PFLOP/s (of fallback)"

        # ---- L2 / HBM read probes: the roofline of the gather kernels (MEASURED_PEAKS.json has no L2 figure)
        def membench(nbytes, iters):
            buf = torch.empty(nbytes // 4, dtype=torch.int32, device=dev).zero_(); sink = torch.zeros(1, dtype=torch.int32, device=dev)
            LL.call("b2n_membench_read", LL.ptr(buf), nbytes, 2, LL.ptr(sink))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); LL.call("b2n_membench_read", LL.ptr(buf), nbytes, iters, LL.ptr(sink)); b.record(); torch.cuda.synchronize()
            return nbytes * iters / (a.elapsed_time(b) * 1e-3) / 1e9
        l2_gbs = membench(32 << 20, 200)

        def gatherbench(nbytes, iters):
            import ctypes
            buf = torch.empty(nbytes, dtype=torch.uint8, device=dev); sink = torch.zeros(4, dtype=torch.int32, device=dev)
            nl = ctypes.c_int64(0)
            LL.call("b2n_membench_gather", LL.ptr(buf), nbytes, 4, LL.ptr(sink), ctypes.byref(nl))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); LL.call("b2n_membench_gather", LL.ptr(buf), nbytes, iters, LL.ptr(sink), ctypes.byref(nl)); b.record()
            torch.cuda.synchronize()
            return nl.value / (a.elapsed_time(b) * 1e-3)                       # 4-byte loads per second
        gather_loads_per_s = gatherbench(16 << 20, 64)       # 16 MiB: about the fp16 table (21.8 MiB), L2-resident

        table_in_l2 = cfg["log2_T"] <= 20
        costs = kernel_costs(N_RAYS, samples, alive, tr.shard, table_in_l2, world, tr.grad_fp16)
        # the dominant kernel of the step's critical path (ray generation / AABB / marching of the NEXT batch run on
        # the side stream underneath it and are reported separately in `marcher`)
        side = ("b2n_raymarching", "b2n_ray_aabb", "b2n_clamp_near", "b2n_rays_from_indices")
        crit = {k: v for k, v in table.items() if not k.startswith(side)}
        top = max(crit, key=crit.get)
        bound, alg = costs.get(top, ("hbm", None))
        dur_s = table[top] * 1e-3
        peak = {"hbm": hbm_peak, "l2": l2_gbs, "nvlink": 770.0}[bound]
        ach = alg / dur_s / 1e9 if alg else None
        ncu = load_ncu_metrics()
        unit_count = {"alive_sample": alive, "sample": samples, "param": tr.shard, "ray": N_RAYS}

        def ncu_per_launch(kernel, field):
            if not ncu or kernel not in ncu.get("kernels", {}) or ncu.get("config") != args.config:
                return None
            e = ncu["kernels"][kernel]
            v, u = e.get(field), unit_count.get(e.get("unit"))
            return v * u if v is not None and u is not None else None
        roofline = dict(kernel=top, bound=bound, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak if ach else None,
                        traffic=ncu_per_launch(top, "dram_bytes_per_unit"),
                        peak_source={"hbm": peak_src, "nvlink": "B200_PROFILING.md: measured 770 GB/s peer copy per direction",
                                     "l2": "b2n_membench_read over a 32 MiB (L2-resident) buffer, measured in this run; "
                                           "MEASURED_PEAKS.json holds no L2 figure"}[bound],
                        ms_per_launch=table[top], share_of_step=table[top] / sum(crit.values()),
                        algorithmic_per_launch=alg,
                        traffic_source=("profiles/ncu_metrics.json (" + ncu.get("round", "?") + "), per-unit DRAM bytes of the "
                                        "committed ncu capture scaled to this run's unit count") if ncu else None)
        if args.profile:
            for k, v in sorted(table.items(), key=lambda kv: -kv[1]):
                print(f"  {k:40s} {v * 1e3:9.1f} us", file=sys.stderr)
        # ---- hash encode vs its rooflines
        t_fw, t_bw = table.get("b2n_hashgrid_fw", 0) * 1e-3, table.get("b2n_hashgrid_bw", 0) * 1e-3
        req_fw = ncu_per_launch("b2n_hashgrid_fw", "lts_requests_per_unit")
        req_bw = ncu_per_launch("b2n_hashgrid_bw", "lts_requests_per_unit")
        sec_fw = ncu_per_launch("b2n_hashgrid_fw", "lts_sectors_per_unit")
        probe_req_per_load = ncu["kernels"]["b2n_membench_gather"]["lts_requests_per_unit"] if ncu and \
            "b2n_membench_gather" in ncu.get("kernels", {}) else None
        hash_encode = dict(
            fw_gbs=588 * samples / t_fw / 1e9 if t_fw else None, bw_gbs=1100 * alive / t_bw / 1e9 if t_bw else None,
            bound="l2" if table_in_l2 else "hbm", l2_read_gbs_measured=l2_gbs, hbm_gbs_measured=hbm_peak,
            fw_frac_of_bound=(588 * samples / t_fw / 1e9) / (l2_gbs if table_in_l2 else hbm_peak) if t_fw else None,
            bw_frac_of_bound=(1100 * alive / t_bw / 1e9) / (l2_gbs if table_in_l2 else hbm_peak) if t_bw else None,
            # what bounds a 4-byte gather on the L1-miss path is the REQUEST rate (one per clock and SM), not bytes:
            # requests per sample come from the committed ncu capture, the achievable rate from the live probe
            fw_l2_requests_per_s=req_fw / t_fw if (req_fw and t_fw) else None,
            fw_l2_sectors_per_s=sec_fw / t_fw if (sec_fw and t_fw) else None,
            probe_loads_per_s=gather_loads_per_s,
            probe_l2_requests_per_s=gather_loads_per_s * probe_req_per_load if probe_req_per_load else None,
            fw_frac_of_request_rate=(req_fw / t_fw) / (gather_loads_per_s * probe_req_per_load)
            if (req_fw and t_fw and probe_req_per_load) else None,
            bw_l2_requests_per_s=req_bw / t_bw if (req_bw and t_bw) else None,
            bw_frac_of_request_rate=(req_bw / t_bw) / (gather_loads_per_s * probe_req_per_load)
            if (req_bw and t_bw and probe_req_per_load) else None,
            note="algorithmic bytes: 588 B/sample fw, 1100 B/sample bw (SURVEY 8d); table %.1f MiB fp16" %
                 (model.xyz_encoder.enc.n_params * 2 / 2 ** 20))
        t_m = (table.get("b2n_raymarching_train_count", 0) + table.get("b2n_raymarching_train_write", 0)) * 1e-3
        marcher = dict(samples_per_s=samples / t_m if t_m else None, rays_per_s=N_RAYS / t_m if t_m else None,
                       gbs=(120 * N_RAYS + 32 * samples) / t_m / 1e9 if t_m else None)
        t_c = table.get("b2n_composite_loss_fwbw", 0) * 1e-3
        compositing = dict(gbs=(40 * samples + 76 * N_RAYS) / t_c / 1e9 if t_c else None,
                           note="fused fw + loss + bw launch; algorithmic 40 B/sample + 76 B/ray; latency-bound at 8192 rays")
        t_f, t_b = table.get("b2n_field_mlp_fw", 0) * 1e-3, table.get("b2n_field_mlp_bw", 0) * 1e-3
        mlp = dict(fw_tflops=20480 * samples / t_f / 1e12 if t_f else None,
                   bw_tflops=61440 * alive / t_b / 1e12 if t_b else None, tensor_peak_tflops=tf_peak,
                   fw_frac_of_tensor_peak=20480 * samples / t_f / 1e12 / tf_peak if t_f else None,
                   bw_frac_of_tensor_peak=61440 * alive / t_b / 1e12 / tf_peak if t_b else None,
                   note="tcgen05 kind::f16; the MLPs are 64 wide: dependency latency bounds them, not the tensor pipe "
                        "(ncu sm__pipe_tensor_cycles_active in profiles/); bw counts the recomputed forward layers")
        exchange = None
        if world > 1:
            wire = costs["b2n_adam_step_peer"][1]
            t_x = table.get("b2n_adam_step_peer", 0) * 1e-3
            exchange = dict(comm=tr.comm, grad_wire="fp16 table / fp32 MLP" if tr.grad_fp16 else "fp32",
                            us={k: round(v * 1e3, 1) for k, v in table.items()
                                if k.startswith(("b2n_peer", "b2n_adam_step", "b2n_grad_pack", "b2n_scaler"))},
                            nvlink_bytes_per_step=wire, nvlink_gbs=wire / t_x / 1e9 if t_x else None,
                            nvlink_peak_gbs=770.0,
                            nvlink_counters_bytes_per_step=dict(tx=(nv1[0] - nv0[0]) / args.steps, rx=(nv1[1] - nv0[1]) / args.steps,
                                                                source="nvidia-smi nvlink -gt d around the timed region, rank 0")
                            if (nv0 and nv1) else "nvidia-smi nvlink -gt d reports N/A for every link on this pool",
                            note="per rank and direction: (W-1) remote gradient reads + (W-1) remote fp16 parameter stores "
                                 "per owned element; the two barrier launches include the wait for the slowest rank")
        cb = None
        if not args.skip_cpu and world == 1:                  # rank 0 at N = 1 only (the other ranks would spin in NCCL)
            cb, _ = cpu_baseline(args.config, args.cpu_rays, 12, 1)   # ~10 s of CPU work on the host cores
        rays = N_RAYS * world * args.steps
        n_updates = sum(1 for s in range(args.steps) if s % tr.S == 0)
        h2d = N_RAYS * (8 + 8 + 12 + (4 if cfg["lambda_depth"] > 0 else 0))
        line = dict(metric="train_rays_per_s", value=rays / (ms * 1e-3), unit="rays/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f16", data="synthetic (analytic scene, random-init weights "
                    f"trained {args.pretrain} untimed steps to steady-state occupancy)",
                    config=dict(workload=cfg["workload"], config=args.config, rays_per_gpu=N_RAYS, samples_per_step=samples,
                                alive_samples_per_step=alive, samples_per_ray=samples / N_RAYS, cuda_graph=not args.no_graph,
                                l2="per-step working set (optimiser state %.0f MB + sample buffers) exceeds the 126 MB L2; "
                                   "no explicit flush" % (18 * tr.shard / 1e6), parallelism=f"dp{world}", comm=tr.comm,
                                loss_scale=loss_scale_now, skipped_steps=skipped,
                                cpu_sample_rays_per_step=args.cpu_rays if cb else None),
                    clocks=clk,
                    e2e=dict(value=rays / (ms_e2e * 1e-3), unit="rays/s", h2d_bytes_per_step=h2d,
                             d2h_bytes_per_step=4, ms_per_step=ms_e2e / args.steps, last_loss=last_loss),
                    gpu_launches=launches_per_step * args.steps + 12 * n_updates,
                    roofline=roofline, cpu_baseline=cb, hash_encode=hash_encode, marcher=marcher, compositing=compositing,
                    mlp=mlp, exchange=exchange, render=render_info, quality=quality, api_path=api,
                    grid_update_ms=grid_update_ms, kernels_us={k: round(v * 1e3, 1) for k, v in table.items()})
        print(json.dumps(line), flush=True)
    tr.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
