#!/usr/bin/env python
"""bench.py -- headline benchmark: Instant-NGP training throughput (rays/s) on the Lego-shaped synthetic
workload (BASELINE.json configs[1]: 800x800x100 views, 8192-ray batch, scale 0.5, HashGrid T=2^19).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" = one training step of ngp_pl/train.py:144-170: ray generation from (img_idxs, pix_idxs), AABB,
marcher, hash-grid encode + density/colour MLPs, compositing, NeRFLoss, backward, Adam on all parameters, and
the density-grid update every 16 steps.  One JSON line is printed by rank 0.
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

SCALE, N_RAYS, W_IMG, H_IMG, N_IMG = 0.5, 8192, 800, 800, 100
WORKLOAD = "NeRF-synthetic Lego-shaped 800x800x100 views, 8192-ray batch, scale 0.5, HashGrid L=16 F=2 T=2^19"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pretrain", type=int, default=512,
                    help="untimed training steps before warm-up so that occupancy/density reach steady state")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--comm", default=None, choices=["p2p", "nccl"],
                    help="N > 1: gradient/parameter exchange (default p2p = fused kernel over NVLink peer memory)")
    ap.add_argument("--cpu-rays", type=int, default=2048, help="rays per step of the bounded CPU-baseline sample")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--profile", action="store_true", help="print the per-kernel event timing table to stderr")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- CPU baseline
def cpu_baseline(n_rays, steps, warmup):
    """The same training step as torch CPU ops (oracle/ngp_ref.py) on a bounded sample of the workload."""
    from google_nerf_b200 import synthetic as syn
    from oracle import ngp_ref as O
    torch.set_num_threads(os.cpu_count())
    g = torch.Generator().manual_seed(0)
    ref = O.NGPRef(SCALE, seed=1337)
    ref.density_bitfield = syn.bitfield_from_grid(syn.density_grid(SCALE, 1))
    K = syn.intrinsics(W_IMG, H_IMG); dirs = syn.directions(W_IMG, H_IMG, K); poses = syn.hemisphere_poses(N_IMG)
    opt = O.AdamRef([ref.xyz_params, ref.rgb_params], lr=1e-2, eps=1e-15)
    times = []
    for it in range(warmup + steps):
        ii = torch.randint(N_IMG, (n_rays,), generator=g); pi = torch.randint(W_IMG * H_IMG, (n_rays,), generator=g)
        t0 = time.perf_counter()
        rays_o, rays_d = syn.get_rays(dirs[pi], poses[ii])
        target = syn.shade(rays_o, rays_d, SCALE)
        O.train_step(ref, opt, rays_o, rays_d, target, torch.rand(n_rays, generator=g))
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.median(times))
    return dict(value=n_rays / sec, unit="rays/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{steps} steps x {n_rays} rays of the same workload (oracle/ngp_ref.py train_step, torch CPU ops; "
                       f"median step {sec:.2f} s; analytic occupancy, random-init weights)"), sec


def run_reference(args, rank):
    if rank != 0:
        return
    # exactly K timed steps after W warm-up steps; each step is a bounded sample of the workload (a CPU step over 2048
    # rays takes ~0.65 s on 16 cores -- the same sample size as the main arm's cpu_baseline), shrunk further when K + W
    # is large so that the arm ends within a few minutes
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    n_rays = args.cpu_rays if steps + warmup <= 300 else max(128, int(args.cpu_rays * 300 / (steps + warmup)) // 128 * 128)
    cb, sec = cpu_baseline(n_rays, steps, warmup)
    line = dict(impl="reference", metric="train_rays_per_s", value=cb["value"], unit="rays/s", n_gpus=args.gpus,
                steps=steps, warmup=warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f16", data="synthetic", config=dict(workload=WORKLOAD),
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
# algorithmic bytes / flops per unit (DESIGN.md "Cost model")
def kernel_costs(n_rays, n_samples, n_alive, n_params_xyz, n_params_rgb):
    # the two backward kernels run over the alive samples only (those composited before their ray's early stop)
    s, r, a = n_samples, n_rays, n_alive
    return {
        "b2n_ray_aabb_intersect": ("hbm", 32 * r),
        "b2n_raymarching_train_count": ("hbm", 36 * r + 24 * r),
        "b2n_raymarching_train_write": ("hbm", 60 * r + 32 * s),
        "b2n_hashgrid_fw": ("hbm", 588 * s),
        "b2n_hashgrid_bw": ("hbm", 1100 * a),
        "b2n_mlp_fw": ("tensor", None),
        "b2n_mlp_bw": ("tensor", None),
        # fused tcgen05 field MLPs: 20.5 / 61 kFLOP per sample (the backward pass recomputes the hidden layers) are ~1% of
        # the tensor roofline; their algorithmic traffic is enc 64 + dirs 12 + sigma 4 + rgb 12 + h 32 = 124 B/sample
        # forward, enc 64 + h 32 + dirs 12 + rgb 12 + dL 16 + index 4 in and dL/denc 64 out = 204 B/alive sample backward
        "b2n_field_mlp_fw": ("hbm", 124 * s),
        "b2n_field_mlp_bw": ("hbm", 204 * a),
        "b2n_field_pack_weights": ("hbm", 40960),
        "b2n_sh4_fw": ("hbm", 44 * s),
        "b2n_composite_loss_fwbw": ("hbm", 40 * s + 76 * r),
        "b2n_adam_step": ("hbm", None),
    }


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
    """CUDA-event duration of every libb2n launch inside real (eager) training steps."""
    from google_nerf_b200 import _lib as L
    orig = L.call
    rec = []

    def timed(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(name, *a); e1.record()
        rec.append((name, a, e0, e1))

    import google_nerf_b200.trainer as T
    L.call = timed
    try:
        for _ in range(reps):
            sset = tr.sets[tr.cur]
            tr.step_count += 1; tr._set_hyper(); tr._march(sset); tr._forward_backward(sset); tr._optimizer()
        torch.cuda.synchronize()
    finally:
        L.call = orig
    out = {}
    for name, a, e0, e1 in rec:
        key = name
        if name == "b2n_mlp_fw" or name == "b2n_mlp_bw":
            key = f"{name}[{'rgb' if a[4 if name == 'b2n_mlp_fw' else 5] == 2 else 'sigma'}]"
        d = out.setdefault(key, [0.0, 0])
        d[0] += e0.elapsed_time(e1); d[1] += 1
    return {k: v[0] / v[1] for k, v in out.items()}, len(rec) // reps


# ---------------------------------------------------------------------------------------------- main arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch.distributed as dist
    from google_nerf_b200 import synthetic as syn
    from google_nerf_b200.models.networks import NGP
    from google_nerf_b200.trainer import NGPTrainer

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic dataset (same on every rank), model replicated with identical seeds
    torch.manual_seed(1337)
    K = syn.intrinsics(W_IMG, H_IMG); dirs = syn.directions(W_IMG, H_IMG, K); poses = syn.hemisphere_poses(N_IMG)
    model = NGP(SCALE, encoding="HashGrid").to(dev)
    tr = NGPTrainer(model, n_rays=N_RAYS, use_graph=not args.no_graph, seed=1234 + rank, samples_per_ray=128,
                    comm=args.comm)
    tr.set_dataset(dirs, poses)
    model.mark_invisible_cells(K.to(dev), poses.to(dev), (W_IMG, H_IMG))

    total_steps = args.pretrain + args.warmup + 2 * args.steps + 8
    g = torch.Generator().manual_seed(100 + rank)
    img_all = torch.randint(N_IMG, (total_steps, N_RAYS), generator=g)
    pix_all = torch.randint(W_IMG * H_IMG, (total_steps, N_RAYS), generator=g)
    # ground truth of every batch, shaded on the GPU in chunks, kept in pinned host memory (the dataloader's
    # `rays[img_idxs, pix_idxs]`, datasets/base.py:31)
    rgb_all = torch.empty(total_steps, N_RAYS, 3).pin_memory()
    dd, pp = dirs.to(dev), poses.to(dev)
    for s0 in range(0, total_steps, 64):
        ii = img_all[s0:s0 + 64].reshape(-1).to(dev); pi = pix_all[s0:s0 + 64].reshape(-1).to(dev)
        ro, rd = syn.get_rays(dd[pi], pp[ii])
        rgb_all[s0:s0 + 64] = syn.shade(ro, rd, SCALE).view(-1, N_RAYS, 3).cpu()
    img_all, pix_all = img_all.pin_memory(), pix_all.pin_memory()
    img_dev, pix_dev, rgb_dev = img_all.to(dev), pix_all.to(dev), rgb_all.to(dev)

    it = [0]
    dev_batch = lambda i: {"img_idxs": img_dev[i], "pix_idxs": pix_dev[i], "rgb": rgb_dev[i]}
    host_batch = lambda i: {"img_idxs": img_all[i], "pix_idxs": pix_all[i], "rgb": rgb_all[i]}
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

    def host_step():                      # the call a user makes: host (pinned) batch in, loss out
        return float(_step(host_batch).item())

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

    # ---- timed: device-resident inputs
    clocks = ClockSampler(local); clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); e0.record()
    for _ in range(args.steps):
        dev_step()
    e1.record(); barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    samples = tr.samples_last_step()
    # ---- timed: end to end through the public API with host batches (H2D + loss D2H inside)
    barrier(); t0 = time.perf_counter(); e0.record()
    for _ in range(args.steps):
        last_loss = host_step()
    e1.record(); barrier()
    ms_e2e = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)], device=dev)
    clk = clocks.stop()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(ms.item()), float(ms_e2e.item())
    if tr.peer is not None:
        tr.peer.check()                   # a timed-out peer barrier would have invalidated the run

    # ---- every rank: gather the model (sharded optimiser) and time the occupancy-grid update (it max-reduces)
    tr.sync_model()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        tr.update_density_grid(warmup=False)
    torch.cuda.synchronize(); grid_update_ms = (time.perf_counter() - t0) / 3 * 1e3
    # ---- every rank: test-time render of full 800x800 frames (BASELINE.json configs[2]), sharded over the ranks in
    # round-robin row tiles (dist_utils.render_sharded); the time is the max over ranks and includes the all-gather
    from google_nerf_b200.models.rendering import render
    from google_nerf_b200.dist_utils import render_sharded
    frames, render_samples = [], 0
    with torch.no_grad():
        for f in range(4):
            ro, rd = syn.get_rays(dd, pp[f % N_IMG])
            barrier(); t0 = time.perf_counter()
            res = render_sharded(lambda o, d, **kw: render(model, o, d, **kw), ro, rd, tile=W_IMG, test_time=True,
                                 T_threshold=1e-2)
            torch.cuda.synchronize(); dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            frames.append(float(dt.item())); render_samples = int(res["total_samples"])
    render_info = dict(mrays_per_s=W_IMG * H_IMG / min(frames[1:]) / 1e6, ms_per_frame=min(frames[1:]) * 1e3,
                       rays=W_IMG * H_IMG, samples_per_ray=render_samples / (W_IMG * H_IMG), n_gpus=world,
                       note="render(test_time=True): the loop of rendering.py:42-114 driven from the device (8 rounds per "
                            "CUDA-graph replay), T_threshold 1e-2 as in test.ipynb; N > 1: row tiles dealt round-robin to "
                            "the ranks, max over ranks, all-gather of rgb/depth/opacity included")
    # ---- image quality of what was just trained (sanity of the whole path, not a timed number): PSNR of a training view
    # and of a held-out view against the analytic ground truth (T_threshold 1e-4 like validation, train.py:178-183)
    quality = None
    if rank == 0:
        from google_nerf_b200.metrics import psnr
        with torch.no_grad():
            novel = syn.hemisphere_poses(4, seed=12345).to(dev)
            vals = []
            for pose in (pp[0], novel[0], novel[1]):
                ro, rd = syn.get_rays(dd, pose)
                img = render(model, ro, rd, test_time=True)["rgb"]
                vals.append(float(psnr(img, syn.shade(ro, rd, SCALE))))
        quality = dict(psnr_train_view=vals[0], psnr_heldout_views=vals[1:], train_steps=tr.step_count,
                       note="800x800 frames against the analytic scene's closed-form shading")
    if rank == 0:
        # ---- per-kernel table + roofline of the dominant kernel (eager replays of the same step)
        tr.use_graph = False
        table, launches_per_step = profile_kernels(tr)
        tr.use_graph = not args.no_graph
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
        hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
        tf_peak = peaks["bf16_tflops_sustained"] if peaks else 1400.0
        samples = int(tr.sets[tr.cur].counter[0].item())      # sample / alive-sample counts of the profiled batch
        alive = int(tr.alive_cnt.item())
        costs = kernel_costs(N_RAYS, samples, alive, tr.p_xyz.numel(), tr.p_rgb.numel())
        # the dominant kernel of the step's critical path (ray generation / AABB / marching of the NEXT batch run on
        # the side stream underneath it and are reported separately in `marcher`)
        side = ("b2n_raymarching", "b2n_ray_aabb", "b2n_clamp_near", "b2n_rays_from_indices")
        top = max((k for k in table if not k.startswith(side)), key=table.get)
        base = top.split("[")[0]
        bound = costs.get(base, ("hbm", None))[0]
        if base == "b2n_adam_step":
            alg = 34.0 * tr.shard                            # fp32 p,g,m,v read + p,g(zero),m,v written + fp16 copy
        elif base == "b2n_mlp_fw":
            alg = samples * (2 * (32 * 64 + 64 * 16) if "sigma" in top else 2 * (32 * 64 + 64 * 64 + 64 * 16))
        elif base == "b2n_mlp_bw":
            alg = samples * 2 * (2 * (32 * 64 + 64 * 16) if "sigma" in top else 2 * (32 * 64 + 64 * 64 + 64 * 16))
        else:
            alg = costs.get(base, ("hbm", 0))[1]
        dur_s = table[top] * 1e-3
        if bound == "hbm":
            ach, peak, unit = alg / dur_s / 1e9, hbm_peak, "GB/s"
        else:
            ach, peak, unit = alg / dur_s / 1e12, tf_peak, "TFLOP/s"
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed capture
        # profiles/r01d_ncu_summary.md (580k samples, 11.43 M parameters), scaled to this run's units
        ncu = load_ncu_metrics()
        traffic = None
        if ncu and base in ncu.get("kernels", {}):
            per_unit = ncu["kernels"][base].get("dram_bytes_per_unit")
            units = {"alive_sample": alive, "sample": samples, "param": tr.shard}.get(ncu["kernels"][base].get("unit"))
            traffic = per_unit * units if per_unit is not None and units is not None else None
        roofline = dict(kernel=top, bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak,
                        traffic=traffic,
                        peak_source="MEASURED_PEAKS.json" if peaks else "fallback",
                        ms_per_launch=table[top], share_of_step=table[top] / sum(table.values()),
                        algorithmic_per_launch=alg)
        if args.profile:
            for k, v in sorted(table.items(), key=lambda kv: -kv[1]):
                print(f"  {k:40s} {v * 1e3:9.1f} us", file=sys.stderr)
        # ---- L2 / HBM read roofline for the gather kernels, marcher and hash-encode rates
        from google_nerf_b200 import _lib as LL
        def membench(nbytes, iters):
            buf = torch.empty(nbytes // 4, dtype=torch.int32, device=dev).zero_(); sink = torch.zeros(1, dtype=torch.int32, device=dev)
            LL.call("b2n_membench_read", LL.ptr(buf), nbytes, 2, LL.ptr(sink))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); LL.call("b2n_membench_read", LL.ptr(buf), nbytes, iters, LL.ptr(sink)); b.record(); torch.cuda.synchronize()
            return nbytes * iters / (a.elapsed_time(b) * 1e-3) / 1e9
        l2_gbs, hbm_read_gbs = membench(32 << 20, 200), membench(2 << 30, 4)

        def gatherbench(nbytes, iters):
            import ctypes
            buf = torch.empty(nbytes, dtype=torch.uint8, device=dev); sink = torch.zeros(4, dtype=torch.int32, device=dev)
            nl = ctypes.c_int64(0)
            LL.call("b2n_membench_gather", LL.ptr(buf), nbytes, 4, LL.ptr(sink), ctypes.byref(nl))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); LL.call("b2n_membench_gather", LL.ptr(buf), nbytes, iters, LL.ptr(sink), ctypes.byref(nl)); b.record()
            torch.cuda.synchronize()
            return nl.value / (a.elapsed_time(b) * 1e-3)                       # sector requests per second
        l2_gather_rate = gatherbench(16 << 20, 64)           # 16 MiB: about the fp16 table (21.8 MiB), L2-resident
        t_fw, t_bw = table.get("b2n_hashgrid_fw", 0) * 1e-3, table.get("b2n_hashgrid_bw", 0) * 1e-3
        hash_encode = dict(fw_gbs=588 * samples / t_fw / 1e9 if t_fw else None, bw_gbs=1100 * alive / t_bw / 1e9 if t_bw else None,
                           l2_read_gbs_measured=l2_gbs, hbm_read_gbs_measured=hbm_read_gbs,
                           fw_frac_of_l2=(588 * samples / t_fw / 1e9) / l2_gbs if t_fw else None,
                           # what actually bounds the gather: 32-byte sector REQUESTS to L2.  58.2 per sample = 128
                           # gathers x (1 - 0.545 L1 hit rate) from the committed ncu capture (profiles/r01d_ncu_summary.md)
                           l2_gather_gsectors_per_s_measured=l2_gather_rate / 1e9,
                           fw_gsectors_per_s=58.2 * samples / t_fw / 1e9 if t_fw else None,
                           fw_frac_of_l2_gather=(58.2 * samples / t_fw) / l2_gather_rate if t_fw else None,
                           note="algorithmic bytes: 588 B/sample fw, 1100 B/sample bw (SURVEY 8d); table 21.8 MiB fp16, L2-resident")
        t_m = (table.get("b2n_raymarching_train_count", 0) + table.get("b2n_raymarching_train_write", 0)) * 1e-3
        marcher = dict(samples_per_s=samples / t_m if t_m else None, rays_per_s=N_RAYS / t_m if t_m else None,
                       gbs=(120 * N_RAYS + 32 * samples) / t_m / 1e9 if t_m else None)
        # ---- SURVEY 8d extras: hash encode against HBM at the C5 table size (T = 2^22, 185 MiB fp16: not L2-resident),
        # compositing bandwidth, field-MLP tensor throughput
        from google_nerf_b200 import tinycudann as tcnn_b
        lay5 = tcnn_b.hashgrid_layout(16, 2, 22, 16, float(np.exp(np.log(2048 * 16 / 16) / 15)))
        tab5 = (torch.rand(lay5.n_params, device=dev) * 2e-4 - 1e-4).half()
        n5 = 1 << 20
        x5 = torch.rand(n5, 3, device=dev); enc5 = torch.empty(n5, 32, dtype=torch.float16, device=dev)
        t5 = []
        for _ in range(4):
            a5, b5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a5.record(); LL.call("b2n_hashgrid_fw", LL.ptr(x5), LL.ptr(tab5), lay5, n5, None, LL.ptr(enc5), 32); b5.record()
            torch.cuda.synchronize(); t5.append(a5.elapsed_time(b5) * 1e-3)
        # uniformly random points: every fine-level gather pulls a distinct 32-byte sector from HBM
        hash_encode["c5_T22_random_points"] = dict(
            fw_gbs=588 * n5 / min(t5[1:]) / 1e9, table_mib=lay5.n_params * 2 / 2 ** 20, points=n5,
            sector_gbs=(12 + 64 + 16 * 8 * 32) * n5 / min(t5[1:]) / 1e9, frac_of_hbm_sectors=((12 + 64 + 16 * 8 * 32) * n5 /
                                                                                     min(t5[1:]) / 1e9) / hbm_peak,
            note="T=2^22 (SURVEY C5): algorithmic 588 B/point; sector_gbs counts one 32-B sector per 4-B gather (upper "
                 "bound of the DRAM traffic, coarse levels hit in cache)")
        del tab5, x5, enc5
        t_c = table.get("b2n_composite_loss_fwbw", 0) * 1e-3
        compositing = dict(gbs=(40 * samples + 76 * N_RAYS) / t_c / 1e9 if t_c else None,
                           note="fused fw + loss + bw launch; algorithmic 40 B/sample + 76 B/ray; latency-bound at 8192 rays")
        t_f, t_b = table.get("b2n_field_mlp_fw", 0) * 1e-3, table.get("b2n_field_mlp_bw", 0) * 1e-3
        mlp = dict(fw_tflops=20480 * samples / t_f / 1e12 if t_f else None,
                   bw_tflops=40960 * alive / t_b / 1e12 if t_b else None, tensor_peak_tflops=tf_peak,
                   fw_frac_of_tensor_peak=20480 * samples / t_f / 1e12 / tf_peak if t_f else None,
                   note="tcgen05 kind::f16; the MLPs are 64 wide: activation traffic and dependency latency bound them, "
                        "not the tensor pipe (ncu sm__pipe_tensor_cycles_active in profiles/)")
        cb = None
        if not args.skip_cpu and world == 1:                  # rank 0 at N = 1 only (the other ranks would spin in NCCL)
            cb, _ = cpu_baseline(args.cpu_rays, 12, 1)         # ~10 s of CPU work on 16 host cores
        rays = N_RAYS * world * args.steps
        n_updates = sum(1 for s in range(args.steps) if s % tr.S == 0)
        line = dict(metric="train_rays_per_s", value=rays / (ms * 1e-3), unit="rays/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f16", data="synthetic (analytic 3-sphere+box scene, random-init weights "
                    f"trained {args.pretrain} untimed steps to steady-state occupancy)",
                    config=dict(workload=WORKLOAD, rays_per_gpu=N_RAYS, samples_per_step=samples, alive_samples_per_step=alive,
                                samples_per_ray=samples / N_RAYS, cuda_graph=not args.no_graph,
                                l2="per-step working set (206 MB optimiser state + sample buffers) exceeds the 126 MB "
                                   "L2; no explicit flush", parallelism=f"dp{world}", comm=tr.comm),
                    clocks=clk,
                    e2e=dict(value=rays / (ms_e2e * 1e-3), unit="rays/s", h2d_bytes_per_step=N_RAYS * (8 + 8 + 12),
                             d2h_bytes_per_step=4, ms_per_step=ms_e2e / args.steps, last_loss=last_loss),
                    gpu_launches=launches_per_step * args.steps + 12 * n_updates,
                    roofline=roofline, cpu_baseline=cb, hash_encode=hash_encode, marcher=marcher, compositing=compositing, mlp=mlp, render=render_info, quality=quality, grid_update_ms=grid_update_ms,
                    kernels_us={k: round(v * 1e3, 1) for k, v in table.items()})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
