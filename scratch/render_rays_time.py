"""Whole-ray kernel vs round loop on a trained c2-like scene: full 800x800 frame and a 1/8 share of its rows.
python scratch/render_rays_time.py [train_steps]"""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from google_nerf_b200 import synthetic as syn
from google_nerf_b200.models.networks import NGP
from google_nerf_b200.models.rendering import render
from google_nerf_b200.trainer import NGPTrainer
dev = torch.device("cuda")
torch.manual_seed(1337)
K = syn.intrinsics(800, 800); dirs = syn.directions(800, 800, K); poses = syn.hemisphere_poses(100)
model = NGP(0.5).to(dev)
tr = NGPTrainer(model, n_rays=8192, use_graph=True, samples_per_ray=160)
tr.set_dataset(dirs, poses)
model.mark_invisible_cells(K.to(dev), poses.to(dev), (800, 800))
dd, pp = dirs.to(dev), poses.to(dev)
g = torch.Generator().manual_seed(1)
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1000):
    ii = torch.randint(100, (8192,), generator=g).to(dev); pi = torch.randint(640000, (8192,), generator=g).to(dev)
    ro, rd = syn.get_rays(dd[pi], pp[ii])
    tr.step_batch({"img_idxs": ii, "pix_idxs": pi, "rgb": syn.shade(ro, rd, 0.5)})
tr.sync_model()
ro_f, rd_f = syn.get_rays(dd, pp[0])
rows = torch.arange(800, device=dev)
share = (rows % 8 == 0).repeat_interleave(800)            # every 8th row: one rank's tiles at N = 8
for name, ro, rd in (("full", ro_f, rd_f), ("eighth", ro_f[share].contiguous(), rd_f[share].contiguous())):
    for wr, fh in (((True, True),) if os.environ.get('RR_ONLY') else ((False, True), (True, False), (True, True))):
        from google_nerf_b200.models.rendering import _WholeRays
        _WholeRays.FIRST_HIT = fh
        with torch.no_grad():
            for _ in range(3):
                res = render(model, ro, rd, test_time=True, T_threshold=1e-2, whole_rays=wr)
            torch.cuda.synchronize()
            ts = []
            for _ in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); res = render(model, ro, rd, test_time=True, T_threshold=1e-2, whole_rays=wr); e1.record()
                torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        extra = f"rounds {model._whole_rays.rounds}" if wr else f"rounds {int(model._device_loop.ctl_host[5])}"
        if wr and fh:
            import ctypes
            from google_nerf_b200 import _lib as L
            lib = ctypes.CDLL(L.LIB_PATH)
            if hasattr(lib, "b2n_render_profile"):
                buf = (ctypes.c_ulonglong * 8)()
                lib.b2n_render_profile(buf, 1)
                render(model, ro, rd, test_time=True, T_threshold=1e-2, whole_rays=True)
                lib.b2n_render_profile(buf, 1)
                tot = sum(buf[:5])
                extra += "  phases A/B/C/D/E % " + " ".join(f"{100 * buf[i] / tot:.1f}" for i in range(5)) + \
                         f"  cycles/round {tot / max(buf[5], 1):.0f}  cta-rounds {buf[5]}"
            from torch.profiler import profile, ProfilerActivity
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                render(model, ro, rd, test_time=True, T_threshold=1e-2, whole_rays=True)
                torch.cuda.synchronize()
            extra += "\n      kernels: " + ", ".join(f"{e.key[:28]} {e.device_time_total:.0f}us" for e in
                                                   sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6])
        print(f"{name:7s} whole_rays={wr!s:5s} first_hit={fh!s:5s} ms {min(ts):.3f} (median {sorted(ts)[4]:.3f})  Mrays/s {len(ro) / min(ts) / 1e3:.1f}  "
              f"samples/ray {res['total_samples'] / len(ro):.2f}  {extra}", flush=True)
    a = render(model, ro, rd, test_time=True, T_threshold=1e-2, whole_rays=False)
    b = render(model, ro, rd, test_time=True, T_threshold=1e-2, whole_rays=True)
    print("   max |d rgb|", float((a["rgb"] - b["rgb"]).abs().max()), "frac > 1e-5", float(((a["rgb"] - b["rgb"]).abs() > 1e-5).float().mean()))
