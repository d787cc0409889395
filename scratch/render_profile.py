"""Kernel-level breakdown of one full-frame test-time render (torch.profiler).  python scratch/render_profile.py [steps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import profile, ProfilerActivity
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
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 600):
    ii = torch.randint(100, (8192,), generator=g).to(dev); pi = torch.randint(640000, (8192,), generator=g).to(dev)
    ro, rd = syn.get_rays(dd[pi], pp[ii])
    tr.step_batch({"img_idxs": ii, "pix_idxs": pi, "rgb": syn.shade(ro, rd, 0.5)})
tr.sync_model()
ro, rd = syn.get_rays(dd, pp[0])
with torch.no_grad():
    for _ in range(2):
        res = render(model, ro, rd, test_time=True, T_threshold=1e-2, whole_rays=False)
    torch.cuda.synchronize()
    st = model._device_loop
    print("rounds", int(st.ctl_host[5]), "samples", res["total_samples"], "alive left", int(st.ctl_host[4]))
    import time
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); res = render(model, ro, rd, test_time=True, T_threshold=1e-2, whole_rays=False); torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print("frame ms", [round(t, 3) for t in ts])
    if os.environ.get("NOPROF"):
        sys.exit(0)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        res = render(model, ro, rd, test_time=True, T_threshold=1e-2, whole_rays=False)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=50))
