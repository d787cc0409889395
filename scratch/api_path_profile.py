"""Where does a step of the reference call path (render + NeRFLoss + backward + FusedAdam) spend its time?
torch.profiler table of 20 steps at the headline batch size (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "google-nerf_b200", "shims"))
import torch
import bench as B
from apex.optimizers import FusedAdam
from google_nerf_b200.losses import NeRFLoss
from google_nerf_b200.models.networks import NGP
from google_nerf_b200.models.rendering import render

dev = torch.device("cuda", 0)
wl = B.Workload("c2"); c, syn = wl.cfg, wl.syn
model = NGP(c["scale"], log2_T=c["log2_T"]).to(dev).init_grid_buffers()
model.mark_invisible_cells(wl.K.to(dev), wl.poses.to(dev), (c["W"], c["H"]))
opt = FusedAdam(model.parameters(), 1e-2, eps=1e-15)
loss_fn = NeRFLoss()
dd, pp = wl.dirs.to(dev), wl.poses.to(dev)
g = torch.Generator(device=dev).manual_seed(5)
N = 80
img = torch.randint(c["n_img"], (N, c["n_rays"]), device=dev, generator=g)
pix = torch.randint(c["W"] * c["H"], (N, c["n_rays"]), device=dev, generator=g)
tgt = torch.stack([wl.shade(*syn.get_rays(dd[pix[k]], pp[img[k]]))[0] for k in range(N)])
def step(k):
    if k % 16 == 0:
        model.update_density_grid(0.01 * 1024 / 3 ** 0.5, warmup=k < 256)
    ro, rd = syn.get_rays(dd[pix[k]], pp[img[k]])
    res = render(model, ro, rd)
    loss = sum(v.mean() for v in loss_fn(res, {"rgb": tgt[k]}).values())
    opt.zero_grad(); loss.backward(); opt.step()
for k in range(40): step(k)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(40, 60): step(k)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host issue {1e3*(t1-t0)/20:.3f} ms/step, with drain {1e3*(t2-t0)/20:.3f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for k in range(60, 80): step(k)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=22, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=22, max_name_column_width=60))
