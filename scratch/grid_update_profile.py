"""Kernel-level breakdown of one occupancy-grid update and of one eager training step (torch.profiler / CUPTI).
Run on the GPU box:  python scratch/grid_update_profile.py [pretrain_steps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import profile, ProfilerActivity
from google_nerf_b200 import synthetic as syn
from google_nerf_b200.models.networks import NGP
from google_nerf_b200.trainer import NGPTrainer
dev = torch.device("cuda")
torch.manual_seed(1337)
K = syn.intrinsics(800, 800); dirs = syn.directions(800, 800, K); poses = syn.hemisphere_poses(100)
model = NGP(0.5).to(dev)
tr = NGPTrainer(model, n_rays=8192, use_graph=False, samples_per_ray=160)
tr.set_dataset(dirs, poses)
model.mark_invisible_cells(K.to(dev), poses.to(dev), (800, 800))
dd, pp = dirs.to(dev), poses.to(dev)
g = torch.Generator().manual_seed(1)


def batch():
    ii = torch.randint(100, (8192,), generator=g).to(dev); pi = torch.randint(640000, (8192,), generator=g).to(dev)
    ro, rd = syn.get_rays(dd[pi], pp[ii])
    return {"img_idxs": ii, "pix_idxs": pi, "rgb": syn.shade(ro, rd, 0.5)}


for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 400):
    if step % 16 == 0:
        tr.update_density_grid(warmup=step < 256)
    tr.step_batch(batch())
torch.cuda.synchronize()
N = s_rays = tr.sets[tr.cur].rays_a[:, 2]
q = torch.quantile(N.float(), torch.tensor([0.5, 0.9, 0.99, 1.0], device=dev))
print("samples/ray quantiles 50/90/99/100:", q.tolist(), "rays with samples", (N > 0).float().mean().item())
for name, fn in (("grid update", lambda: tr.update_density_grid(False)), ("train step", lambda: tr.step_batch(b))):
    b = batch()
    fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
    print("=====", name, "(4 repetitions)")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
