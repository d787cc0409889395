"""Round-loop render time of the tree this file is run against (python scratch/round_loop_time.py <tree root>)."""
import sys, os, torch
root = os.path.abspath(sys.argv[1]); sys.path.insert(0, root)
from google_nerf_b200 import synthetic as syn
from google_nerf_b200.models.networks import NGP
from google_nerf_b200.models.rendering import render
from google_nerf_b200.trainer import NGPTrainer
import google_nerf_b200
print("package from", os.path.dirname(google_nerf_b200.__file__))
dev = torch.device("cuda")
torch.manual_seed(1337)
K = syn.intrinsics(800, 800); dirs = syn.directions(800, 800, K); poses = syn.hemisphere_poses(100)
model = NGP(0.5).to(dev)
tr = NGPTrainer(model, n_rays=8192, use_graph=True, samples_per_ray=160)
tr.set_dataset(dirs, poses)
model.mark_invisible_cells(K.to(dev), poses.to(dev), (800, 800))
dd, pp = dirs.to(dev), poses.to(dev)
g = torch.Generator().manual_seed(1)
for step in range(1000):
    ii = torch.randint(100, (8192,), generator=g).to(dev); pi = torch.randint(640000, (8192,), generator=g).to(dev)
    ro, rd = syn.get_rays(dd[pi], pp[ii])
    tr.step_batch({"img_idxs": ii, "pix_idxs": pi, "rgb": syn.shade(ro, rd, 0.5)})
tr.sync_model()
ro, rd = syn.get_rays(dd, pp[0])
kw = dict(test_time=True, T_threshold=1e-2)
from google_nerf_b200.models import rendering
if hasattr(rendering, "_WholeRays"):
    kw["whole_rays"] = False
with torch.no_grad():
    for _ in range(3):
        res = render(model, ro, rd, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); res = render(model, ro, rd, **kw); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"round loop: ms {min(ts):.3f} median {sorted(ts)[4]:.3f} samples/ray {res['total_samples'] / len(ro):.2f}")
from torch.profiler import profile, ProfilerActivity
with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
    render(model, ro, rd, **kw); torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6]:
    print(f"   {e.key[:60]:60s} {e.device_time_total:8.0f} us  x{e.count}")
