"""Where a sharded test-time frame spends its time (torchrun, >= 2 GPUs): wall clock per frame + torch.profiler of one
frame on rank 0.   python -m torch.distributed.run --nproc-per-node 2 ... scratch/sharded_profile.py"""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from google_nerf_b200 import synthetic as syn
from google_nerf_b200.models.networks import NGP
from google_nerf_b200.models.rendering import render
from google_nerf_b200.dist_utils import render_sharded
from google_nerf_b200.trainer import NGPTrainer
torch.manual_seed(1337)
K = syn.intrinsics(800, 800); dirs = syn.directions(800, 800, K); poses = syn.hemisphere_poses(100)
model = NGP(0.5).to(dev)
tr = NGPTrainer(model, n_rays=8192, use_graph=True, samples_per_ray=160, data_parallel=False)
tr.set_dataset(dirs, poses)
model.mark_invisible_cells(K.to(dev), poses.to(dev), (800, 800))
dd, pp = dirs.to(dev), poses.to(dev)
g = torch.Generator().manual_seed(1)
for step in range(600):
    ii = torch.randint(100, (8192,), generator=g).to(dev); pi = torch.randint(640000, (8192,), generator=g).to(dev)
    ro, rd = syn.get_rays(dd[pi], pp[ii])
    tr.step_batch({"img_idxs": ii, "pix_idxs": pi, "rgb": syn.shade(ro, rd, 0.5)})
tr.sync_model()
ro, rd = syn.get_rays(dd, pp[0])
tile = int(os.environ.get("TILE", 800))
fn = lambda o, d, **kw: render(model, o, d, **kw)
kw = dict(test_time=True, T_threshold=1e-2)
with torch.no_grad():
    for _ in range(3):
        render_sharded(fn, ro, rd, tile=tile, **kw)
    ts = []
    for _ in range(8):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        render_sharded(fn, ro, rd, tile=tile, **kw)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    if rank == 0:
        print("world", dist.get_world_size(), "frame ms", [round(t, 3) for t in ts])
    from torch.profiler import profile, ProfilerActivity
    dist.barrier(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        render_sharded(fn, ro, rd, tile=tile, **kw); torch.cuda.synchronize()
    if rank == 0:
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=48))
        print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=48))
dist.barrier(); dist.destroy_process_group()
