"""Where the data-parallel step spends its time: CUDA-event segments around the peer exchange, eager mode.
torchrun --nproc-per-node N scratch/dp_segments.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from google_nerf_b200 import synthetic as syn, _lib as L
from google_nerf_b200.models.networks import NGP
from google_nerf_b200.trainer import NGPTrainer
torch.manual_seed(1337)
K = syn.intrinsics(800, 800); dirs = syn.directions(800, 800, K); poses = syn.hemisphere_poses(100)
model = NGP(0.5).to(dev)
tr = NGPTrainer(model, n_rays=8192, use_graph=True, seed=1234 + rank, samples_per_ray=160)
tr.set_dataset(dirs, poses)
model.mark_invisible_cells(K.to(dev), poses.to(dev), (800, 800))
dd, pp = dirs.to(dev), poses.to(dev)
g = torch.Generator().manual_seed(100 + rank)
def batch():
    ii = torch.randint(100, (8192,), generator=g).to(dev); pi = torch.randint(640000, (8192,), generator=g).to(dev)
    ro, rd = syn.get_rays(dd[pi], pp[ii])
    return {"img_idxs": ii, "pix_idxs": pi, "rgb": syn.shade(ro, rd, 0.5)}
for step in range(512):
    tr.step_batch(batch())
torch.cuda.synchronize(); dist.barrier()
# eager, instrumented steps
tr.use_graph = False
P, call, pb = L.ptr, L.call, tr.peer
names = ["march", "fw+bw", "-", "-", "exchange (all)", "-"]
acc = torch.zeros(len(names))
reps = 24
for it in range(reps + 4):
    b = batch(); s = tr.sets[tr.cur]; tr._load(s, b)
    tr.step_count += 1; tr._set_hyper()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record(); tr._march(s)
    ev[1].record(); tr._forward_backward(s)
    ev[2].record(); ev[3].record(); ev[4].record()
    tr._optimizer_peer()                                      # (pack16 +) barrier + adam_peer + barrier + zero + pack
    ev[5].record()
    ev[6].record(); torch.cuda.synchronize()
    if it >= 4:
        acc += torch.tensor([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(len(names))])
acc /= reps
allv = [torch.zeros_like(acc).to(dev) for _ in range(dist.get_world_size())]
dist.all_gather(allv, acc.to(dev))
if rank == 0:
    print("world", dist.get_world_size(), "segments (us), one row per rank:", names)
    for r, v in enumerate(allv):
        print(r, [round(float(x), 1) for x in v])
pb.check()
dist.barrier(); dist.destroy_process_group()
