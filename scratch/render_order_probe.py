"""Why is the whole-ray frame faster after round-loop frames?  Event-timed frames in alternating blocks + SM clock."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
pynvml.nvmlInit(); H = pynvml.nvmlDeviceGetHandleByIndex(0)
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
for step in range(1000):
    ii = torch.randint(100, (8192,), generator=g).to(dev); pi = torch.randint(640000, (8192,), generator=g).to(dev)
    ro, rd = syn.get_rays(dd[pi], pp[ii])
    tr.step_batch({"img_idxs": ii, "pix_idxs": pi, "rgb": syn.shade(ro, rd, 0.5)})
tr.sync_model()
ro, rd = syn.get_rays(dd, pp[0])

def block(name, wr, n=5):
    ts = []
    with torch.no_grad():
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res = render(model, ro, rd, test_time=True, T_threshold=1e-2, whole_rays=wr); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    clk = pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM); pw = pynvml.nvmlDeviceGetPowerUsage(H) / 1000
    print(f"{name:24s} ms {[round(t, 2) for t in ts]}  samples/ray {res['total_samples'] / len(ro):.2f}  sm {clk} MHz {pw:.0f} W", flush=True)

block("whole (first)", True)
block("whole again", True)
block("round loop", False)
block("whole after round", True)
block("whole again", True)
time.sleep(1.0)
block("whole after 1 s idle", True)
for step in range(300):
    ii = torch.randint(100, (8192,), generator=g).to(dev); pi = torch.randint(640000, (8192,), generator=g).to(dev)
    ro2, rd2 = syn.get_rays(dd[pi], pp[ii])
    tr.step_batch({"img_idxs": ii, "pix_idxs": pi, "rgb": syn.shade(ro2, rd2, 0.5)})
tr.sync_model(); torch.cuda.synchronize()
block("whole after training", True)
block("round after that", False)
block("whole", True)
