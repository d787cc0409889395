"""In-situ timeline of the graph-replayed training step (CUPTI through torch.profiler): start / duration / stream of every
kernel of a few consecutive steps, plus the gaps along the main-stream chain.
Run on the GPU box:  python scratch/step_timeline.py [pretrain_steps]"""
import sys, os, json, tempfile, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import profile, ProfilerActivity
from google_nerf_b200 import synthetic as syn
from google_nerf_b200.models.networks import NGP
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


def batch():
    ii = torch.randint(100, (8192,), generator=g).to(dev); pi = torch.randint(640000, (8192,), generator=g).to(dev)
    ro, rd = syn.get_rays(dd[pi], pp[ii])
    return {"img_idxs": ii, "pix_idxs": pi, "rgb": syn.shade(ro, rd, 0.5)}


batches = [batch() for _ in range(8)]
n_pre = int(sys.argv[1]) if len(sys.argv) > 1 else 400
step = 0


def one(nb=True):
    global step
    if step % 16 == 0:
        tr.update_density_grid(warmup=step < 256)
    b = batches[step % 8]; nxt = batches[(step + 1) % 8]
    tr.step_batch(b if not tr.sets[tr.cur].marched else None, next_batch=None if (step + 1) % 16 == 0 else nxt)
    step += 1


for _ in range(n_pre):
    one()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(64):
    one()
e1.record(); torch.cuda.synchronize()
print("unprofiled: %.1f us/step, samples %d, alive %d" % (e0.elapsed_time(e1) / 64 * 1e3, tr.samples_last_step(), int(tr.alive_cnt.item())))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(20):
        one()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
# pick steps: a step begins at each hashgrid_fw launch on the main chain with ~394k samples (skip grid-update ones)
names = lambda e: e["name"].split("(")[0].replace("void ", "")[:44]
starts = [i for i, e in enumerate(ev) if names(e).startswith("composite_loss")]
print("events", len(ev), "steps seen", len(starts))
# print the window of the 6th..8th step
lo = ev[starts[5]]["ts"] - 250
hi = ev[starts[8]]["ts"]
prev_end = {}
for e in ev:
    if lo <= e["ts"] <= hi:
        st = e["args"].get("stream")
        gap = e["ts"] - prev_end.get(st, e["ts"])
        print(f"{e['ts'] - lo:9.1f} us  dur {e['dur']:7.1f}  gap {gap:6.1f}  stream {st}  {names(e)}")
        prev_end[st] = e["ts"] + e["dur"]
# step period
per = [ev[starts[i + 1]]["ts"] - ev[starts[i]]["ts"] for i in range(len(starts) - 1)]
print("step periods (us):", [round(p, 1) for p in per])
