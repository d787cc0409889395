"""Kernel micro-benchmark on the real sample distribution: trains config c2 to steady state with the product library,
then times the hot kernels one by one (CUDA events, median of 30 launches, back to back like inside a step).
With B2N_LIB pointing at a variant build (google-nerf_b200/build.py: B2N_VARIANT / B2N_EXTRA_FLAGS) this compares
tuning constants:  for v in ...; do B2N_LIB=google-nerf_b200/lib/libb2n_$v.so python scratch/bench_kernels.py; done"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B
from google_nerf_b200 import _lib as L
from google_nerf_b200.models.networks import NGP
from google_nerf_b200.trainer import NGPTrainer

dev = torch.device("cuda", 0)
wl = B.Workload(os.environ.get("CFG", "c2")); cfg, syn = wl.cfg, wl.syn
torch.manual_seed(1337)
model = NGP(cfg["scale"], log2_T=cfg["log2_T"]).to(dev)
tr = NGPTrainer(model, n_rays=cfg["n_rays"], use_graph=True, seed=1234, samples_per_ray=cfg["spr"], exp_step_factor=cfg["esf"])
tr.set_dataset(wl.dirs, wl.poses)
model.mark_invisible_cells(wl.K.to(dev), wl.poses.to(dev), (cfg["W"], cfg["H"]))
g = torch.Generator(device=dev).manual_seed(7)
dd, pp = wl.dirs.to(dev), wl.poses.to(dev)
def batch():
    ii = torch.randint(cfg["n_img"], (cfg["n_rays"],), device=dev, generator=g)
    pi = torch.randint(cfg["W"] * cfg["H"], (cfg["n_rays"],), device=dev, generator=g)
    ro, rd = syn.get_rays(dd[pi], pp[ii])
    return {"img_idxs": ii, "pix_idxs": pi, "rgb": wl.shade(ro, rd)[0]}
for k in range(int(os.environ.get("PRETRAIN", 400))):
    tr.step_batch(batch())
    if k % 50 == 49 and tr.overflowed():
        tr.grow(1.5)
torch.cuda.synchronize()
tr.use_graph = False
s = tr.sets[tr.cur]
tr.step_count += 1; tr._set_hyper(); tr._march(s); tr._forward_backward(s)      # fills every buffer of the step
torch.cuda.synchronize()
P, cap, nd = L.ptr, tr.capacity, s.counter
din_enc = tr.din_enc                     # filled by the real step above (zeros would let the scatter skip its reds)
calls = {
    "hashgrid_fw": lambda: L.call("b2n_hashgrid_fw", P(s.xyzs), P(tr.h_xyz[tr.n_mlp:]), tr.layout, cap, P(nd), P(tr.enc), 32),
    "field_mlp_fw": lambda: L.call("b2n_field_mlp_fw", P(tr.enc), 32, P(s.dirs), P(tr.w_image), cap, P(nd), P(tr.sigmas), P(tr.rgbs), P(tr.h)),
    "field_mlp_bw": lambda: L.call("b2n_field_mlp_bw", P(tr.dL_dsigmas), P(tr.dL_drgbs), P(tr.enc), 32, P(s.dirs), P(tr.w_image), cap,
                                   P(tr.alive_cnt), P(tr.rgbs), P(tr.h), 1.0, P(din_enc), P(tr.g_xyz), P(tr.g_rgb), P(tr.alive_idx), 0, None),
    "hashgrid_bw": lambda: L.call("b2n_hashgrid_bw", P(s.xyzs), P(din_enc), 32, tr.layout, cap, P(tr.alive_cnt), 1.0,
                                  P(tr.g_xyz[tr.n_mlp:]), P(tr.alive_idx)),
}
only = os.environ.get("ONLY")
out = {"lib": os.path.basename(L.LIB_PATH), "samples": int(nd[0]), "alive": int(tr.alive_cnt)}
for name, fn in calls.items():
    if only and name not in only.split(","):
        continue
    for _ in range(3):
        fn()
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    out[name] = round(ts[len(ts) // 2], 1)
print(json.dumps(out))
