import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 600):
    ii = torch.randint(100, (8192,), generator=g).to(dev); pi = torch.randint(640000, (8192,), generator=g).to(dev)
    ro, rd = syn.get_rays(dd[pi], pp[ii])
    tr.step_batch({"img_idxs": ii, "pix_idxs": pi, "rgb": syn.shade(ro, rd, 0.5)})
    if step % 100 == 99:
        n = tr.samples_last_step()
        ds, dc = tr.dL_dsigmas[:n], tr.dL_drgbs[:n]
        alive = (ds != 0) | (dc != 0).any(1)
        s = tr.sets[tr.cur]
        N = s.rays_a[:, 2]
        print(step, "samples", n, "alive frac %.3f" % alive.float().mean().item(), "rays with samples %.3f" % (N > 0).float().mean().item(),
              "max N", int(N.max()), "mean opacity %.3f" % tr.opacity.mean().item(), "loss %.5f" % tr.loss.item(),
              "dead-tile frac %.3f" % (1 - alive[: n // 128 * 128].view(-1, 128).any(1).float().mean().item()))
