"""Which part of the end-to-end table-gradient error is fp16 underflow?  One C2-size step vs the oracle at several
loss scales (run on the GPU box: python scratch/loss_scale_parity.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import make_scene
from test_gpu_baseline import _grad_report
from google_nerf_b200.models.networks import NGP
from google_nerf_b200.trainer import NGPTrainer
from oracle import ngp_ref as O

scale, log2_T, n_rays = 0.5, 19, 4096
s = make_scene(scale, n_rays, seed=21)
ref = O.NGPRef(scale, log2_T=log2_T, seed=3)
g = torch.Generator().manual_seed(9)
with torch.no_grad():
    ref.xyz_params[ref.n_mlp:] = (torch.rand(ref.layout["n_params"], generator=g) * 2 - 1) * 0.5
ref.density_bitfield = s["bitfield"].clone()
target = torch.rand(n_rays, 3, generator=g)
res = O.render(ref, s["rays_o"], s["rays_d"].clone(), noise=s["noise"])
O.nerf_loss(res, target).backward()
for ls in (128.0, 1024.0, 8192.0, 65536.0):
    model = NGP(scale, log2_T=log2_T).to("cuda")
    model.xyz_encoder.params.data.copy_(ref.xyz_params.detach()); model.rgb_net.params.data.copy_(ref.rgb_params.detach())
    model.density_bitfield.copy_(s["bitfield"])
    tr = NGPTrainer(model, n_rays=n_rays, use_graph=False, samples_per_ray=128, grid_update_interval=10 ** 9, loss_scale=ls)
    tr.step_count = 1; tr.fixed_noise = s["noise"].cuda()
    tr.set_batch(s["rays_o"].cuda(), s["rays_d"].cuda(), target.cuda())
    ss = tr.sets[tr.cur]
    tr._set_hyper(); tr._march(ss); tr._forward_backward(ss)
    rep = _grad_report((tr.g_xyz, tr.g_rgb), ref, ref.n_mlp, ref.layout["offsets"], unscale=1.0 / ls)
    print("norms got/want W3", float((tr.g_rgb[:2048] / ls).norm()), float(ref.rgb_params.grad[:2048].norm()),
          "loss", float(tr.loss), "samples", int(ss.counter[0]), res["total_samples"])
    print(f"loss_scale {ls:8.0f} found_inf {int(tr.hyper[2])} max|din_enc| {float(tr.din_enc[:int(tr.alive_cnt)].abs().max()):.3g} " +
          " ".join(f"{k}=({v[0]:.1e},{v[1]:.1e})" for k, v in rep.items()))
