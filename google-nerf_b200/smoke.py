"""One small invocation of the hot path on the GPU, checked against the oracle (the checker, never the
thing shipped): a 512-ray training step of the HashGrid NGP -- AABB, marcher, hash encode, MLPs, compositing,
loss, backward, Adam -- and a 128-ray test-time render, compared with oracle/ngp_ref.py on the same inputs, jitter and
weights."""
import torch


def run(device):
    from . import synthetic as syn
    from .models.networks import NGP
    from .trainer import NGPTrainer
    from oracle import ngp_ref as O                       # smoke() is one of the three allowed oracle users

    torch.manual_seed(0)
    scale, n, log2_T = 0.5, 512, 15
    ref = O.NGPRef(scale, log2_T=log2_T, seed=1337)
    grid = syn.density_grid(scale, 1)
    ref.density_bitfield = syn.bitfield_from_grid(grid)
    K = syn.intrinsics(800, 800); dirs = syn.directions(800, 800, K); poses = syn.hemisphere_poses(8)
    ii = torch.randint(8, (n,)); pi = torch.randint(800 * 800, (n,))
    rays_o, rays_d = syn.get_rays(dirs[pi], poses[ii])
    target = syn.shade(rays_o, rays_d, scale)
    noise = torch.rand(n)

    model = NGP(scale, log2_T=log2_T).to(device)
    model.xyz_encoder.params.data.copy_(ref.xyz_params.detach())
    model.rgb_net.params.data.copy_(ref.rgb_params.detach())
    model.density_bitfield.copy_(ref.density_bitfield)
    # test-time render of a few rays: the whole-ray kernel (march + hash grid + tcgen05 field + compositing in one
    # launch) against the oracle's host loop of rendering.py:42-114
    from .models.rendering import render
    m = 128
    res_t = O.render(ref, rays_o[:m], rays_d[:m].clone(), test_time=True, T_threshold=1e-2)
    got_t = render(model, rays_o[:m].to(device), rays_d[:m].to(device).clone(), test_time=True, T_threshold=1e-2)
    assert model._whole_rays.rounds > 0 and got_t["total_samples"] > 0
    for k in ("opacity", "depth", "rgb"):
        torch.testing.assert_close(got_t[k].cpu(), res_t[k], rtol=1e-2, atol=1e-2, msg=lambda msg: f"render {k}: {msg}")

    tr = NGPTrainer(model, n_rays=n, use_graph=False, samples_per_ray=160, warmup_steps=0, grid_update_interval=10 ** 9)
    tr.step_count = 2                                      # no grid update: the bitfield is the analytic one
    tr.fixed_noise = noise.to(device)
    tr.set_batch(rays_o.to(device), rays_d.to(device), target.to(device))
    p_before = tr.p_xyz.clone()
    sset = tr.sets[tr.cur]
    tr._set_hyper(); tr._march(sset); tr._forward_backward(sset); tr.last_counter = sset.counter
    g_xyz, g_rgb = tr.g_xyz.clone() / tr.loss_scale, tr.g_rgb.clone() / tr.loss_scale
    loss_gpu = float(tr.loss.item())
    n_samples = tr.samples_last_step()

    res = O.render(ref, rays_o, rays_d.clone(), noise=noise)
    loss = O.nerf_loss(res, target)
    (loss * tr.loss_scale).backward()                     # same loss scaling as the CUDA path (fp16 activation grads)
    assert res["total_samples"] == n_samples, (res["total_samples"], n_samples)
    assert abs(loss_gpu - loss.item()) < 2e-3 * abs(loss.item()), (loss_gpu, loss.item())
    torch.testing.assert_close(tr.opacity.cpu(), res["opacity"].detach(), rtol=5e-3, atol=5e-3)
    for got, want, name in ((g_rgb, ref.rgb_params.grad / tr.loss_scale, "rgb_net"),
                            (g_xyz, ref.xyz_params.grad / tr.loss_scale, "xyz_encoder")):
        s = want.abs().max().item()
        err = (got.cpu() - want).abs().max().item()
        assert err <= 1e-2 * s, (name, err, s)
    tr._optimizer()
    assert float((tr.p_xyz - p_before).abs().max()) > 0
    torch.cuda.synchronize()
    print(f"smoke ok: {n} rays, {n_samples} samples, loss {loss_gpu:.6f} (oracle {loss.item():.6f}); "
          f"test-time render of {m} rays: {got_t['total_samples']} samples")
