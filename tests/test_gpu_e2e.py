"""End-to-end parity of the drop-in API (NGP + render() + autograd) and of the fused trainer against the
oracle, on the same weights, occupancy bitfield and jitter."""
import numpy as np
import pytest
import torch

from conftest import make_scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def pair(built_lib):
    """(oracle NGPRef, CUDA NGP) sharing weights and the analytic bitfield; HashGrid with a small table."""
    from google_nerf_b200.models.networks import NGP
    from oracle import ngp_ref as O
    s = make_scene(0.5, 768, seed=11)
    ref = O.NGPRef(0.5, log2_T=15, seed=1337)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():                                   # larger table values so that sigma varies in space
        ref.xyz_params[ref.n_mlp:] = (torch.rand(ref.layout["n_params"], generator=g) * 2 - 1) * 0.5
    ref.density_bitfield = s["bitfield"].clone()
    model = NGP(0.5, log2_T=15).to(DEV)
    model.xyz_encoder.params.data.copy_(ref.xyz_params.detach())
    model.rgb_net.params.data.copy_(ref.rgb_params.detach())
    model.density_bitfield.copy_(s["bitfield"])
    return ref, model, s


def test_state_dict_contract(pair):
    _, model, _ = pair
    sd = model.state_dict()
    assert set(sd) == {"center", "xyz_min", "xyz_max", "half_size", "density_bitfield", "xyz_encoder.params",
                       "dir_encoder.params", "rgb_net.params"}
    assert sd["density_bitfield"].dtype == torch.uint8 and sd["density_bitfield"].numel() == 128 ** 3 // 8
    assert sd["dir_encoder.params"].numel() == 0 and sd["rgb_net.params"].numel() == 7168
    assert sd["xyz_encoder.params"].dtype == torch.float32
    model.init_grid_buffers()
    assert model.density_grid.shape == (1, 128 ** 3) and model.grid_coords.shape == (128 ** 3, 3)
    from google_nerf_b200.models.networks import NGP
    full = NGP(0.5)                                          # reference defaults: T = 2^19
    assert full.xyz_encoder.params.numel() == 11420064 + 3072
    freq = NGP(0.5, encoding="Frequency")                    # this fork's active encoding (networks.py:49-53)
    assert freq.xyz_encoder.params.numel() == 6144


def test_render_train_forward_backward(pair):
    from google_nerf_b200.models.rendering import render
    from google_nerf_b200.models.custom_functions import RayMarcher
    from oracle import ngp_ref as O
    ref, model, s = pair
    target = torch.rand(s["rays_o"].shape[0], 3, generator=torch.Generator().manual_seed(1))
    res_ref = O.render(ref, s["rays_o"], s["rays_d"].clone(), noise=s["noise"])
    loss_ref = O.nerf_loss(res_ref, target)
    (loss_ref * 1024.0).backward()                          # GradScaler-style scaling: fp16 grads must not underflow

    RayMarcher.noise = s["noise"].to(DEV)
    try:
        res = render(model, s["rays_o"].to(DEV), s["rays_d"].to(DEV).clone())
    finally:
        RayMarcher.noise = None
    assert int(res["total_samples"]) == res_ref["total_samples"]
    for k in ("opacity", "depth", "rgb", "depth_sq"):
        torch.testing.assert_close(res[k].float().cpu(), res_ref[k].detach(), rtol=5e-3, atol=5e-3, msg=lambda m: f"{k}: {m}")
    loss = O.nerf_loss({k: v.float() for k, v in res.items() if torch.is_tensor(v) and v.ndim > 0}, target.to(DEV))
    assert abs(loss.item() - loss_ref.item()) < 2e-3 * abs(loss_ref.item())
    (loss * 1024.0).backward()
    for got, want, name in ((model.rgb_net.params.grad, ref.rgb_params.grad, "rgb_net"),
                            (model.xyz_encoder.params.grad, ref.xyz_params.grad, "xyz_encoder")):
        sc = want.abs().max().item()
        assert (got.cpu() - want).abs().max().item() <= 2e-3 * sc, name
    model.zero_grad(); ref.xyz_params.grad = None; ref.rgb_params.grad = None


def test_render_test_time(pair):
    from google_nerf_b200.models.rendering import render
    from oracle import ngp_ref as O
    ref, model, s = pair
    n = 256
    res_ref = O.render(ref, s["rays_o"][:n], s["rays_d"][:n].clone(), test_time=True, T_threshold=1e-2)
    res = render(model, s["rays_o"][:n].to(DEV), s["rays_d"][:n].to(DEV).clone(), test_time=True, T_threshold=1e-2,
                 whole_rays=False)                           # the reference's round schedule (sample count included)
    # fp16 field noise can move a ray's early-termination by a sample, so counts agree to within a few samples
    assert abs(int(res["total_samples"]) - res_ref["total_samples"]) <= 0.01 * res_ref["total_samples"] + 2
    for k in ("opacity", "depth", "rgb"):
        torch.testing.assert_close(res[k].cpu(), res_ref[k], rtol=1e-2, atol=1e-2, msg=lambda m: f"{k}: {m}")
    cpu = render(model, s["rays_o"][:n].to(DEV), s["rays_d"][:n].to(DEV).clone(), test_time=True, to_cpu=True)
    assert not cpu["rgb"].is_cuda


@pytest.mark.parametrize("esf,thr", [(0.0, 1e-2), (0.0, 1e-4), (1.0 / 256, 1e-3)])
def test_render_device_loop_equals_host_loop(pair, esf, thr):
    """The graph-replayed device-driven loop (alive list, ray counts and samples per ray on the GPU) produces the very
    same image and sample count as the reference-style host loop of rendering.py:42-114, also when called again on
    other rays (buffers and graph are reused)."""
    from google_nerf_b200.models.rendering import render
    ref, model, s = pair
    for lo, hi in ((0, 700), (68, 768)):
        ro, rd = s["rays_o"][lo:hi].to(DEV), s["rays_d"][lo:hi].to(DEV)
        a = render(model, ro, rd.clone(), test_time=True, T_threshold=thr, exp_step_factor=esf, device_loop=False)
        b = render(model, ro, rd.clone(), test_time=True, T_threshold=thr, exp_step_factor=esf, whole_rays=False)
        assert int(a["total_samples"]) == int(b["total_samples"]) > 0
        for k in ("opacity", "depth", "rgb"):
            assert torch.equal(a[k], b[k]), k


def test_frequency_variant_forward(pair):
    from google_nerf_b200.models.networks import NGP
    from oracle import ngp_ref as O
    _, _, s = pair
    ref = O.NGPRef(0.5, encoding="Frequency", seed=7)
    model = NGP(0.5, encoding="Frequency").to(DEV)
    model.xyz_encoder.params.data.copy_(ref.xyz_params.detach())
    model.rgb_net.params.data.copy_(ref.rgb_params.detach())
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(2000, 3, generator=g) - 0.5); d = torch.randn(2000, 3, generator=g)
    with torch.no_grad():
        sr, cr = ref(x, d.clone())
        sg, cg = model(x.to(DEV), d.to(DEV).clone())
    torch.testing.assert_close(sg.cpu(), sr, rtol=2e-2, atol=2e-3)
    torch.testing.assert_close(cg.float().cpu(), cr.float(), rtol=1e-2, atol=5e-3)


def test_update_density_grid_consistency(pair):
    _, model, _ = pair
    from google_nerf_b200.models.networks import NGP
    m = NGP(0.5, log2_T=15).to(DEV).init_grid_buffers()
    m.xyz_encoder.params.data.copy_(model.xyz_encoder.params.data)
    m.density_grid[0, :1000] = -1.0                                      # "invisible" cells are never updated
    m.update_density_grid(5.912, warmup=True)
    grid = m.density_grid
    assert float(grid[0, :1000].max()) == -1.0 and float(grid[0, 1000:].min()) >= 0.0
    mean = grid[grid > 0].mean().item()
    thr = min(mean, 5.912)
    assert abs(m._grid_stats[0].item() - thr) < 1e-5 * max(1.0, thr)
    ref_bits = np.packbits((grid.cpu().numpy().reshape(-1) > m._grid_stats[0].item()), bitorder="little")
    assert np.array_equal(m.density_bitfield.cpu().numpy(), ref_bits)
    # the device-side occupied-cell sampler only returns occupied cells, roughly uniformly over them
    # (the returned list is sorted by Morton index, so the two halves cannot be told apart afterwards: at least the M
    # occupied draws must land on occupied cells, plus the uniform draws' share)
    (idx_all, coords_all), = m.sample_uniform_and_occupied_cells(20000)
    from google_nerf_b200 import vren
    assert idx_all.numel() == 40000 and torch.equal(vren.morton3D(coords_all).long(), idx_all)
    n_occ = int((grid[0] > 0).sum())
    assert 0 < n_occ < 128 ** 3
    hits = int((grid[0, idx_all] > 0).sum())
    expect = 20000 + 20000 * n_occ / 128 ** 3
    assert abs(hits - expect) < 6 * (20000 * 0.25) ** 0.5 + 1, (hits, expect)
    # and the occupied draws spread over the occupied cells instead of repeating a few
    occ_hit = idx_all[grid[0, idx_all] > 0]
    assert occ_hit.unique().numel() > 0.5 * min(hits, n_occ)
    before = grid.clone()
    m.update_density_grid(5.912, warmup=False)                           # uniform + occupied sampling path
    assert float((m.density_grid - before).abs().max()) > 0
    assert float(m.density_grid[0, :1000].max()) == -1.0


def test_trainer_matches_oracle_step(built_lib):
    import __graft_entry__ as g
    g.smoke()


def test_trainer_graph_equals_eager_and_learns(pair):
    """CUDA-graph replay == eager launch sequence (same seeds), and the loss goes down on the analytic scene."""
    from google_nerf_b200 import synthetic as syn
    from google_nerf_b200.models.networks import NGP
    from google_nerf_b200.trainer import NGPTrainer
    _, _, s = pair
    n = 768
    ro, rd = s["rays_o"][:n].to(DEV), s["rays_d"][:n].to(DEV)
    tgt = syn.shade(ro, rd, 0.5)
    losses = {}
    for use_graph in (False, True):
        torch.manual_seed(0)
        m = NGP(0.5, log2_T=15).to(DEV)
        m.density_bitfield.copy_(s["bitfield"])
        tr = NGPTrainer(m, n_rays=n, use_graph=use_graph, samples_per_ray=200, grid_update_interval=10 ** 9, seed=3)
        tr.step_count = 1                                               # keep the analytic bitfield
        tr.fixed_noise = s["noise"][:n].to(DEV)
        ls = []
        for _ in range(30):
            ls.append(float(tr.step(ro, rd, tgt).item()))
        assert not tr.overflowed()
        losses[use_graph] = ls
    assert losses[True][-1] < 0.5 * losses[True][0]
    np.testing.assert_allclose(losses[True], losses[False], rtol=2e-2)


def test_trainer_pipelined_march_matches_sequential(pair):
    """Handing over the next batch (marched on the side stream while the current step trains) gives the same
    training trajectory as marching every batch inline: the samples depend only on the bitfield and the jitter."""
    from google_nerf_b200 import synthetic as syn
    from google_nerf_b200.models.networks import NGP
    from google_nerf_b200.trainer import NGPTrainer
    _, _, s = pair
    n, steps = 384, 24
    batches = []
    for k in range(steps + 1):
        ro, rd = s["rays_o"][(k % 2) * n:(k % 2 + 1) * n].to(DEV), s["rays_d"][(k % 2) * n:(k % 2 + 1) * n].to(DEV)
        batches.append((ro, rd, syn.shade(ro, rd, 0.5)))
    out = {}
    for pipelined in (False, True):
        torch.manual_seed(0)
        m = NGP(0.5, log2_T=15).to(DEV)
        m.density_bitfield.copy_(s["bitfield"])
        tr = NGPTrainer(m, n_rays=n, use_graph=True, samples_per_ray=200, grid_update_interval=10 ** 9, seed=3)
        tr.step_count = 1
        tr.fixed_noise = s["noise"][:n].to(DEV)
        ls = []
        for k in range(steps):
            if pipelined:
                cur = batches[0] if k == 0 else (None, None, None)
                ls.append(float(tr.step(*cur, next_batch=batches[k + 1]).item()))
            else:
                ls.append(float(tr.step(*batches[k]).item()))
        out[pipelined] = ls
    np.testing.assert_allclose(out[True], out[False], rtol=3e-2)
    assert out[True][-1] < 0.6 * out[True][0]


def test_trainer_unbounded_config_matches_oracle_step(built_lib):
    """BASELINE config 5 shape in miniature: scale 4 (3 cascades), exp_step_factor 1/256, black background.  One fused
    training step (march with exponential stepping through the cascades, field, compositing, backward) against the
    oracle on the same weights / bitfield / jitter."""
    from google_nerf_b200 import synthetic as syn
    from google_nerf_b200.models.networks import NGP
    from google_nerf_b200.trainer import NGPTrainer
    from oracle import ngp_ref as O
    scale, n = 4.0, 384
    s = make_scene(scale, n, seed=13)
    ref = O.NGPRef(scale, log2_T=14, seed=3)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        ref.xyz_params[ref.n_mlp:] = (torch.rand(ref.layout["n_params"], generator=g) * 2 - 1) * 0.3
    ref.density_bitfield = s["bitfield"].clone()
    target = torch.rand(n, 3, generator=g)
    model = NGP(scale, log2_T=14).to(DEV)
    assert model.cascades == 4
    model.xyz_encoder.params.data.copy_(ref.xyz_params.detach()); model.rgb_net.params.data.copy_(ref.rgb_params.detach())
    model.density_bitfield.copy_(s["bitfield"])
    tr = NGPTrainer(model, n_rays=n, exp_step_factor=1 / 256, use_graph=False, samples_per_ray=400,
                    grid_update_interval=10 ** 9)
    tr.step_count = 1
    tr.fixed_noise = s["noise"].to(DEV)
    tr.set_batch(s["rays_o"].to(DEV), s["rays_d"].to(DEV), target.to(DEV))
    sset = tr.sets[tr.cur]
    tr._march(sset); tr._forward_backward(sset); tr.last_counter = sset.counter
    res = O.render(ref, s["rays_o"], s["rays_d"].clone(), noise=s["noise"], exp_step_factor=1 / 256)
    loss = O.nerf_loss(res, target)
    (loss * tr.loss_scale).backward()
    assert tr.samples_last_step() == res["total_samples"] > 1000 and not tr.overflowed()
    torch.testing.assert_close(tr.opacity.cpu(), res["opacity"].detach(), rtol=5e-3, atol=5e-3)
    torch.testing.assert_close(tr.rgb_out.cpu(), res["rgb"].detach(), rtol=5e-3, atol=5e-3)
    assert abs(tr.loss.item() - loss.item()) < 3e-3 * abs(loss.item())
    for got, want, name in ((tr.g_rgb, ref.rgb_params.grad, "rgb_net"), (tr.g_xyz, ref.xyz_params.grad, "xyz_encoder")):
        sc = want.abs().max().item()
        assert (got.cpu() - want).abs().max().item() <= 1e-2 * sc, name


@pytest.mark.parametrize("esf", [0.0, 1.0 / 256])
def test_render_graph_form_equals_launch_by_launch(pair, esf):
    """render() at training time: the single-node form (two CUDA-graph replays: AABB, march, field, compositing, blend
    forward; compositing / field / scatter backward) gives the results and parameter gradients of the launch-by-launch
    form with one autograd node per reference Function -- on the first call (capture) and on replays with new rays."""
    from google_nerf_b200.models.rendering import render
    from google_nerf_b200.models.custom_functions import RayMarcher
    _, model, s = pair
    g = torch.Generator().manual_seed(4)
    model.zero_grad()
    try:
        for lo in (0, 256, 0):                                    # third pass: a replay on the first rays again
            ro, rd = s["rays_o"][lo:lo + 512].to(DEV), s["rays_d"][lo:lo + 512].to(DEV)
            RayMarcher.noise = s["noise"][lo:lo + 512].to(DEV)
            w = [torch.rand(512, generator=g).to(DEV) for _ in range(3)] + [torch.rand(512, 3, generator=g).to(DEV)]
            out = {}
            for graph in (False, True):
                res = render(model, ro, rd.clone(), exp_step_factor=esf, graph=graph)
                loss = sum((res[k] * wk).sum() for k, wk in zip(("opacity", "depth", "depth_sq", "rgb"), w))
                model.zero_grad()
                (loss * 64.0).backward()
                out[graph] = (res, model.xyz_encoder.params.grad.clone(), model.rgb_net.params.grad.clone())
            a, b = out[False], out[True]
            assert int(a[0]["total_samples"]) == int(b[0]["total_samples"]) > 0
            for k in ("opacity", "depth", "depth_sq", "rgb"):
                torch.testing.assert_close(b[0][k], a[0][k].float(), rtol=1e-6, atol=1e-6)
            for ga, gb in zip(a[1:], b[1:]):
                assert (ga - gb).abs().max().item() <= 2e-4 * ga.abs().max().item()   # atomics order only
    finally:
        RayMarcher.noise = None
        model.zero_grad()


def test_render_full_frame_device_loop_equals_host_loop(built_lib):
    """BASELINE config 3 at full size: an 800x800 frame (640,000 rays; the marcher's thread-per-ray path above 16,384
    live rays, the warp path below) rendered by the device-driven loop and by the reference-style host loop -- the
    same image, depth, opacity and sample count, bit for bit."""
    from google_nerf_b200 import synthetic as syn
    from google_nerf_b200.models.networks import NGP
    from google_nerf_b200.models.rendering import render
    torch.manual_seed(0)
    model = NGP(0.5, log2_T=16).to(DEV)
    g = torch.Generator().manual_seed(3)
    model.xyz_encoder.params.data[model.xyz_encoder.mlp.n_params:] = \
        ((torch.rand(model.xyz_encoder.enc.n_params, generator=g) * 2 - 1) * 0.5).to(DEV)
    model.density_bitfield.copy_(syn.bitfield_from_grid(syn.density_grid(0.5, 1)).to(DEV))
    K = syn.intrinsics(800, 800); dirs = syn.directions(800, 800, K).to(DEV)
    ro, rd = syn.get_rays(dirs, syn.hemisphere_poses(3)[1].to(DEV))
    with torch.no_grad():
        a = render(model, ro, rd.clone(), test_time=True, T_threshold=1e-2, device_loop=False)
        b = render(model, ro, rd.clone(), test_time=True, T_threshold=1e-2, whole_rays=False)
    assert int(a["total_samples"]) == int(b["total_samples"]) > 640000
    for k in ("opacity", "depth", "rgb"):
        assert torch.equal(a[k], b[k]), k
    assert float(b["opacity"].max()) > 0.1 and float(b["opacity"].min()) == 0.0      # hit and missed rays both present


def test_render_test_time_frequency_encoding(pair):
    """Test-time render of the fork's active Frequency-12 configuration (device-driven loop on the K1 = 80 kernels)
    against the oracle's host loop."""
    from google_nerf_b200.models.networks import NGP
    from google_nerf_b200.models.rendering import render
    from oracle import ngp_ref as O
    _, _, s = pair
    ref = O.NGPRef(0.5, encoding="Frequency", seed=7)
    ref.density_bitfield = s["bitfield"].clone()
    model = NGP(0.5, encoding="Frequency").to(DEV)
    model.xyz_encoder.params.data.copy_(ref.xyz_params.detach()); model.rgb_net.params.data.copy_(ref.rgb_params.detach())
    model.density_bitfield.copy_(s["bitfield"])
    n = 256
    res_ref = O.render(ref, s["rays_o"][:n], s["rays_d"][:n].clone(), test_time=True, T_threshold=1e-2)
    res = render(model, s["rays_o"][:n].to(DEV), s["rays_d"][:n].to(DEV).clone(), test_time=True, T_threshold=1e-2)
    assert abs(int(res["total_samples"]) - res_ref["total_samples"]) <= 0.01 * res_ref["total_samples"] + 2
    for k in ("opacity", "depth", "rgb"):
        torch.testing.assert_close(res[k].cpu(), res_ref[k], rtol=1e-2, atol=1e-2, msg=lambda m: f"{k}: {m}")
    host = render(model, s["rays_o"][:n].to(DEV), s["rays_d"][:n].to(DEV).clone(), test_time=True, T_threshold=1e-2,
                  device_loop=False)
    for k in ("opacity", "depth", "rgb"):
        assert torch.equal(host[k], res[k]), k
