"""Whole-ray test-time renderer (csrc/render_tc.cu, b2n_render_rays): one persistent kernel per frame instead of the
round loop of rendering.py:42-114.  Checked against (a) the marcher on its own (sample counts per ray are exact),
(b) the round loop on the same kernels (images agree to fp32 rounding: only the points where T is re-derived from the
opacity differ), (c) the oracle's host loop, (d) the sample-budget fallback."""
import pytest
import torch

from conftest import make_scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(scale=0.5, log2_T=16, seed=3, table_amp=0.5, s=None):
    from google_nerf_b200 import synthetic as syn
    from google_nerf_b200.models.networks import NGP
    torch.manual_seed(0)
    model = NGP(scale, log2_T=log2_T).to(DEV)
    g = torch.Generator().manual_seed(seed)
    model.xyz_encoder.params.data[model.xyz_encoder.mlp.n_params:] = \
        ((torch.rand(model.xyz_encoder.enc.n_params, generator=g) * 2 - 1) * table_amp).to(DEV)
    cascades = model.cascades
    model.density_bitfield.copy_(syn.bitfield_from_grid(syn.density_grid(scale, cascades)).to(DEV))
    return model


def _close(a, b, thr, what):
    """fp32-rounding agreement; a ray whose transmittance lands within rounding of T_threshold may composite one sample
    more or less in one of the two schedules (weight <= T_threshold), so a handful of rays get that much slack."""
    d = (a - b).abs()
    tol = 1e-5 * (1 + b.abs())
    bad = d > tol
    assert float(bad.float().mean()) < 2e-3, (what, float(bad.float().mean()), float(d.max()))
    lim = 4.0 * thr * (8.0 if what == "depth" else 1.0) + 1e-5
    assert float(d.max()) <= lim, (what, float(d.max()))


@pytest.mark.parametrize("esf,scale", [(0.0, 0.5), (1.0 / 256, 4.0)])
def test_whole_rays_march_counts_are_exact(built_lib, esf, scale):
    """T_threshold = 0 on a thin medium: no ray retires early, so the samples a ray marched inside the fused kernel
    are all the samples the stand-alone test marcher finds for it (b2n_raymarching_test with room for 1024)."""
    from google_nerf_b200 import _lib as L
    from google_nerf_b200.models.rendering import _WholeRays, MAX_SAMPLES, NEAR_DISTANCE
    from google_nerf_b200.models.custom_functions import RayAABBIntersector
    model = _model(scale, log2_T=15, table_amp=1e-4)
    s = make_scene(scale, 3000, seed=5)
    ro, rd = s["rays_o"].to(DEV), s["rays_d"].to(DEV)
    _, hits_t, _ = RayAABBIntersector.apply(ro, rd, model.center, model.half_size, 1)
    hits_t[(hits_t[:, 0, 0] >= 0) & (hits_t[:, 0, 0] < NEAR_DISTANCE), 0, 0] = NEAR_DISTANCE
    hits = hits_t[:, 0].contiguous()
    n = len(ro)
    counts = torch.full((n,), -1, dtype=torch.int32, device=DEV)
    wr = _WholeRays.get(model, n)
    res = wr.run(ro, rd, hits, esf, 0.0, ray_samples=counts)
    P = L.ptr
    alive = torch.arange(n, device=DEV)
    N = MAX_SAMPLES
    xyzs = torch.empty(n * N, 3, device=DEV); dirs = torch.empty(n * N, 3, device=DEV)
    deltas = torch.empty(n * N, device=DEV); ts = torch.empty(n * N, device=DEV)
    n_eff = torch.empty(n, dtype=torch.int32, device=DEV)
    h2 = hits.clone()
    L.call("b2n_raymarching_test", P(ro), P(rd), P(h2), P(alive), P(model.density_bitfield), model.cascades,
           float(model.scale), float(esf), model.grid_size, MAX_SAMPLES, N, n, P(xyzs), P(dirs), P(deltas), P(ts), P(n_eff))
    assert int(n_eff.max()) > 50 and int((n_eff == 0).sum()) > 0
    assert torch.equal(counts, n_eff)
    if res is None:                                          # rays that filled the budget: the caller falls back
        assert int((n_eff >= MAX_SAMPLES).sum()) > 0
    else:
        assert int((n_eff >= MAX_SAMPLES).sum()) == 0 and int(res["total_samples"]) == int(n_eff.sum())


@pytest.mark.parametrize("esf,scale,thr", [(0.0, 0.5, 1e-2), (0.0, 0.5, 1e-4), (1.0 / 256, 4.0, 1e-3)])
def test_whole_rays_equal_round_loop(built_lib, esf, scale, thr):
    from google_nerf_b200.models.rendering import render
    model = _model(scale, log2_T=16)
    s = make_scene(scale, 20000, seed=9)
    ro, rd = s["rays_o"].to(DEV), s["rays_d"].to(DEV)
    with torch.no_grad():
        a = render(model, ro, rd.clone(), test_time=True, T_threshold=thr, exp_step_factor=esf, whole_rays=False)
        b = render(model, ro, rd.clone(), test_time=True, T_threshold=thr, exp_step_factor=esf, whole_rays=True)
        c = render(model, ro, rd.clone(), test_time=True, T_threshold=thr, exp_step_factor=esf, whole_rays=True)
    assert model._whole_rays.rounds > 0                      # the fused kernel ran (and did not fall back)
    for k in ("opacity", "depth", "rgb"):
        _close(b[k], a[k], thr, k)
        assert torch.equal(b[k], c[k]), k                    # deterministic
    assert 0 < int(b["total_samples"]) <= int(a["total_samples"]) * 1.5
    assert float(b["opacity"].max()) > 0.1 and float(b["opacity"].min()) == 0.0


def test_whole_rays_full_frame_and_ragged_sizes(built_lib):
    """An 800x800 frame (more rays than ray slots: the queue refills the CTAs) and sizes that do not fill a CTA."""
    from google_nerf_b200 import synthetic as syn
    from google_nerf_b200.models.rendering import render
    model = _model(0.5, log2_T=16)
    K = syn.intrinsics(800, 800); dirs = syn.directions(800, 800, K).to(DEV)
    ro, rd = syn.get_rays(dirs, syn.hemisphere_poses(3)[1].to(DEV))
    with torch.no_grad():
        a = render(model, ro, rd.clone(), test_time=True, T_threshold=1e-2, whole_rays=False)
        b = render(model, ro, rd.clone(), test_time=True, T_threshold=1e-2, whole_rays=True)
        for k in ("opacity", "depth", "rgb"):
            _close(b[k], a[k], 1e-2, k)
        for n in (1, 127, 129, 1000):
            a = render(model, ro[5000:5000 + n], rd[5000:5000 + n].clone(), test_time=True, T_threshold=1e-2, whole_rays=False)
            b = render(model, ro[5000:5000 + n], rd[5000:5000 + n].clone(), test_time=True, T_threshold=1e-2, whole_rays=True)
            for k in ("opacity", "depth", "rgb"):
                _close(b[k], a[k], 1e-2, k)


def test_whole_rays_against_oracle(built_lib):
    from google_nerf_b200.models.networks import NGP
    from google_nerf_b200.models.rendering import render
    from oracle import ngp_ref as O
    s = make_scene(0.5, 256, seed=11)
    ref = O.NGPRef(0.5, log2_T=15, seed=1337)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        ref.xyz_params[ref.n_mlp:] = (torch.rand(ref.layout["n_params"], generator=g) * 2 - 1) * 0.5
    ref.density_bitfield = s["bitfield"].clone()
    model = NGP(0.5, log2_T=15).to(DEV)
    model.xyz_encoder.params.data.copy_(ref.xyz_params.detach()); model.rgb_net.params.data.copy_(ref.rgb_params.detach())
    model.density_bitfield.copy_(s["bitfield"])
    res_ref = O.render(ref, s["rays_o"], s["rays_d"].clone(), test_time=True, T_threshold=1e-2)
    res = render(model, s["rays_o"].to(DEV), s["rays_d"].to(DEV).clone(), test_time=True, T_threshold=1e-2, whole_rays=True)
    assert model._whole_rays.rounds > 0
    for k in ("opacity", "depth", "rgb"):
        torch.testing.assert_close(res[k].cpu(), res_ref[k], rtol=1e-2, atol=1e-2, msg=lambda m: f"{k}: {m}")


def test_whole_rays_budget_fallback(built_lib):
    """Fully occupied grid, a medium that never saturates (T_threshold = 0) and cameras inside a scale-16 box at
    exp_step_factor 1/256: every ray has more than MAX_SAMPLES samples (226 at the minimum step up to t = 0.43, then
    256 ln(16 / 0.43)), the fused kernel reports them and render() returns the round loop's result."""
    from google_nerf_b200.models.rendering import render
    model = _model(16.0, log2_T=15, table_amp=1e-4)
    model.density_bitfield.fill_(255)
    g = torch.Generator().manual_seed(2)
    ro = ((torch.rand(1500, 3, generator=g) - 0.5) * 0.2).to(DEV)
    rd = torch.nn.functional.normalize(torch.randn(1500, 3, generator=g), dim=1).to(DEV)
    with torch.no_grad():
        a = render(model, ro, rd.clone(), test_time=True, T_threshold=0.0, exp_step_factor=1 / 256, whole_rays=False)
        b = render(model, ro, rd.clone(), test_time=True, T_threshold=0.0, exp_step_factor=1 / 256, whole_rays=True)
    assert int(model._whole_rays.ctl_host[1]) > 0
    for k in ("opacity", "depth", "rgb"):
        assert torch.equal(a[k], b[k]), k
    assert int(a["total_samples"]) == int(b["total_samples"]) >= 1024 * 1000
