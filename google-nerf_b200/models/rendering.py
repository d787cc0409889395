"""render() with the signature, kwargs and result keys of ngp_pl/models/rendering.py:12-166.

kwargs: test_time (False), exp_step_factor (0.), T_threshold (1e-4), to_cpu (False).
results: opacity, depth, rgb, total_samples (+ depth_sq when training)."""
import torch

from .. import vren
from .custom_functions import RayAABBIntersector, RayMarcher, VolumeRenderer

MAX_SAMPLES = 1024
NEAR_DISTANCE = 0.05


def render(model, rays_o, rays_d, **kwargs):
    """rays_o, rays_d (N_rays,3) -> dict.  AABB-clip, clamp the near hit to NEAR_DISTANCE, then train- or
    test-time rendering (rendering.py:26-39)."""
    rays_o = rays_o.contiguous(); rays_d = rays_d.contiguous()
    _, hits_t, _ = RayAABBIntersector.apply(rays_o, rays_d, model.center, model.half_size, 1)
    near = hits_t[:, 0, 0]
    hits_t[(near >= 0) & (near < NEAR_DISTANCE), 0, 0] = NEAR_DISTANCE

    fn = _render_rays_test if kwargs.get("test_time", False) else _render_rays_train
    results = fn(model, rays_o, rays_d, hits_t, **kwargs)
    if kwargs.get("to_cpu", False):
        results = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in results.items()}
    return results


def _background(exp_step_factor, device):
    # synthetic scenes (exp_step_factor == 0) composite over white, real scenes over black (rendering.py:108-111)
    return torch.ones(3, device=device) if exp_step_factor == 0 else torch.zeros(3, device=device)


@torch.no_grad()
def _render_rays_test(model, rays_o, rays_d, hits_t, **kwargs):
    """Iterative alive-ray marching (rendering.py:42-114): every pass marches each live ray to its next
    N_samples occupied samples, evaluates the field there and composites in place; converged rays drop out."""
    exp_step_factor = kwargs.get("exp_step_factor", 0.)
    T_threshold = kwargs.get("T_threshold", 1e-4)
    N_rays, device = len(rays_o), rays_o.device
    opacity = torch.zeros(N_rays, device=device)
    depth = torch.zeros(N_rays, device=device)
    rgb = torch.zeros(N_rays, 3, device=device)
    hits = hits_t[:, 0]                                   # (N_rays,2) view, advanced in place by the marcher

    samples = total_samples = 0
    fused = getattr(model, "encoding", None) == "HashGrid" and hasattr(model, "_forward_fused")
    alive_indices = torch.arange(N_rays, device=device)
    min_samples = 1 if exp_step_factor == 0 else 4
    while samples < MAX_SAMPLES:
        N_alive = len(alive_indices)
        if N_alive == 0:
            break
        N_samples = max(min(N_rays // N_alive, 64), min_samples)
        samples += N_samples
        xyzs, dirs, deltas, ts, N_eff_samples = vren.raymarching_test(
            rays_o, rays_d, hits, alive_indices, model.density_bitfield, model.cascades, model.scale,
            exp_step_factor, model.grid_size, MAX_SAMPLES, N_samples)
        total_samples += N_eff_samples.sum()
        xyzs = xyzs.view(-1, 3); dirs = dirs.view(-1, 3)
        if fused:
            # Mask-free variant of rendering.py:87-95: the field is evaluated on every slot in one fused launch;
            # slots beyond N_eff_samples are never read by composite_test_fw, so the results are identical and the
            # boolean-mask gathers/scatters (and their host syncs) disappear.  A round without any sample marks
            # every ray dead in composite_test_fw, which ends the loop exactly like the reference's early break.
            sigmas, rgbs = model._forward_fused(xyzs, dirs, rgb_fp32=True)
        else:
            valid_mask = ~torch.all(dirs == 0, dim=1)
            if valid_mask.sum() == 0:
                break
            sigmas = torch.zeros(len(xyzs), device=device)
            rgbs = torch.zeros(len(xyzs), 3, device=device)
            _sigmas, _rgbs = model(xyzs[valid_mask], dirs[valid_mask])
            sigmas[valid_mask], rgbs[valid_mask] = _sigmas.float(), _rgbs.float()
        vren.composite_test_fw(sigmas.view(-1, N_samples), rgbs.view(-1, N_samples, 3), deltas, ts, hits,
                               alive_indices, T_threshold, N_eff_samples, opacity, depth, rgb)
        alive_indices = alive_indices[alive_indices >= 0]

    rgb = rgb + _background(exp_step_factor, device) * (1 - opacity)[:, None]
    return {"opacity": opacity, "depth": depth, "rgb": rgb, "total_samples": total_samples}


def _render_rays_train(model, rays_o, rays_d, hits_t, **kwargs):
    """march -> field -> composite under autocast (rendering.py:117-166)."""
    exp_step_factor = kwargs.get("exp_step_factor", 0.)
    with torch.autocast("cuda", dtype=torch.float16):
        rays_a, xyzs, dirs, deltas, ts, total_samples = RayMarcher.apply(
            rays_o, rays_d, hits_t[:, 0], model.density_bitfield, model.cascades, model.scale, exp_step_factor,
            model.grid_size, MAX_SAMPLES)
        sigmas, rgbs = model(xyzs, dirs)
        opacity, depth, depth_sq, rgb = VolumeRenderer.apply(sigmas, rgbs.contiguous(), deltas, ts, rays_a,
                                                             kwargs.get("T_threshold", 1e-4))
        rgb = rgb + _background(exp_step_factor, rays_o.device) * (1 - opacity)[:, None]
    return {"total_samples": total_samples, "opacity": opacity, "depth": depth, "depth_sq": depth_sq, "rgb": rgb}


# the reference spells these with a leading double underscore (module-private names)
globals()["__render_rays_test"] = _render_rays_test
globals()["__render_rays_train"] = _render_rays_train
