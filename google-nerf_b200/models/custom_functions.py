"""Autograd wrappers with the names, argument order and return arity of
ngp_pl/models/custom_functions.py (RayAABBIntersector :8-29, RaySphereIntersector :32-52, RayMarcher :55-113,
VolumeRenderer :116-159, TruncExp :162-173), bound to libb2n through the `vren` drop-in."""
import torch
from torch.amp import custom_bwd, custom_fwd

from .. import vren

_fwd32 = custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = custom_bwd(device_type="cuda")


class RayAABBIntersector(torch.autograd.Function):
    """rays_o, rays_d (N_rays,3); centers, half_sizes (N_voxels,3); max_hits ->
    hits_cnt (N_rays), hits_t (N_rays,max_hits,2) near-to-far (-1 = no hit), hits_voxel_idx (N_rays,max_hits)."""

    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, center, half_size, max_hits):
        return vren.ray_aabb_intersect(rays_o, rays_d, center, half_size, max_hits)


class RaySphereIntersector(torch.autograd.Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, center, radii, max_hits):
        return vren.ray_sphere_intersect(rays_o, rays_d, center, radii, max_hits)


class RayMarcher(torch.autograd.Function):
    """-> rays_a (N_rays,3) [ray_idx, start_idx, N_samples], xyzs, dirs (N,3), deltas, ts (N), total_samples.

    ``RayMarcher.noise`` may be set to a (N_rays) tensor to fix the per-ray jitter (parity tests share it with
    the oracle); otherwise it is drawn with torch.rand_like as at custom_functions.py:84.

    ``RayMarcher.sync_free`` (default True): the reference reads the sample count back to the host to size its outputs
    (custom_functions.py:92-97), which stalls the host once per step.  Here the first call does the same; later calls
    size the buffers from the previous step's count with 4x headroom (capped at N_rays * max_samples, the reference's
    own allocation), keep the count on the device -- it travels to the field kernels as the ``_b2n_n_dev`` attribute of
    ``xyzs`` -- and return ``total_samples`` as a device scalar.  Rows past the count are never read.  The count of a
    step is checked one step later; a step whose samples did not fit is truncated ray-wise by the marcher and reported
    with a warning (it takes a 4x jump of the sample count between two consecutive steps)."""
    noise = None
    sync_free = True
    _state = {}

    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, grid_size,
                max_samples):
        noise = RayMarcher.noise if RayMarcher.noise is not None else torch.rand_like(rays_o[:, 0])
        args = (rays_o.contiguous(), rays_d.contiguous(), hits_t.contiguous(), density_bitfield, cascades, scale,
                exp_step_factor, noise, grid_size, max_samples)
        n = rays_o.shape[0]
        st = RayMarcher._state.setdefault((rays_o.device, n), {"last": None, "pending": False})
        if st["pending"] and st["event"].query():
            st["pending"] = False
            st["last"] = int(st["host"][3])
            if int(st["host"][2]):
                import warnings
                warnings.warn("RayMarcher: a step produced more than 4x the samples of the step before; its rays were "
                              "truncated to the buffer (set RayMarcher.sync_free = False for exact sizing)")
        if not RayMarcher.sync_free or st["last"] is None or n == 0:
            rays_a, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(*args)
            st["last"] = int(xyzs.shape[0])
        else:
            cap = min(n * int(max_samples), max(4 * st["last"], 16 * n))
            rays_a, counter, ws = vren.raymarching_train_count(*args, capacity=cap)
            xyzs, dirs, deltas, ts = vren.raymarching_train_write(*args, rays_a, cap, ws)
            xyzs._b2n_n_dev = counter                       # device-side sample count for the field kernels
            if not st["pending"]:
                if "host" not in st:
                    st["host"], st["event"] = torch.zeros(4, dtype=torch.int32).pin_memory(), torch.cuda.Event()
                st["host"].copy_(counter, non_blocking=True)
                st["event"].record()
                st["pending"] = True
        total_samples = counter[0]
        ctx.save_for_backward(rays_a, ts)
        return rays_a, xyzs, dirs, deltas, ts, total_samples

    @staticmethod
    @_bwd
    def backward(ctx, dL_drays_a, dL_dxyzs, dL_ddirs, dL_ddeltas, dL_dts, dL_dtotal_samples):
        # per-ray segmented sums (the reference uses torch_scatter.segment_csr, custom_functions.py:108-111) in one
        # warp-per-ray launch
        rays_a, ts = ctx.saved_tensors
        n_rays = rays_a.shape[0]
        from .. import _lib as L
        g_x = dL_dxyzs.float().contiguous()
        g_d = dL_ddirs.float().contiguous() if dL_ddirs is not None else None
        dL_drays_o = torch.zeros(n_rays, 3, device=ts.device); dL_drays_d = torch.zeros(n_rays, 3, device=ts.device)
        L.call("b2n_raymarcher_bw", L.ptr(g_x), L.ptr(g_d), L.ptr(ts), L.ptr(rays_a), n_rays, L.ptr(dL_drays_o),
               L.ptr(dL_drays_d))
        return dL_drays_o, dL_drays_d, None, None, None, None, None, None, None


class VolumeRenderer(torch.autograd.Function):
    """sigmas (N), rgbs (N,3), deltas, ts (N), rays_a (N_rays,3), T_threshold ->
    opacity, depth, depth_sq (N_rays), rgb (N_rays,3)."""

    @staticmethod
    @_fwd32
    def forward(ctx, sigmas, rgbs, deltas, ts, rays_a, T_threshold):
        opacity, depth, depth_sq, rgb = vren.composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T_threshold)
        ctx.save_for_backward(sigmas, rgbs, deltas, ts, rays_a, opacity, depth, depth_sq, rgb)
        ctx.T_threshold = T_threshold
        return opacity, depth, depth_sq, rgb

    @staticmethod
    @_bwd
    def backward(ctx, dL_dopacity, dL_ddepth, dL_ddepth_sq, dL_drgb):
        sigmas, rgbs, deltas, ts, rays_a, opacity, depth, depth_sq, rgb = ctx.saved_tensors
        dL_dsigmas, dL_drgbs = vren.composite_train_bw(
            dL_dopacity.contiguous(), dL_ddepth.contiguous(), dL_ddepth_sq.contiguous(), dL_drgb.contiguous(),
            sigmas, rgbs, deltas, ts, rays_a, opacity, depth, depth_sq, rgb, ctx.T_threshold)
        return dL_dsigmas, dL_drgbs, None, None, None, None


class TruncExp(torch.autograd.Function):
    """exp forward; backward uses exp(clamp(x, -15, 15))  (custom_functions.py:162-173)."""

    @staticmethod
    @_fwd32
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @_bwd
    def backward(ctx, dL_dout):
        x = ctx.saved_tensors[0]
        return dL_dout * torch.exp(x.clamp(-15, 15))
