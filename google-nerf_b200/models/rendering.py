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
    if hits_t.is_cuda:       # hits_t[(t1 >= 0) & (t1 < NEAR), 0, 0] = NEAR (rendering.py:29) without the boolean-mask sync
        from .. import _lib as L
        L.call("b2n_clamp_near", L.ptr(hits_t), hits_t.shape[0], NEAR_DISTANCE)
    else:
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
    fused = hasattr(model, "_forward_fused")            # this repo's NGP (HashGrid or Frequency): fused field kernels
    if fused and kwargs.get("device_loop", True) and not torch.cuda.is_current_stream_capturing():
        return _DeviceLoop.get(model, N_rays, exp_step_factor, T_threshold).run(rays_o, rays_d, hits)
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


class _DeviceLoop:
    """The loop of rendering.py:64-102 driven from the device.  One round = schedule -> march -> hash-grid gather ->
    fused field MLP -> composite (+ compaction of the alive list); ray counts, samples per ray and the alive list
    stay on the GPU, every launch has a fixed grid, so ROUNDS_PER_SYNC rounds (the alive list ping-pongs) are captured
    as ONE CUDA graph and replayed; the host looks at the live-ray count once per replay instead of twice per round.  Same schedule (N_samples = max(min(N_rays // N_alive, 64), min_samples)), same kernels and the same
    per-ray arithmetic as the host loop, so the images are identical."""
    ROUNDS_PER_SYNC = 8                                  # host looks at the live-ray count this often
    ROUNDS_PER_GRAPH = 8                                 # rounds captured per graph (even; 2 measured the same)

    @classmethod
    def get(cls, model, n_rays, esf, T_threshold):
        p16, image = model._fused_state(model.center.device)
        key = (n_rays, float(esf), float(T_threshold), p16.data_ptr(), image.data_ptr(), model.density_bitfield.data_ptr())
        st = getattr(model, "_device_loop", None)
        if st is None or st.key != key:
            st = cls(model, n_rays, float(esf), float(T_threshold), key)
            model._device_loop = st
        return st

    def __init__(self, model, n, esf, T_threshold, key):
        self.key, self.model, self.n, self.esf, self.T = key, model, n, esf, T_threshold
        dev = model.center.device
        self.min_samples = 1 if esf == 0 else 4
        cap = self.cap = n * self.min_samples            # slots per round never exceed max(N_rays, N_alive*min_samples)
        f = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)
        self.rays_o, self.rays_d, self.hits = f(n, 3), f(n, 3), f(n, 2)
        self.alive = f(2, n, dt=torch.int64); self.arange = torch.arange(n, device=dev)
        self.n_eff = f(n, dt=torch.int32)
        self.xyzs, self.dirs, self.deltas, self.ts = f(cap, 3), f(cap, 3), f(cap), f(cap)
        self.enc = f(cap, model.k1, dt=torch.float16); self.sigmas, self.rgbs = f(cap), f(cap, 3)
        self.opacity, self.depth, self.rgb = f(n), f(n), f(n, 3)
        self.ctl = torch.zeros(8, dtype=torch.int32, device=dev)
        self.ctl_init = torch.tensor([0, 0, 0, 0, n, 0, 0, 0], dtype=torch.int32, device=dev)
        self.ctl_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self.graph = None

    def _reset(self):
        self.alive[0].copy_(self.arange)
        self.opacity.zero_(); self.depth.zero_(); self.rgb.zero_()
        self.ctl.copy_(self.ctl_init)

    def _round(self, cur):
        from .. import _lib as L
        m, P, call, n, cap = self.model, L.ptr, L.call, self.n, self.cap
        p16, image = m._fused_state(m.center.device)
        slots = self.ctl[2:]
        call("b2n_render_schedule", P(self.ctl), n, self.min_samples, MAX_SAMPLES)
        call("b2n_raymarching_test_dev", P(self.rays_o), P(self.rays_d), P(self.hits), P(self.alive[cur]),
             P(m.density_bitfield), m.cascades, float(m.scale), self.esf, m.grid_size, MAX_SAMPLES, n, P(self.ctl),
             P(self.xyzs), P(self.dirs), P(self.deltas), P(self.ts), P(self.n_eff))
        m._encode(self.xyzs, p16, out=self.enc, n_dev=slots)
        call("b2n_field_mlp_fw", P(self.enc), m.k1, P(self.dirs), P(image), cap, P(slots), P(self.sigmas), P(self.rgbs),
             None)
        call("b2n_composite_test_fw_dev", P(self.sigmas), P(self.rgbs), P(self.deltas), P(self.ts), P(self.alive[cur]),
             P(self.alive[1 - cur]), self.T, P(self.n_eff), n, P(self.ctl), P(self.opacity), P(self.depth), P(self.rgb))

    def _rounds(self):
        for k in range(self.ROUNDS_PER_GRAPH):
            self._round(k & 1)

    def run(self, rays_o, rays_d, hits):
        self.rays_o.copy_(rays_o); self.rays_d.copy_(rays_d); self.hits.copy_(hits)
        if self.graph is None:
            self._reset()
            side = torch.cuda.Stream(device=rays_o.device)   # eager warm-up, then capture
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._rounds()
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._rounds()
            self.hits.copy_(hits)                            # the warm-up rounds advanced the rays
        self._reset()
        while True:
            for _ in range(self.ROUNDS_PER_SYNC // self.ROUNDS_PER_GRAPH):
                self.graph.replay()
            self.ctl_host.copy_(self.ctl, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            c = self.ctl_host
            if int(c[4]) == 0 or int(c[3]) >= MAX_SAMPLES:
                break
        total = (int(c[7]) << 32) | (int(c[6]) & 0xffffffff)
        rgb = self.rgb + _background(self.esf, self.rgb.device) * (1 - self.opacity)[:, None]
        return {"opacity": self.opacity.clone(), "depth": self.depth.clone(), "rgb": rgb, "total_samples": total}


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
