"""render() with the signature, kwargs and result keys of ngp_pl/models/rendering.py:12-166.

kwargs: test_time (False), exp_step_factor (0.), T_threshold (1e-4), to_cpu (False).
results: opacity, depth, rgb, total_samples (+ depth_sq when training)."""
import torch

from .. import vren
from .custom_functions import RayAABBIntersector, RayMarcher, VolumeRenderer

MAX_SAMPLES = 1024
NEAR_DISTANCE = 0.05
WHOLE_RAYS = "auto"     # default of render(test_time=True, whole_rays=...): one persistent kernel per call (render_tc.cu)
WHOLE_RAYS_MAX_TABLE_BYTES = 96 << 20   # "auto": only while the fp16 hash table stays L2-resident (B200: 126 MB L2)


def render(model, rays_o, rays_d, **kwargs):
    """rays_o, rays_d (N_rays,3) -> dict.  AABB-clip, clamp the near hit to NEAR_DISTANCE, then train- or
    test-time rendering (rendering.py:26-39)."""
    rays_o = rays_o.contiguous(); rays_d = rays_d.contiguous()
    if not kwargs.get("test_time", False) and _train_graph_ok(model, rays_o, rays_d, kwargs):
        # training with this repo's NGP: the whole of render() is ONE autograd node replaying two CUDA graphs
        # (`graph=False` or RayMarcher.sync_free = False select the launch-by-launch form below)
        results = _render_train_graph(model, rays_o, rays_d, **kwargs)
        if kwargs.get("to_cpu", False):
            results = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in results.items()}
        return results
    if (kwargs.get("test_time", False) and kwargs.get("packed_out") is not None and rays_o.is_cuda
            and getattr(model, "fused", False) and kwargs.get("device_loop", True)
            and _WholeRays.supports(model, kwargs.get("whole_rays", WHOLE_RAYS))
            and not torch.cuda.is_current_stream_capturing()):
        # a rank's part of a sharded frame: box clip, near clamp and the whole-ray kernel replayed as ONE CUDA graph
        # that writes pixels and totals into the caller's block; nothing is read back here
        return _WholeRays.get(model, len(rays_o)).run_packed(
            rays_o, rays_d, kwargs.get("exp_step_factor", 0.), kwargs.get("T_threshold", 1e-4),
            kwargs["packed_out"], kwargs.get("tail_out"))
    if kwargs.get("test_time", False) and rays_o.is_cuda and not (rays_o.requires_grad or rays_d.requires_grad):
        # same kernel without the autograd.Function round trip (~0.1 ms of host time per frame)
        _, hits_t, _ = vren.ray_aabb_intersect(rays_o.float(), rays_d.float(), model.center, model.half_size, 1)
    else:
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
    fused = getattr(model, "fused", False)              # this repo's NGP (HashGrid L=16 or Frequency): fused field kernels
    if fused and kwargs.get("device_loop", True) and not torch.cuda.is_current_stream_capturing():
        if _WholeRays.supports(model, kwargs.get("whole_rays", WHOLE_RAYS)):
            res = _WholeRays.get(model, N_rays).run(rays_o, rays_d, hits, exp_step_factor, T_threshold)
            if res is not None:                           # None: a ray met the per-call sample budget -> round loop
                return res
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


class _WholeRays:
    """Test-time rendering with ONE persistent kernel per call (csrc/render_tc.cu): every ray runs the per-ray
    arithmetic of rendering.py:64-102 -- march, field, compositing -- start to finish; rays are independent at test
    time, the reference's rounds only batch them.  HashGrid field (k1 = 32).  Pixels agree with the round loop to fp32
    rounding (the transmittance is carried across a ray's rounds instead of being re-derived from the opacity at each
    round start), do not depend on which rays share a launch (a sharded frame equals the unsharded one bit for bit) and
    are reproducible.  The one thing a ray cannot know without the rounds is the per-call sample budget
    (`samples < MAX_SAMPLES`, rendering.py:66): the kernel counts rays that reach MAX_SAMPLES alive and `run` then
    returns None so that the caller renders that frame with the round loop (a box of scale 0.5 cannot get there:
    sqrt(3) / dt = 1024).  `total_samples` counts the samples marched under THIS kernel's row schedule (like the
    reference's it includes samples marched behind a ray's termination point, so it is schedule-dependent).
    `whole_rays=False` selects the round loop."""

    FIRST_HIT = True      # pre-pass: one thread per ray walks the empty space in front of it (see render_tc.cu)

    @staticmethod
    def supports(model, mode=True):
        """mode True: the HashGrid field (k1 = 32); "auto": additionally the fp16 table must fit the L2 -- the kernel
        gathers all 16 levels of a sample at once with 512 threads per SM, which is fine against the L2 but not against
        HBM (T = 2^22, 185 MB: 75 ms per 1920x1080 frame against 30 ms for the round loop with its level-major gather)."""
        if not mode or getattr(model, "encoding", None) != "HashGrid" or model.k1 != 32:
            return False
        return mode is True or model.xyz_encoder.enc.n_params * 2 <= WHOLE_RAYS_MAX_TABLE_BYTES

    @classmethod
    def get(cls, model, n_rays):
        """One state (control block, first-hit list, captured graph) per ray count; a few are kept so that a frame
        rendered in chunks with a shorter last chunk does not rebuild them every call."""
        pool = model.__dict__.setdefault("_whole_rays_pool", {})
        st = pool.get(n_rays)
        if st is None or st.dev != model.center.device:
            if len(pool) >= 4:
                pool.pop(next(iter(pool)))
            st = pool[n_rays] = cls(model, n_rays)
        model.__dict__["_whole_rays"] = st                   # the state of the last call (tests / bench look at it)
        return st

    def __init__(self, model, n):
        self.model, self.n, self.dev = model, n, model.center.device
        self.ctl = torch.zeros(8, dtype=torch.int32, device=self.dev)
        self.ctl_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self.list = torch.empty(max(n, 1), 2, dtype=torch.int32, device=self.dev)   # (ray, t at its first sample)

    def _launch(self, rays_o, rays_d, hits, esf, T_threshold, outs, stride, tail, ray_samples=None):
        from .. import _lib as L
        m, P = self.model, L.ptr
        p16, image = m._fused_state(self.dev)
        bg = 1.0 if esf == 0 else 0.0                            # rendering.py:108-111, blended by the kernel
        L.call("b2n_render_rays", P(rays_o), P(rays_d), P(hits), self.n, P(m.density_bitfield), m.cascades, float(m.scale),
               float(esf), m.grid_size, MAX_SAMPLES, m._layout, P(p16[m.xyz_encoder.mlp.n_params:]), P(image),
               float(T_threshold), bg, *outs, stride, P(tail), P(self.ctl), P(ray_samples),
               P(self.list) if self.FIRST_HIT else None)

    def run(self, rays_o, rays_d, hits, esf, T_threshold, ray_samples=None):
        """Dense outputs; returns the result dict, or None when a ray met the sample budget."""
        from .. import _lib as L
        P, n, dev = L.ptr, self.n, self.dev
        rays_o, rays_d, hits = rays_o.contiguous().float(), rays_d.contiguous().float(), hits.contiguous().float()
        opacity, depth, rgb = (torch.empty(n, device=dev), torch.empty(n, device=dev), torch.empty(n, 3, device=dev))
        self._launch(rays_o, rays_d, hits, esf, T_threshold, (P(opacity), P(depth), P(rgb)), 0, None, ray_samples)
        self.ctl_host.copy_(self.ctl, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        c = self.ctl_host
        self.rounds = int(c[4])
        if int(c[1]) > 0:
            return None
        total = (int(c[3]) << 32) | (int(c[2]) & 0xffffffff)
        return {"opacity": opacity, "depth": depth, "rgb": rgb, "total_samples": total}

    def run_packed(self, rays_o, rays_d, esf, T_threshold, packed_out, tail_out):
        """A rank's part of a sharded frame (dist_utils.render_sharded): pixels go to the rows of packed_out (n.., 5) =
        rgb | depth | opacity, the call's totals to tail_out (4 floats, see b2n_render_rays); nothing is read back --
        the caller looks at the gathered tails of all ranks once.  Box clip (rendering.py:27-28), near clamp (:29) and the
        kernel are captured as ONE CUDA graph per (buffers, settings) and replayed: three launches' worth of host time
        per frame instead of ~0.2 ms of Python in front of a kernel that takes 0.5 ms on an eighth of a frame."""
        from .. import _lib as L
        m, P, n, dev = self.model, L.ptr, self.n, self.dev
        assert packed_out.dtype == torch.float32 and packed_out.stride() == (5, 1) and packed_out.shape[0] >= n
        p16, image = m._fused_state(dev)
        key = (float(esf), float(T_threshold), packed_out.data_ptr(), tail_out.data_ptr() if tail_out is not None else 0,
               p16.data_ptr(), image.data_ptr(), m.density_bitfield.data_ptr(), m.center.data_ptr())
        if getattr(self, "_graph_key", None) != key:
            self._ro, self._rd = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
            self._hits = torch.empty(n, 1, 2, device=dev)
            self._hits_cnt, self._hits_idx = (torch.empty(n, dtype=torch.int32, device=dev),
                                              torch.empty(n, 1, dtype=torch.int64, device=dev))
            center, half = m.center.contiguous().float(), m.half_size.contiguous().float()
            base = packed_out.data_ptr()

            def body():
                L.call("b2n_ray_aabb_intersect", P(self._ro), P(self._rd), P(center), P(half), n, 1, 1, P(self._hits_cnt),
                       P(self._hits), P(self._hits_idx))
                L.call("b2n_clamp_near", P(self._hits), n, NEAR_DISTANCE)
                self._launch(self._ro, self._rd, self._hits, esf, T_threshold, (base + 16, base + 12, base), 5, tail_out)
            self._ro.copy_(rays_o); self._rd.copy_(rays_d)
            side = torch.cuda.Stream(device=dev)             # eager warm-up, then capture
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                body()
            torch.cuda.current_stream().wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                body()
            self._graph_key, self._keep = key, (center, half)
        self._ro.copy_(rays_o); self._rd.copy_(rays_d)
        self._graph.replay()
        return {"opacity": packed_out[:n, 4], "depth": packed_out[:n, 3], "rgb": packed_out[:n, 0:3],
                "total_samples": None, "tail": tail_out}


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


class _TrainGraph:
    """The training-time render of rendering.py:117-166 for the autograd API as TWO CUDA-graph replays over static
    buffers: forward = march (count + write) -> encode -> fused field kernel -> compositing, backward = compositing
    backward (+ alive list) -> fused field backward -> hash-grid scatter.  The per-step host work of the reference call
    path (train.py:144-170) drops from ~40 eager launches to two replays; counts stay on the device.  Buffers are sized
    from the sample count with headroom and grown when a step did not fit (that step's rays are truncated ray-wise by
    the marcher, reported once, like NGPTrainer)."""

    @classmethod
    def get(cls, model, n_rays, esf, T_threshold):
        pool = model.__dict__.setdefault("_train_graphs", {})
        key = (n_rays, float(esf), float(T_threshold))
        st = pool.get(key)
        if st is None:
            st = pool[key] = cls(model, n_rays, float(esf), float(T_threshold))
        return st

    def __init__(self, model, n, esf, T_threshold):
        self.model, self.n, self.esf, self.T = model, n, esf, T_threshold
        self.dev = model.center.device
        self.cap = max(128 * n, 4096)
        self.serial = 0
        self.host = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.event, self.pending, self.warned = torch.cuda.Event(), False, False
        self._alloc()

    def _alloc(self):
        n, cap, dev, m = self.n, self.cap, self.dev, self.model
        e = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)
        self.rays_o, self.rays_d, self.hits, self.noise = e(n, 3), e(n, 3), e(n, 1, 2), e(n)
        self.hits_cnt, self.hits_idx = e(n, dt=torch.int32), e(n, 1, dt=torch.int64)
        self.march_ws, self.rays_a = e(n, 64, dt=torch.int32), e(n, 3, dt=torch.int64)
        self.counter = torch.zeros(4, dtype=torch.int32, device=dev)
        self.xyzs, self.dirs, self.deltas, self.ts = e(cap, 3), e(cap, 3), e(cap), e(cap)
        self.enc, self.h = e(cap, m.k1, dt=torch.float16), e(cap, 16, dt=torch.float16)
        self.sigmas, self.rgbs = e(cap), e(cap, 3)
        # per-ray outputs side by side (one clone hands them to autograd): opacity | depth | depth_sq | blended rgb
        self.out = e(6 * n)
        self.opacity, self.depth, self.depth_sq = self.out[:n], self.out[n:2 * n], self.out[2 * n:3 * n]
        self.rgb_out = self.out[3 * n:].view(n, 3)
        self.rgb = e(n, 3)                                    # composited colour before the background blend
        self.gin = e(6 * n)                                   # incoming gradients, same packing
        self.gO, self.gD, self.gD2 = self.gin[:n], self.gin[n:2 * n], self.gin[2 * n:3 * n]
        self.gRGB = self.gin[3 * n:].view(n, 3)
        self.bg = 1.0 if self.esf == 0 else 0.0               # rendering.py:108-111: white for synthetic scenes, else black
        self.dL_dsigmas, self.dL_drgbs = e(cap), e(cap, 3)
        self.alive_idx, self.alive_cnt = e(cap, dt=torch.int32), torch.zeros(1, dtype=torch.int32, device=dev)
        self.din_enc = e(cap, 32, dt=torch.float16) if m.encoding == "HashGrid" else None
        self.g_xyz = torch.zeros(m.xyz_encoder.params.numel(), device=dev)
        self.g_rgb = torch.zeros(m.rgb_net.params.numel(), device=dev)
        self.found = torch.zeros(1, dtype=torch.int32, device=dev)
        self.graphs, self.graph_key = {}, None

    # ---------------------------------------------------------------- kernels (recorded into the graphs)
    def _fw(self, fixed_noise):
        from .. import _lib as L
        m, P, call, n, cap = self.model, L.ptr, L.call, self.n, self.cap
        p16, image = m._fused_state(self.dev)
        if not fixed_noise:
            self.noise.uniform_()
        call("b2n_ray_aabb_intersect", P(self.rays_o), P(self.rays_d), P(m.center), P(m.half_size), n, 1, 1,
             P(self.hits_cnt), P(self.hits), P(self.hits_idx))
        call("b2n_clamp_near", P(self.hits), n, NEAR_DISTANCE)
        march = (P(self.rays_o), P(self.rays_d), P(self.hits), P(m.density_bitfield), m.cascades, float(m.scale),
                 self.esf, P(self.noise), m.grid_size, MAX_SAMPLES, n)
        call("b2n_raymarching_train_count", *march, cap, P(self.rays_a), P(self.counter), P(self.march_ws))
        call("b2n_raymarching_train_write", *march, P(self.rays_a), P(self.xyzs), P(self.dirs), P(self.deltas),
             P(self.ts), P(self.march_ws))
        m._encode(self.xyzs, p16, out=self.enc, n_dev=self.counter)
        call("b2n_field_mlp_fw", P(self.enc), m.k1, P(self.dirs), P(image), cap, P(self.counter), P(self.sigmas),
             P(self.rgbs), P(self.h))
        call("b2n_composite_train_fw", P(self.sigmas), P(self.rgbs), P(self.deltas), P(self.ts), P(self.rays_a),
             self.T, n, P(self.opacity), P(self.depth), P(self.depth_sq), P(self.rgb))
        torch.add(self.rgb, (1 - self.opacity)[:, None], alpha=self.bg, out=self.rgb_out)      # rendering.py:163-164

    def _bw(self):
        from .. import _lib as L
        from .. import tinycudann as tcnn
        m, P, call, n, cap, S = self.model, L.ptr, L.call, self.n, self.cap, tcnn.LOSS_SCALE
        p16, image = m._fused_state(self.dev)
        self.gin.mul_(S)                                       # tcnn's loss scale rides on the per-ray gradients
        if self.bg:
            self.gO.sub_(self.gRGB.sum(-1), alpha=self.bg)       # the background blend's share of dL/dopacity
        call("b2n_composite_train_bw", P(self.gO), P(self.gD), P(self.gD2), P(self.gRGB), P(self.sigmas), P(self.rgbs),
             P(self.deltas), P(self.ts), P(self.rays_a), P(self.opacity), P(self.depth), P(self.depth_sq), P(self.rgb),
             self.T, n, P(self.dL_dsigmas), P(self.dL_drgbs), P(self.alive_idx), P(self.alive_cnt))
        self.g_xyz.zero_(); self.g_rgb.zero_(); self.found.zero_()
        call("b2n_field_mlp_bw", P(self.dL_dsigmas), P(self.dL_drgbs), P(self.enc), m.k1, P(self.dirs), P(image), cap,
             P(self.alive_cnt), P(self.rgbs), P(self.h), 1.0 / S, P(self.din_enc), P(self.g_xyz), P(self.g_rgb),
             P(self.alive_idx), 0, P(self.found))
        if self.din_enc is not None:
            call("b2n_hashgrid_bw", P(self.xyzs), P(self.din_enc), 32, m._layout, cap, P(self.alive_cnt), 1.0 / S,
                 P(self.g_xyz[m.xyz_encoder.mlp.n_params:]), P(self.alive_idx))
        # an fp16 overflow on the way surfaces as inf in the returned gradient (what a GradScaler keys on)
        self.g_rgb[:1] += torch.where(self.found > 0, float("inf"), 0.0)

    def _replay(self, name, fn):
        p16, image = self.model._fused_state(self.dev)        # refreshes the fp16 copies / weight image IN PLACE (eager)
        key = (p16.data_ptr(), image.data_ptr(), self.model.density_bitfield.data_ptr())
        if key != self.graph_key:
            self.graphs, self.graph_key = {}, key
        g = self.graphs.get(name)
        if g is None:
            side = torch.cuda.Stream(device=self.dev)         # eager warm-up, then capture
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)     # (this eager pass IS this call's execution)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            self.graphs[name] = g
        else:
            g.replay()

    def _size_from_first_batch(self, fixed_noise):
        """One exact count pass (and the only host synchronisation) before the first step: a fresh occupancy grid is
        fully occupied, so the first batches are the largest ones (hundreds of samples per ray)."""
        from .. import _lib as L
        m, P, n = self.model, L.ptr, self.n
        if not fixed_noise:
            self.noise.uniform_()
        L.call("b2n_ray_aabb_intersect", P(self.rays_o), P(self.rays_d), P(m.center), P(m.half_size), n, 1, 1,
               P(self.hits_cnt), P(self.hits), P(self.hits_idx))
        L.call("b2n_clamp_near", P(self.hits), n, NEAR_DISTANCE)
        L.call("b2n_raymarching_train_count", P(self.rays_o), P(self.rays_d), P(self.hits), P(m.density_bitfield),
               m.cascades, float(m.scale), self.esf, P(self.noise), m.grid_size, MAX_SAMPLES, n, -1, P(self.rays_a),
               P(self.counter), P(self.march_ws))
        total = int(self.counter[0].item())
        if total * 1.25 > self.cap:
            ro, rd, nz = self.rays_o, self.rays_d, self.noise
            self.cap = int(total * 1.25)
            self._alloc()
            self.rays_o.copy_(ro); self.rays_d.copy_(rd); self.noise.copy_(nz)

    # ---------------------------------------------------------------- the two halves
    def forward(self, rays_o, rays_d, fixed_noise):
        if self.pending and self.event.query():
            self.pending = False
            if int(self.host[2]):                             # the step before last did not fit: grow for the next ones
                if not self.warned:
                    import warnings
                    warnings.warn("render(): a training batch produced more samples than the buffers held; its rays were "
                                  "truncated for that step and the buffers have been enlarged")
                    self.warned = True
                self.cap = int(max(self.cap * 1.5, int(self.host[3]) * 1.25))
                self._alloc()
        self.rays_o.copy_(rays_o); self.rays_d.copy_(rays_d)
        if fixed_noise is not None:
            self.noise.copy_(fixed_noise)
        if self.serial == 0:
            self._size_from_first_batch(fixed_noise is not None)
        self._replay(("fw", fixed_noise is not None), lambda: self._fw(fixed_noise is not None))
        if not self.pending:
            self.host.copy_(self.counter, non_blocking=True); self.event.record(); self.pending = True
        self.serial += 1
        return self.serial

    def backward(self, g_out):
        self.gin.copy_(g_out)
        self._replay("bw", self._bw)
        return self.g_xyz.clone(), self.g_rgb.clone()


class _TrainRenderFn(torch.autograd.Function):
    """render() at training time (rendering.py:12-39 + 117-166: AABB clip, near clamp, march, field, compositing,
    background blend) as one autograd node over _TrainGraph.  Returns the packed per-ray outputs
    [opacity | depth | depth_sq | rgb] (6 * N_rays) and the sample count."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, xyz_params, rgb_params, model, esf, T_threshold):
        st = _TrainGraph.get(model, rays_o.shape[0], esf, T_threshold)
        ctx.st = st
        ctx.serial = st.forward(rays_o, rays_d, RayMarcher.noise)
        total = st.counter[0].clone()
        ctx.mark_non_differentiable(total)
        return st.out.clone(), total

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g_out, _):
        st = ctx.st
        if st.serial != ctx.serial:
            raise RuntimeError("render() ran again on this model before the previous result's backward pass; its "
                               "activations live in reused buffers -- call render(..., graph=False) for that pattern")
        g_xyz, g_rgb = st.backward(g_out)
        return None, None, g_xyz, g_rgb, None, None, None


def _train_graph_ok(model, rays_o, rays_d, kwargs):
    xe = getattr(model, "xyz_encoder", None)
    return (kwargs.get("graph", True) and RayMarcher.sync_free and getattr(model, "fused", False) and rays_o.is_cuda
            and torch.is_grad_enabled() and not rays_o.requires_grad and not rays_d.requires_grad and rays_o.shape[0] > 0
            and (xe.params.requires_grad or model.rgb_net.params.requires_grad)
            and not torch.cuda.is_current_stream_capturing())


def _render_train_graph(model, rays_o, rays_d, **kwargs):
    n = rays_o.shape[0]
    out, total_samples = _TrainRenderFn.apply(rays_o, rays_d, model.xyz_encoder.params, model.rgb_net.params, model,
                                              kwargs.get("exp_step_factor", 0.), kwargs.get("T_threshold", 1e-4))
    return {"total_samples": total_samples, "opacity": out[:n], "depth": out[n:2 * n], "depth_sq": out[2 * n:3 * n],
            "rgb": out[3 * n:].view(n, 3)}


def _render_rays_train(model, rays_o, rays_d, hits_t, **kwargs):
    """march -> field -> composite under autocast (rendering.py:117-166), one autograd node per reference Function
    (RayMarcher, the fused field node behind NGP.forward, VolumeRenderer); differentiates w.r.t. the rays as well.
    render() takes the graph-replayed single-node form (_TrainRenderFn) instead whenever it applies."""
    exp_step_factor = kwargs.get("exp_step_factor", 0.)
    T_threshold = kwargs.get("T_threshold", 1e-4)
    with torch.autocast("cuda", dtype=torch.float16):
        rays_a, xyzs, dirs, deltas, ts, total_samples = RayMarcher.apply(
            rays_o, rays_d, hits_t[:, 0], model.density_bitfield, model.cascades, model.scale, exp_step_factor,
            model.grid_size, MAX_SAMPLES)
        sigmas, rgbs = model(xyzs, dirs)
        opacity, depth, depth_sq, rgb = VolumeRenderer.apply(sigmas, rgbs.contiguous(), deltas, ts, rays_a, T_threshold)
        rgb = rgb + _background(exp_step_factor, rays_o.device) * (1 - opacity)[:, None]
    return {"total_samples": total_samples, "opacity": opacity, "depth": depth, "depth_sq": depth_sq, "rgb": rgb}


# the reference spells these with a leading double underscore (module-private names)
globals()["__render_rays_test"] = _render_rays_test
globals()["__render_rays_train"] = _render_rays_train
