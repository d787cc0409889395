"""NGPTrainer: one training step of ngp_pl/train.py:144-170 (render -> NeRFLoss -> backward -> FusedAdam,
density-grid update every 16 steps) run directly on the libb2n kernels, without autograd, on pre-allocated
buffers and with every count kept on the device, so the whole step can be replayed as one CUDA graph.

Data-parallel: one process per GPU, each with its own ray batch (like PL DDP at
train.py:262-263); gradients of the two flat parameters are summed with NCCL all-reduce and averaged, and the
occupancy grid is max-reduced after every update so that all ranks march the same bitfield.
"""
import math

import torch
import torch.distributed as dist

from . import _lib as L
from . import tinycudann as tc
from . import vren
from .models.rendering import MAX_SAMPLES, NEAR_DISTANCE

_f16, _f32 = torch.float16, torch.float32


class NGPTrainer:
    def __init__(self, model, n_rays=8192, lr=1e-2, eps=1e-15, betas=(0.9, 0.999), exp_step_factor=0.0,
                 T_threshold=1e-4, lambda_opa=1e-3, loss_scale=128.0, samples_per_ray=96, seed=0,
                 use_graph=True, process_group=None, grid_update_interval=16, warmup_steps=256):
        if model.encoding != "HashGrid":
            raise ValueError("NGPTrainer drives the HashGrid configuration; the Frequency variant trains through "
                             "render() + autograd")
        self.model, self.n_rays = model, n_rays
        self.dev = model.center.device
        if self.dev.type != "cuda":
            raise RuntimeError("NGPTrainer needs the model on a CUDA device (no CPU fallback)")
        model.init_grid_buffers()
        self.lr, self.eps, self.betas = lr, eps, betas
        self.esf, self.T_threshold, self.lambda_opa, self.loss_scale = exp_step_factor, T_threshold, lambda_opa, loss_scale
        self.bg = 1.0 if exp_step_factor == 0 else 0.0
        self.capacity = int(n_rays * samples_per_ray)
        torch.cuda.manual_seed(seed)
        self.fixed_noise = None                   # parity tests pin the per-ray jitter here
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.S, self.warmup_steps = grid_update_interval, warmup_steps
        self.step_count = 0
        self.use_graph, self.graph = use_graph, None
        self._from_indices = False

        xe, rn = model.xyz_encoder, model.rgb_net
        self.n_mlp = xe.mlp.n_params
        self.p_xyz, self.p_rgb = xe.params.data, rn.params.data
        z = lambda t: torch.zeros_like(t)
        self.g_xyz, self.g_rgb = z(self.p_xyz), z(self.p_rgb)
        self.m_xyz, self.v_xyz, self.m_rgb, self.v_rgb = z(self.p_xyz), z(self.p_xyz), z(self.p_rgb), z(self.p_rgb)
        self.h_xyz = tc.cast_half(self.p_xyz)
        self.h_rgb = tc.cast_half(self.p_rgb)
        self.hyper = torch.zeros(2, dtype=torch.int32, device=self.dev)        # {float lr; int32 step}
        self.w_image = torch.empty(10240, dtype=_f16, device=self.dev)
        self._pack_weights()
        # fold NGP.density's box normalisation (networks.py:96) into the hash-grid kernels
        self.layout = L.GridLayout.from_buffer_copy(xe.enc.layout)
        self.layout.x_offset = -float(model.scale)
        self.layout.x_scale = 1.0 / (2.0 * float(model.scale))
        self._alloc()

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        n, cap, dev = self.n_rays, self.capacity, self.dev
        e = lambda *s, dt=_f32: torch.empty(*s, dtype=dt, device=dev)
        self.rays_o, self.rays_d, self.target = e(n, 3), e(n, 3), e(n, 3)
        self.hits_cnt = e(n, dt=torch.int32); self.hits_t = e(n, 1, 2); self.hits_idx = e(n, 1, dt=torch.int64)
        self.noise = e(n)
        self.march_ws = e(n, 64, dt=torch.int32)
        self.rays_a = e(n, 3, dt=torch.int64); self.counter = torch.zeros(4, dtype=torch.int32, device=dev)
        self.xyzs, self.dirs, self.deltas, self.ts = e(cap, 3), e(cap, 3), e(cap), e(cap)
        self.enc = e(cap, 32, dt=_f16); self.hid_s = e(cap, 64, dt=_f16); self.h = e(cap, 16, dt=_f16)
        self.hid_r = e(2, cap, 64, dt=_f16)
        self.sigmas, self.rgbs = e(cap), e(cap, 3)
        self.opacity, self.depth, self.depth_sq, self.rgb = e(n), e(n), e(n), e(n, 3)
        self.rgb_out, self.loss = e(n, 3), torch.zeros(1, device=dev)
        self.dL_drgb, self.dL_dopacity = e(n, 3), e(n)
        self.zeros_n = torch.zeros(n, device=dev)
        self.dL_dsigmas, self.dL_drgbs = e(cap), e(cap, 3)
        self.din_enc = e(cap, 32, dt=_f16)
        self.graph = None

    # ------------------------------------------------------------------ the step body (all on the device)
    def _forward_backward(self):
        m, P, call = self.model, L.ptr, L.call
        n, cap = self.n_rays, self.capacity
        nd = self.counter                                             # counter[0] = #samples (device)
        if self._from_indices:
            self._gen_rays()
        call("b2n_ray_aabb_intersect", P(self.rays_o), P(self.rays_d), P(m.center), P(m.half_size), n, 1, 1,
             P(self.hits_cnt), P(self.hits_t), P(self.hits_idx))
        call("b2n_clamp_near", P(self.hits_t), n, NEAR_DISTANCE)
        if self.fixed_noise is None:
            self.noise.uniform_()                # default CUDA generator: graph-safe (philox offset is replayed)
        else:
            self.noise.copy_(self.fixed_noise)
        march = (P(self.rays_o), P(self.rays_d), P(self.hits_t), P(m.density_bitfield), m.cascades, float(m.scale),
                 float(self.esf), P(self.noise), m.grid_size, MAX_SAMPLES, n)
        call("b2n_raymarching_train_count", *march, cap, P(self.rays_a), P(self.counter), P(self.march_ws))
        call("b2n_raymarching_train_write", *march, P(self.rays_a), P(self.xyzs), P(self.dirs), P(self.deltas), P(self.ts),
             P(self.march_ws))
        # field forward: hash-grid gather, then the fused tcgen05 MLP chain (sigma + colour)
        call("b2n_hashgrid_fw", P(self.xyzs), P(self.h_xyz[self.n_mlp:]), self.layout, cap, P(nd), P(self.enc), 32)
        call("b2n_field_mlp_fw", P(self.enc), P(self.dirs), P(self.w_image), cap, P(nd), P(self.sigmas), P(self.rgbs),
             P(self.hid_s), P(self.h), P(self.hid_r))
        # compositing + loss
        call("b2n_composite_train_fw", P(self.sigmas), P(self.rgbs), P(self.deltas), P(self.ts), P(self.rays_a),
             self.T_threshold, n, P(self.opacity), P(self.depth), P(self.depth_sq), P(self.rgb))
        self.loss.zero_()
        call("b2n_nerf_loss_fwbw", P(self.rgb), P(self.opacity), P(self.target), n, self.bg, self.lambda_opa,
             self.loss_scale, P(self.rgb_out), P(self.loss), P(self.dL_drgb), P(self.dL_dopacity))
        call("b2n_composite_train_bw", P(self.dL_dopacity), P(self.zeros_n), P(self.zeros_n), P(self.dL_drgb),
             P(self.sigmas), P(self.rgbs), P(self.deltas), P(self.ts), P(self.rays_a), P(self.opacity), P(self.depth),
             P(self.depth_sq), P(self.rgb), self.T_threshold, n, P(self.dL_dsigmas), P(self.dL_drgbs))
        # field backward (gradients carry loss_scale; parameter gradients are unscaled inside Adam)
        call("b2n_field_mlp_bw", P(self.dL_dsigmas), P(self.dL_drgbs), P(self.enc), P(self.dirs), P(self.w_image), cap,
             P(nd), P(self.rgbs), P(self.hid_s), P(self.h), P(self.hid_r), 1.0, P(self.din_enc), P(self.g_xyz),
             P(self.g_rgb))
        call("b2n_hashgrid_bw", P(self.xyzs), P(self.din_enc), 32, self.layout, cap, P(nd), 1.0,
             P(self.g_xyz[self.n_mlp:]))

    def _allreduce(self):
        if self.world > 1:
            dist.all_reduce(self.g_xyz, group=self.pg)
            dist.all_reduce(self.g_rgb, group=self.pg)

    def _optimizer(self):
        P, call = L.ptr, L.call
        inv = 1.0 / (self.loss_scale * self.world)
        b1, b2 = self.betas
        for p, g, m, v, h in ((self.p_xyz, self.g_xyz, self.m_xyz, self.v_xyz, self.h_xyz),
                              (self.p_rgb, self.g_rgb, self.m_rgb, self.v_rgb, self.h_rgb)):
            call("b2n_adam_step", P(p), P(g), P(m), P(v), P(h), p.numel(), self.lr, b1, b2, self.eps, inv, 1,
                 P(self.hyper))
        self._pack_weights()

    def _pack_weights(self):
        """fp16 MLP weights -> UMMA canonical shared-memory image for the fused field kernels."""
        L.call("b2n_field_pack_weights", L.ptr(self.h_xyz), L.ptr(self.h_rgb), L.ptr(self.w_image))

    # ------------------------------------------------------------------ ray generation (train.py:150-157)
    def set_dataset(self, directions, poses):
        """directions (H*W,3) camera-frame ray directions, poses (N_img,3,4) c2w -- the two buffers the
        reference registers at train.py:99-100.  Batches then arrive as {img_idxs, pix_idxs, rgb}."""
        self.directions = directions.to(self.dev, _f32).contiguous()
        self.poses = poses.to(self.dev, _f32).contiguous()
        self.img_idxs = torch.zeros(self.n_rays, dtype=torch.int64, device=self.dev)
        self.pix_idxs = torch.zeros(self.n_rays, dtype=torch.int64, device=self.dev)
        self.graph = None

    def _gen_rays(self):
        # get_rays (datasets/ray_utils.py:152-175): rotate camera-frame directions, origin = camera centre
        c2w = self.poses[self.img_idxs]                               # (n,3,4)
        d = self.directions[self.pix_idxs]                            # (n,3)
        torch.bmm(d[:, None], c2w[..., :3].transpose(1, 2), out=self.rays_d.view(-1, 1, 3))
        self.rays_o.copy_(c2w[..., 3])

    def _body(self):
        self._forward_backward()
        self._allreduce()
        self._optimizer()

    # ------------------------------------------------------------------ public API
    def set_batch(self, rays_o, rays_d, target_rgb, non_blocking=True):
        """Copy one batch (host pinned or device tensors) into the step's input buffers."""
        if self._from_indices:
            self._from_indices, self.graph = False, None
        self.rays_o.copy_(rays_o, non_blocking=non_blocking)
        self.rays_d.copy_(rays_d, non_blocking=non_blocking)
        self.target.copy_(target_rgb, non_blocking=non_blocking)

    def set_batch_indices(self, img_idxs, pix_idxs, target_rgb, non_blocking=True):
        """The reference's training batch (datasets/base.py:28-33): image / pixel indices + ground-truth rgb."""
        if not self._from_indices:
            self._from_indices, self.graph = True, None
        self.img_idxs.copy_(img_idxs, non_blocking=non_blocking)
        self.pix_idxs.copy_(pix_idxs, non_blocking=non_blocking)
        self.target.copy_(target_rgb, non_blocking=non_blocking)

    def step_batch(self, batch):
        self.set_batch_indices(batch["img_idxs"], batch["pix_idxs"], batch["rgb"])
        return self.step()

    def _set_hyper(self):
        import struct
        packed = struct.unpack("i", struct.pack("f", float(self.lr)))[0]
        self.hyper.copy_(torch.tensor([packed, self.step_count], dtype=torch.int32), non_blocking=True)

    def step(self, rays_o=None, rays_d=None, target_rgb=None):
        """One optimisation step.  Returns the (device) loss tensor of this step; no host synchronisation."""
        if rays_o is not None:
            self.set_batch(rays_o, rays_d, target_rgb)
        if self.step_count % self.S == 0:
            self.update_density_grid(warmup=self.step_count < self.warmup_steps)
        self.step_count += 1
        self._set_hyper()
        if not self.use_graph:
            self._body()
        else:
            if self.graph is None:
                self._capture()
            if self.world == 1:
                self.graph.replay()
            else:                                  # NCCL all-reduce stays outside the captured graphs
                self.graph[0].replay()
                self._allreduce()
                self.graph[1].replay()
        return self.loss

    def _capture(self):
        # warm up on a side stream (allocator, cuBLAS-free path), then capture the whole step
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        state = [t.clone() for t in (self.p_xyz, self.p_rgb, self.m_xyz, self.v_xyz, self.m_rgb, self.v_rgb,
                                      self.h_xyz, self.h_rgb)]
        with torch.cuda.stream(s):
            self._body()
        torch.cuda.current_stream().wait_stream(s)
        for t, sv in zip((self.p_xyz, self.p_rgb, self.m_xyz, self.v_xyz, self.m_rgb, self.v_rgb, self.h_xyz,
                          self.h_rgb), state):
            t.copy_(sv)                                   # the warm-up must not count as a training step
        self.g_xyz.zero_(); self.g_rgb.zero_()
        self._pack_weights()
        if self.world == 1:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.graph = g
        else:
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                self._forward_backward()
            with torch.cuda.graph(g2):
                self._optimizer()
            self.graph = (g1, g2)

    @torch.no_grad()
    def update_density_grid(self, warmup=False):
        """train.py:145-148: threshold 0.01*MAX_SAMPLES/sqrt(3); all ranks end with the same grid."""
        m = self.model
        m.xyz_encoder.set_half_params(self.h_xyz)
        m.update_density_grid(0.01 * MAX_SAMPLES / 3 ** 0.5, warmup=warmup, erode=False)
        if self.world > 1:
            dist.all_reduce(m.density_grid, op=dist.ReduceOp.MAX, group=self.pg)
            ws = torch.empty(3, dtype=torch.float64, device=self.dev); stats = torch.empty(3, device=self.dev)
            L.call("b2n_grid_threshold", L.ptr(m.density_grid), m.density_grid.numel(),
                   float(0.01 * MAX_SAMPLES / 3 ** 0.5), L.ptr(ws), L.ptr(stats))
            vren.packbits(m.density_grid, 0.0, m.density_bitfield, threshold_dev=stats)

    def overflowed(self):
        """True if the last step hit the sample capacity (host sync; call sparingly)."""
        return bool(self.counter[2].item())

    def samples_last_step(self):
        return int(self.counter[0].item())

    def grow(self, factor=1.5):
        self.capacity = int(self.capacity * factor)
        self._alloc()

    def sync_model(self):
        """Make the nn.Module view consistent after direct parameter updates (bump versions, hand over fp16)."""
        self.model.xyz_encoder.set_half_params(self.h_xyz)
        self.model.rgb_net.set_half_params(self.h_rgb)
