"""NGPTrainer: one training step of ngp_pl/train.py:144-170 (render -> NeRFLoss -> backward -> FusedAdam,
density-grid update every 16 steps) run directly on the libb2n kernels, without autograd, on pre-allocated
buffers and with every count kept on the device, so the whole step is replayed as CUDA graphs.

Pipelining: ray generation + AABB + marching of batch k+1 depend only on the occupancy bitfield, not on the
weights, and the marcher is instruction-issue bound while the field kernels are latency / bandwidth bound.  When the
caller hands over the next batch (the reference's DataLoader always has it, train.py:126-131) its copies and its march
graph run on a side stream underneath the training graph of batch k [field forward .. Adam]; the main stream joins at
the end of the step.  Steps that are followed by a density-grid update are not pre-marched, so every batch is marched
against exactly the bitfield the sequential reference loop would use.

Data-parallel: one process per GPU, each with its own ray batch (like PL DDP at
train.py:262-263).  The flat parameter vector is sharded over the ranks (fp32 master + Adam moments); gradients are
summed, Adam runs on the owner's shard and the fp16 working copy is re-distributed -- by ONE kernel over NVLink peer
memory bracketed by flag barriers (comm="p2p", peer.py / csrc/peer.cu) or by NCCL reduce-scatter / all-gather
(comm="nccl").  The sampled cells of the occupancy update are shared out over the ranks and max-reduced, so all ranks
march the same bitfield.
"""
import os
import struct

import torch
import torch.distributed as dist

from . import _lib as L
from . import tinycudann as tc
from . import vren
from .models.rendering import MAX_SAMPLES, NEAR_DISTANCE

_f16, _f32 = torch.float16, torch.float32


class _SampleSet:
    """Inputs of one batch and the packed samples the marcher produces from them."""

    def __init__(self, n, cap, dev):
        e = lambda *s, dt=_f32: torch.empty(*s, dtype=dt, device=dev)
        self.rays_o, self.rays_d, self.target = e(n, 3), e(n, 3), e(n, 3)
        self.img_idxs = torch.zeros(n, dtype=torch.int64, device=dev)
        self.pix_idxs = torch.zeros(n, dtype=torch.int64, device=dev)
        self.hits_cnt = e(n, dt=torch.int32); self.hits_t = e(n, 1, 2); self.hits_idx = e(n, 1, dt=torch.int64)
        self.noise = e(n)
        self.prior = torch.zeros(n, device=dev)                            # depth-prior disparity per ray (<= 0: none)
        self.march_ws = e(n, 64, dt=torch.int32)
        self.rays_a = e(n, 3, dt=torch.int64)
        self.counter = torch.zeros(4, dtype=torch.int32, device=dev)      # [0] = #samples (device-side count)
        self.xyzs, self.dirs, self.deltas, self.ts = e(cap, 3), e(cap, 3), e(cap), e(cap)
        self.marched = False


class NGPTrainer:
    def __init__(self, model, n_rays=8192, lr=1e-2, eps=1e-15, betas=(0.9, 0.999), exp_step_factor=0.0,
                 T_threshold=1e-4, lambda_opa=1e-3, loss_scale=16384.0, samples_per_ray=96, seed=0,
                 use_graph=True, process_group=None, grid_update_interval=16, warmup_steps=256, data_parallel=True,
                 comm=None, comm_in_graph=False, grad_fp16=None, erode=False, scale_growth_interval=0,
                 serialize_mma=False, lambda_depth=0.0):
        """model: NGP with either encoding (HashGrid: networks.py:39-47; Frequency: networks.py:49-53).
        erode: the reference passes erode=True for colmap scenes (train.py:148).  loss_scale: start value of the
        device-side loss scaler.  The reference's fp16 backward pass runs at tcnn's internal 128 TIMES Lightning's
        GradScaler (precision=16, train.py:265: 65536 at start, halved on overflow); at 128 alone the per-sample
        gradients dL/denc of an 8192-ray batch (~1e-5) are fp16 subnormals and lose most of their mantissa, at 2^14
        they sit in the normal range with > 2^20 of headroom.  An overflow skips that optimiser step and halves the
        scale like GradScaler; scale_growth_interval > 0 doubles it again after that many clean steps (0: never)."""
        if not model.fused:
            raise ValueError("NGPTrainer drives the fused field kernels: 16 hash levels x 2 features, or Frequency-12 "
                             "(other --num_levels values train through render() + autograd)")
        self.model, self.n_rays = model, n_rays
        self.hashed = model.encoding == "HashGrid"
        self.k1 = model.k1
        self.erode, self.serialize_mma = bool(erode), int(bool(serialize_mma))
        self.lambda_depth = float(lambda_depth)
        self.dev = model.center.device
        if self.dev.type != "cuda":
            raise RuntimeError("NGPTrainer needs the model on a CUDA device (no CPU fallback)")
        model.init_grid_buffers()
        self.lr, self.eps, self.betas = lr, eps, betas
        self.esf, self.T_threshold, self.lambda_opa, self.loss_scale = exp_step_factor, T_threshold, lambda_opa, loss_scale
        self.bg = 1.0 if exp_step_factor == 0 else 0.0
        self.capacity = int(n_rays * samples_per_ray)
        self.scale_growth_interval = int(scale_growth_interval)
        self.fixed_noise = None                   # parity tests pin the per-ray jitter here
        self.fixed_grid_noise = None              # ... and the occupancy update's per-cascade (G^3,3) cell jitter here
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (data_parallel and dist.is_available() and dist.is_initialized()) else 1
        self.S, self.warmup_steps = grid_update_interval, warmup_steps
        self.step_count = 0
        self.use_graph = use_graph
        self._from_indices = False
        self.directions = self.poses = None
        self.side = torch.cuda.Stream(device=self.dev)
        self._fork = torch.cuda.Event()
        # Kernel nodes of a captured graph inherit the priority of the stream they were captured on: the training chain
        # is captured on a high-priority stream, the pre-march of the next batch (which runs underneath it on the side
        # stream) on a default-priority one, so that the block scheduler serves the chain's CTAs first
        self._prio = os.environ.get("B2N_PRIO", "1") == "1"
        self._cap_streams = (torch.cuda.Stream(device=self.dev, priority=-1), torch.cuda.Stream(device=self.dev, priority=0))
        # thread-per-ray count pass for the prefetched batch: fewer issue slots, but a long per-ray latency -- it only
        # pays once a batch has enough rays to fill the machine with threads (measured at 8192 rays: the tail of the
        # longest rays outlasts the training step, 0.47 vs 0.415 ms/step)
        self.serial_prefetch = os.environ.get("B2N_SERIAL_PREFETCH", "1" if n_rays >= 65536 else "0") == "1"

        xe, rn = model.xyz_encoder, model.rgb_net
        self.n_mlp = xe.mlp.n_params
        z = lambda t: torch.zeros_like(t)
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        # every rank draws its own jitter / occupancy cells: the sharded occupancy update relies on the ranks sampling
        # different cells (M / world each) before the max-reduce
        torch.cuda.manual_seed(seed + self.rank)
        n_xyz, n_rgb = xe.params.numel(), rn.params.numel()
        # All trainable parameters live in ONE flat vector [xyz_encoder (MLP + hash table) | rgb_net | pad]: one Adam
        # launch, and for world > 1 one exchange.  Sharded optimiser (world > 1): the vector is padded to world * shard;
        # the gradients are summed shard-wise, every rank runs Adam on ITS shard of the fp32 master / moments only, and
        # the fp16 working copy goes back to all ranks.  Less traffic than an all-reduce (fp32 one way, fp16 back) and
        # Adam's HBM sweep shrinks by the world size.  The model's Parameters are views of the master buffer.
        assert n_xyz % 8 == 0
        n_all = n_xyz + n_rgb
        self.n_pad = -(-n_all // (8 * self.world)) * (8 * self.world)
        self.shard = self.n_pad // self.world
        p_pad = torch.zeros(self.n_pad, dtype=_f32, device=self.dev)
        p_pad[:n_xyz].copy_(xe.params.data); p_pad[n_xyz:n_all].copy_(rn.params.data)
        xe.params.data, rn.params.data = p_pad[:n_xyz], p_pad[n_xyz:n_all]
        self.p_pad = p_pad
        self.p_xyz, self.p_rgb = xe.params.data, rn.params.data
        lo = self.rank * self.shard
        self.comm = (comm or "p2p") if self.world > 1 else "none"
        if self.comm not in ("none", "p2p", "nccl"):
            raise ValueError(f"unknown comm {comm!r}")
        self.peer, self.grad_fp16 = None, False
        if self.comm == "p2p":
            # gradient vector and fp16 working copy live in NVLink peer memory: the fused reduce + Adam + broadcast
            # kernel of every rank addresses all of them (peer.py / csrc/peer.cu)
            from .peer import PeerBlock
            # wire format of the gradients: fp16 copies halve the NVLink bytes of the exchange at the price of one
            # local pack pass (which also clears the fp32 vector); measured to pay from 4 ranks up
            self.grad_fp16 = os.environ.get("B2N_GRAD_FP16", "1" if self.world >= 4 else "0") == "1" \
                if grad_fp16 is None else bool(grad_fp16)
            regions = {"g": (_f32, self.n_pad), "h": (_f16, self.n_pad), "hy": (torch.int32, 8)}
            if self.grad_fp16:
                regions["g16"] = (_f16, self.n_pad)
            try:
                self.peer = PeerBlock(regions, self.dev, process_group)      # raises on ALL ranks or on none
            except RuntimeError as e:
                if comm == "p2p":                            # asked for explicitly
                    raise
                import warnings
                warnings.warn(f"NVLink peer memory is not available ({e}); exchanging through NCCL instead")
                self.comm, self.grad_fp16 = "nccl", False
        if self.comm == "p2p":
            self.g_all, self.h_all = self.peer.tensor("g"), self.peer.tensor("h")
            self.h_all.copy_(tc.cast_half(p_pad))
            torch.cuda.synchronize(self.dev)
            dist.barrier(group=process_group)
        else:
            # gradient vector and fp16 copy side by side: one L2 persistence window covers both
            self._gh = torch.zeros(6 * self.n_pad, dtype=torch.uint8, device=self.dev)
            self.g_all = self._gh[:4 * self.n_pad].view(_f32)
            self.h_all = self._gh[4 * self.n_pad:].view(_f16)
            self.h_all.copy_(tc.cast_half(p_pad))
        # (an L2 persistence window over gradient + fp16 copy was measured NEGATIVE on B200 in round 1 -- 0.460 vs
        # 0.418 ms/step: the 69 MB carve-out halves the L2 left for write-combining the activation stream -- and is gone)
        self.g_xyz, self.g_rgb = self.g_all[:n_xyz], self.g_all[n_xyz:n_all]
        self.p_shard = p_pad[lo:lo + self.shard]
        self.g_shard = torch.zeros(self.shard, dtype=_f32, device=self.dev) if self.comm == "nccl" else \
            self.g_all[lo:lo + self.shard]
        self.m, self.v = z(self.p_shard), z(self.p_shard)
        self.h_xyz, self.h_rgb = self.h_all[:n_xyz], self.h_all[n_xyz:n_all]
        self.h_shard = self.h_all[lo:lo + self.shard]
        # comm == "nccl" only: True captures the collectives into the training graph.  Measured: 22 us/step faster at 2
        # GPUs, but communicator teardown after captured collectives hung on this stack, so the default keeps them
        # eager between two graphs.
        self.comm_in_graph = comm_in_graph
        # b2n_hyper (include/b2n.h): {lr, step, found_inf, skipped, loss_scale, good_steps, growth_interval, -}; in
        # peer memory when ranks exchange through it (an overflow on ANY rank skips the step on all of them)
        self.hyper = self.peer.tensor("hy") if self.comm == "p2p" else torch.zeros(8, dtype=torch.int32, device=self.dev)
        self.hyper.copy_(torch.tensor([0, 0, 0, 0, struct.unpack("i", struct.pack("f", float(loss_scale)))[0], 0,
                                       self.scale_growth_interval, 0], dtype=torch.int32))
        self.w_image = torch.empty(64 * self.k1 + 8192, dtype=_f16, device=self.dev)
        self._pack_weights()
        self.sync_model_half()                    # the model's fused paths read the fp16 copies / image kept current here
        # fold NGP.density's box normalisation (networks.py:96) into the hash-grid kernels
        self.layout = model._layout
        self._peer_err = None
        if self.comm == "p2p":
            self._peer_err = (torch.zeros(2, dtype=torch.int32).pin_memory(), torch.cuda.Event())
            self._peer_err_pending = False
            torch.cuda.synchronize(self.dev)
            dist.barrier(group=process_group)     # every rank's control block is initialised before anyone reads it
        self._alloc()

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        n, cap, dev = self.n_rays, self.capacity, self.dev
        e = lambda *s, dt=_f32: torch.empty(*s, dtype=dt, device=dev)
        self.sets = [_SampleSet(n, cap, dev), _SampleSet(n, cap, dev)]
        self.cur = 0
        # saved activations: enc and h only (the hidden layers are recomputed by the backward kernel)
        self.enc = e(cap, self.k1, dt=_f16); self.h = e(cap, 16, dt=_f16)
        self.sigmas, self.rgbs = e(cap), e(cap, 3)
        self.opacity, self.depth = e(n), e(n)
        self.rgb_out = e(n, 3)
        if self.lambda_depth > 0:
            self.depth_sq, self.rgb_lin = e(n), e(n, 3)
            self.dL_drgb, self.dL_dopacity, self.dL_ddepth = e(n, 3), e(n), e(n)
            self.zeros_n, self.depth_stats = torch.zeros(n, device=dev), torch.zeros(8, device=dev)
        # [alive count, loss] side by side, so that the compositing launch clears both with one memset
        self._scalars = torch.zeros(4, dtype=torch.int32, device=dev)
        self.alive_cnt, self.loss = self._scalars[0:1], self._scalars[1:2].view(_f32)
        self.dL_dsigmas, self.dL_drgbs = e(cap), e(cap, 3)
        self.din_enc = e(cap, 32, dt=_f16) if self.hashed else None
        self.alive_idx = e(cap, dt=torch.int32)
        self.last_counter = self.sets[0].counter
        self.graphs = {}

    # ------------------------------------------------------------------ the step body (all on the device)
    def _march(self, s, serial=False):
        """ray generation (train.py:150-157) -> AABB (+ near clamp) -> occupancy marcher, into sample set s.
        serial: count pass with one thread per ray (prefetch on the side stream: latency is hidden there, and it
        takes ~10x fewer issue slots away from the training kernels than the warp-per-ray form)."""
        m, P, call = self.model, L.ptr, L.call
        n, cap = self.n_rays, self.capacity
        if self._from_indices:
            # get_rays (datasets/ray_utils.py:152-175): rotate camera-frame directions, origin = camera centre
            call("b2n_rays_from_indices", P(self.directions), P(self.poses), P(s.img_idxs), P(s.pix_idxs), n,
                 P(s.rays_o), P(s.rays_d))
        call("b2n_ray_aabb_intersect", P(s.rays_o), P(s.rays_d), P(m.center), P(m.half_size), n, 1, 1,
             P(s.hits_cnt), P(s.hits_t), P(s.hits_idx))
        call("b2n_clamp_near", P(s.hits_t), n, NEAR_DISTANCE)
        if self.fixed_noise is None:
            s.noise.uniform_()                   # default CUDA generator: graph-safe (philox offset is replayed)
        else:
            s.noise.copy_(self.fixed_noise)
        march = (P(s.rays_o), P(s.rays_d), P(s.hits_t), P(m.density_bitfield), m.cascades, float(m.scale),
                 float(self.esf), P(s.noise), m.grid_size, MAX_SAMPLES, n)
        call("b2n_raymarching_train_count_serial" if serial else "b2n_raymarching_train_count", *march, cap,
             P(s.rays_a), P(s.counter), P(s.march_ws))
        call("b2n_raymarching_train_write", *march, P(s.rays_a), P(s.xyzs), P(s.dirs), P(s.deltas), P(s.ts), P(s.march_ws))

    def _forward_backward(self, s):
        P, call = L.ptr, L.call
        n, cap, nd = self.n_rays, self.capacity, s.counter
        # field forward: encode (hash-grid gather / frequency), then the fused tcgen05 MLP chain (sigma + colour)
        if self.hashed:
            call("b2n_hashgrid_fw", P(s.xyzs), P(self.h_xyz[self.n_mlp:]), self.layout, cap, P(nd), P(self.enc), 32)
        else:
            call("b2n_frequency_fw", P(s.xyzs), 12, cap, P(nd), P(self.enc), self.k1, -float(self.model.scale),
                 2.0 * float(self.model.scale))
        call("b2n_field_mlp_fw", P(self.enc), self.k1, P(s.dirs), P(self.w_image), cap, P(nd), P(self.sigmas),
             P(self.rgbs), P(self.h))
        # compositing + loss (the loss scale is read from the device-side scaler)
        if self.lambda_depth > 0:
            # depth prior: its medians need every ray's depth before any gradient exists, so compositing forward, the
            # two loss kernels and compositing backward (with dL/ddepth) are separate launches here
            ls_dev = P(self.hyper[4:])
            self._scalars.zero_()
            call("b2n_composite_train_fw", P(self.sigmas), P(self.rgbs), P(s.deltas), P(s.ts), P(s.rays_a),
                 self.T_threshold, n, P(self.opacity), P(self.depth), P(self.depth_sq), P(self.rgb_lin))
            call("b2n_nerf_loss_fwbw", P(self.rgb_lin), P(self.opacity), P(s.target), n, self.bg, self.lambda_opa,
                 self.loss_scale, P(self.rgb_out), P(self.loss), P(self.dL_drgb), P(self.dL_dopacity), ls_dev)
            call("b2n_ssi_depth_loss_fwbw", P(self.depth), P(s.prior), n, self.lambda_depth, self.loss_scale, ls_dev,
                 P(self.loss), P(self.dL_ddepth), P(self.depth_stats))
            call("b2n_composite_train_bw", P(self.dL_dopacity), P(self.dL_ddepth), P(self.zeros_n), P(self.dL_drgb),
                 P(self.sigmas), P(self.rgbs), P(s.deltas), P(s.ts), P(s.rays_a), P(self.opacity), P(self.depth),
                 P(self.depth_sq), P(self.rgb_lin), self.T_threshold, n, P(self.dL_dsigmas), P(self.dL_drgbs),
                 P(self.alive_idx), P(self.alive_cnt))
        else:
            call("b2n_composite_loss_fwbw", P(self.sigmas), P(self.rgbs), P(s.deltas), P(s.ts), P(s.rays_a), P(s.target),
                 self.T_threshold, n, self.bg, self.lambda_opa, self.loss_scale, P(self.opacity), P(self.depth),
                 P(self.rgb_out), P(self.loss), P(self.dL_dsigmas), P(self.dL_drgbs), P(self.alive_idx),
                 P(self.alive_cnt), P(self.hyper[4:]))
        # field backward (gradients carry the loss scale; parameter gradients are unscaled inside Adam)
        # only the samples composited before each ray's early stop carry gradient: the two heavy backward kernels
        # run over that compacted list (alive_cnt is a device-side count)
        call("b2n_field_mlp_bw", P(self.dL_dsigmas), P(self.dL_drgbs), P(self.enc), self.k1, P(s.dirs), P(self.w_image),
             cap, P(self.alive_cnt), P(self.rgbs), P(self.h), 1.0, P(self.din_enc), P(self.g_xyz), P(self.g_rgb),
             P(self.alive_idx), self.serialize_mma, P(self.hyper[2:]))
        if self.hashed:
            call("b2n_hashgrid_bw", P(s.xyzs), P(self.din_enc), 32, self.layout, cap, P(self.alive_cnt), 1.0,
                 P(self.g_xyz[self.n_mlp:]), P(self.alive_idx))

    def _reduce_grads(self):
        """comm == "nccl": sum the gradients over ranks; every rank keeps its shard (reduce-scatter).  The overflow
        flags are max-reduced so that every rank skips (or takes) the step together."""
        dist.reduce_scatter_tensor(self.g_shard, self.g_all, op=dist.ReduceOp.SUM, group=self.pg)
        dist.all_reduce(self.hyper[2:3], op=dist.ReduceOp.MAX, group=self.pg)

    def _optimizer(self):
        """Adam on this rank's shard from the (already reduced, or local) gradient shard."""
        P, call = L.ptr, L.call
        inv = 1.0 / self.world                               # the loss scale is divided out from the device block
        b1, b2 = self.betas
        call("b2n_adam_step", P(self.p_shard), P(self.g_shard), P(self.m), P(self.v), P(self.h_shard), self.shard,
             self.lr, b1, b2, self.eps, inv, 1, P(self.hyper), 1)
        if self.world == 1:
            self._pack_weights(scaler=True)                  # + loss-scaler bookkeeping, found_inf consumed
        else:
            call("b2n_scaler_update", P(self.hyper), None, 1)
            self.hyper[2:3].zero_()                          # found_inf consumed
            self.g_all.zero_()                               # Adam only zeroed this rank's (reduced) shard

    def _gather_params(self):
        """comm == "nccl": every rank receives the other shards of the fp16 working copy, then re-packs the MLP image."""
        dist.all_gather_into_tensor(self.h_all, self.h_shard, group=self.pg)
        self._pack_weights()

    def _optimizer_peer(self):
        """comm == "p2p": barrier, then ONE kernel that sums this rank's gradient slice over all ranks through NVLink
        peer loads, runs Adam on the local master shard and stores the fp16 result into every rank's working copy;
        barrier; clear the local gradient vector; re-pack the MLP image."""
        P, call, pb = L.ptr, L.call, self.peer
        inv = 1.0 / self.world
        b1, b2 = self.betas
        lo16, hi16 = self.n_mlp, self.p_xyz.numel()           # the hash table; the MLP weights stay fp32 on the wire
        if self.grad_fp16:
            call("b2n_grad_pack_half", P(self.g_all), P(pb.tensor("g16")), lo16, hi16)
        pb.barrier()
        call("b2n_adam_step_peer", P(self.p_shard), P(self.m), P(self.v), pb.table("g"),
             pb.table("g16") if self.grad_fp16 else None, lo16, hi16 if self.grad_fp16 else lo16, pb.table("h"),
             self.world, self.rank * self.shard, self.shard, self.lr, b1, b2, self.eps, inv, 1, P(self.hyper),
             pb.table("hy"), P(pb.state))
        call("b2n_scaler_update", P(self.hyper), pb.table("hy"), self.world)     # same OR on every rank
        pb.barrier()
        self.hyper[2:3].zero_()                              # every rank has read every flag
        if self.grad_fp16:
            self.g_all[:lo16].zero_(); self.g_all[hi16:].zero_()
        else:
            self.g_all.zero_()
        self._pack_weights()

    def _pack_weights(self, scaler=False):
        """fp16 MLP weights -> UMMA canonical shared-memory image for the fused field kernels."""
        L.call("b2n_field_pack_weights", L.ptr(self.h_xyz), L.ptr(self.h_rgb), L.ptr(self.w_image), self.k1,
               L.ptr(self.hyper) if scaler else None)

    def _train(self, p):
        """Forward, backward and the optimiser on sample set p (world > 1: with the gradient reduce-scatter and the
        parameter all-gather around the optimiser, unless the collectives are kept out of the graph)."""
        self._forward_backward(self.sets[p])
        if self.world == 1:
            self._optimizer()
        elif self.comm == "p2p":
            self._optimizer_peer()
        elif self.comm_in_graph:
            self._reduce_grads()
            self._optimizer()
            self._gather_params()

    # ------------------------------------------------------------------ graphs
    def _run(self, key, fn):
        """Replay the CUDA graph of `fn` (captured on first use), or run it eagerly when graphs are off."""
        if not self.use_graph:
            fn()
            return
        g = self.graphs.get(key)
        if g is None:
            g = self._capture(fn, touches_params=key[0] in ("train", "opt"), touches_grid=key[0].startswith("grid"),
                              low_priority=key[0] == "march")
            self.graphs[key] = g
        g.replay()

    def _capture(self, fn, touches_params, touches_grid=False, low_priority=False):
        # one eager warm-up on a side stream, with the optimiser state restored afterwards so that it does not
        # count as a training step; then capture
        state_t = (self.p_pad, self.m, self.v, self.h_all, self.g_all, self.hyper) + \
            ((self.g_shard,) if self.comm == "nccl" else ())
        if touches_grid:
            if getattr(self.model, "_grid_tmp", None) is None:        # the persistent scratch grid of the update
                self.model._grid_tmp = torch.zeros_like(self.model.density_grid)
            state_t, touches_params = (self.model.density_grid, self.model.density_bitfield, self.model._grid_tmp), True
        saved = [t.clone() for t in state_t] if touches_params else None
        warm = torch.cuda.Stream(device=self.dev)
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):
            fn()
        torch.cuda.current_stream().wait_stream(warm)
        if saved is not None:
            for t, sv in zip(state_t, saved):
                t.copy_(sv)
            if not touches_grid:
                self._pack_weights()
        g = torch.cuda.CUDAGraph()
        if self._prio:
            with torch.cuda.graph(g, stream=self._cap_streams[1 if low_priority else 0]):
                fn()
        else:
            with torch.cuda.graph(g):
                fn()
        return g

    # ------------------------------------------------------------------ public API
    def set_dataset(self, directions, poses):
        """directions (H*W,3) camera-frame ray directions, poses (N_img,3,4) c2w -- the two buffers the
        reference registers at train.py:99-100.  Batches then arrive as {img_idxs, pix_idxs, rgb}."""
        self.directions = directions.to(self.dev, _f32).contiguous()
        self.poses = poses.to(self.dev, _f32).contiguous()
        self.graphs = {}

    def _load(self, s, batch, non_blocking=True):
        """Copy one batch (host pinned or device tensors) into sample set s.  batch = (rays_o, rays_d, rgb) or the
        reference's {img_idxs, pix_idxs, rgb} dict (datasets/base.py:28-33)."""
        from_idx = isinstance(batch, dict)
        cur = torch.cuda.current_stream()
        for t in (batch.values() if from_idx else batch):
            if t.is_cuda:
                t.record_stream(cur)                         # the copy may run on the side stream
        if from_idx != self._from_indices:
            self._from_indices, self.graphs = from_idx, {}
        if from_idx:
            if self.directions is None:
                raise RuntimeError("set_dataset(directions, poses) must be called before index batches are used")
            s.img_idxs.copy_(batch["img_idxs"], non_blocking=non_blocking)
            s.pix_idxs.copy_(batch["pix_idxs"], non_blocking=non_blocking)
            s.target.copy_(batch["rgb"], non_blocking=non_blocking)
            if self.lambda_depth > 0:
                s.prior.copy_(batch["disp"], non_blocking=non_blocking)
        else:
            s.rays_o.copy_(batch[0], non_blocking=non_blocking)
            s.rays_d.copy_(batch[1], non_blocking=non_blocking)
            s.target.copy_(batch[2], non_blocking=non_blocking)
            if self.lambda_depth > 0:
                s.prior.copy_(batch[3], non_blocking=non_blocking)
        s.marched = False

    def set_batch(self, rays_o, rays_d, target_rgb, prior_disp=None):
        self._load(self.sets[self.cur], (rays_o, rays_d, target_rgb) + ((prior_disp,) if prior_disp is not None else ()))

    def _set_hyper(self):
        # a fresh pageable tensor per step: the driver stages the copy at once, so the host may run many graph replays
        # ahead without ever overwriting a value the GPU has not consumed yet
        packed = struct.unpack("i", struct.pack("f", float(self.lr)))[0]
        self.hyper[:2].copy_(torch.tensor([packed, self.step_count], dtype=torch.int32), non_blocking=True)

    def step(self, rays_o=None, rays_d=None, target_rgb=None, next_batch=None, prior_disp=None):
        """One optimisation step on the given (or previously set / pre-marched) batch.  `next_batch`, if given, is
        marched concurrently for the following step.  Returns the device loss tensor; no host synchronisation."""
        p = self.cur
        s = self.sets[p]
        if rays_o is not None:
            self._load(s, (rays_o, rays_d, target_rgb) + ((prior_disp,) if prior_disp is not None else ()))
        if self.step_count % self.S == 0:
            self.update_density_grid(warmup=self.step_count < self.warmup_steps)
            s.marched = False                                # (never pre-marched across an update anyway)
        self.step_count += 1
        self._set_hyper()
        if not s.marched:
            self._run(("march", p, False), lambda: self._march(s))
        prefetch = next_batch is not None and (self.step_count % self.S != 0)
        main = torch.cuda.current_stream()
        if prefetch:
            self._fork.record(main)
        self._run(("train", p), lambda: self._train(p))      # issued first: the GPU starts on it while the host goes on
        if prefetch:
            # the next batch is loaded and marched on the side stream while this one trains: its copies and the march
            # graph are ordered after everything issued before this step's training graph (the previous step's use of
            # that sample set) and the main stream joins at the end of the step
            self.side.wait_event(self._fork)
            with torch.cuda.stream(self.side):
                self._load(self.sets[1 - p], next_batch)
                self._run(("march", 1 - p, self.serial_prefetch),
                          lambda: self._march(self.sets[1 - p], serial=self.serial_prefetch))
        if self.comm == "nccl" and not self.comm_in_graph:   # collectives between two graphs
            self._reduce_grads()
            self._run(("opt",), self._optimizer)
            self._gather_params()
        if prefetch:
            main.wait_stream(self.side)
        s.marched = False
        self.last_counter = s.counter
        if next_batch is not None:
            if not prefetch:
                self._load(self.sets[1 - p], next_batch)
            else:
                self.sets[1 - p].marched = True
            self.cur = 1 - p
        return self.loss

    def step_batch(self, batch, next_batch=None):
        """The reference's training_step input: {img_idxs, pix_idxs, rgb}."""
        s = self.sets[self.cur]
        if batch is not None:
            self._load(s, batch)
        return self.step(next_batch=next_batch)

    @torch.no_grad()
    def update_density_grid(self, warmup=False):
        """train.py:145-148: threshold 0.01*MAX_SAMPLES/sqrt(3).  The update has no host synchronisation, so it is
        replayed as CUDA graphs as well (it is ~25 small launches).  world > 1: the sampled cells are shared out over
        the ranks (each evaluates 1/world of them), the sampled densities are max-reduced and every rank merges the
        same values, so all ranks keep the same grid and march the same bitfield."""
        thr = 0.01 * MAX_SAMPLES / 3 ** 0.5
        m, w = self.model, bool(warmup)
        self.sync_model_half()
        if self.world == 1:
            self._run(("grid", w), lambda: m.update_density_grid(thr, warmup=w, erode=self.erode,
                                                                 noise=self.fixed_grid_noise if w else None))
            return
        self._check_peer()
        self._run(("grid_eval", w), lambda: m._grid_eval(w, (self.rank, self.world)))
        dist.all_reduce(m._grid_tmp, op=dist.ReduceOp.MAX, group=self.pg)
        self._run(("grid_commit",), lambda: m._grid_commit(thr, 0.95, self.erode))

    def _check_peer(self):
        """comm == "p2p": the barrier's sticky error word is copied to pinned host memory at every grid update and the
        PREVIOUS copy is inspected (no host synchronisation on the training path): a peer that stopped arriving raises
        here on every rank within two update intervals (the exchange kernel is already a no-op from the first timeout)."""
        if self._peer_err is None:
            return
        host, ev = self._peer_err
        if self._peer_err_pending and ev.query():
            err = int(host[1])
            if err:
                raise RuntimeError(f"peer barrier timed out waiting for rank {err - 1}; parameters were left untouched")
        host.copy_(self.peer.state, non_blocking=True)
        ev.record()
        self._peer_err_pending = True

    def close(self):
        """Collective: release the NVLink peer mappings (comm == "p2p").  Call on every rank when training ends."""
        self.graphs = {}
        if self.peer is not None:
            # the fp16 working copies the model's fused paths read live in the peer block: hand the model private copies
            # before the block is freed (render() / density() keep working after training ends)
            self.h_xyz, self.h_rgb = self.h_xyz.clone(), self.h_rgb.clone()
            self.sync_model_half()
            self.model._device_loop = None                   # (graphs captured over the old copies)
            self.peer.close()
            self.peer = None

    def skipped_steps(self):
        """Optimiser steps skipped by the loss scaler so far, and the current loss scale (host sync)."""
        h = self.hyper.cpu()
        return int(h[3]), struct.unpack("f", struct.pack("i", int(h[4])))[0]

    def sync_model_half(self):
        """Hand the fp16 working copies (always current) to the nn.Module view; no communication."""
        self.model.xyz_encoder.set_half_params(self.h_xyz)
        self.model.rgb_net.set_half_params(self.h_rgb)
        # the model's fused paths -- and the CUDA graphs captured over them (occupancy update, render loop) -- read
        # the weight image this trainer re-packs after every optimiser step: never a stale or reallocated copy
        self.model.adopt_image(self.w_image)

    def overflowed(self):
        """True if the last step hit the sample capacity (host sync; call sparingly)."""
        return bool(self.last_counter[2].item())

    def samples_last_step(self):
        return int(self.last_counter[0].item())

    def grow(self, factor=1.5):
        self.capacity = int(self.capacity * factor)
        self._alloc()

    def sync_model(self):
        """Make the nn.Module view consistent after direct parameter updates (hand over the fp16 copies; with a
        sharded optimiser also gather the fp32 master shards so that state_dict() is complete on every rank)."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.p_pad, self.p_shard, group=self.pg)
        self.sync_model_half()
