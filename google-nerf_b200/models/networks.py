"""NGP: the field + occupancy-grid state holder with the attribute names, method signatures and
checkpoint keys of ngp_pl/models/networks.py:12-252.

state_dict keys: center, xyz_min, xyz_max, half_size (1,3); density_bitfield
(cascades*G^3/8) uint8; xyz_encoder.params, dir_encoder.params (0 elements), rgb_net.params (flat fp32);
`density_grid` (cascades, G^3) and `grid_coords` (G^3, 3) int32 are registered by the training system
(ngp_pl/train.py:73-77) -- or by ``init_grid_buffers()`` here -- and are dropped by utils.slim_ckpt.

Extra keyword arguments (reference constants as defaults):
  encoding: "Frequency" is what this fork's NGP builds (networks.py:49-53); "HashGrid" is the Instant-NGP
            configuration left commented at networks.py:39-47 and the one north_star names.
  log2_T, grid_size: networks.py:25,30.
"""
import numpy as np
import torch
from torch import nn

from .. import _lib as L
from .. import tinycudann as tcnn
from .. import vren
from torch.amp import custom_bwd, custom_fwd

from .custom_functions import TruncExp
from .rendering import NEAR_DISTANCE


class _FieldFn(torch.autograd.Function):
    """NGP.forward (networks.py:102-117) as ONE autograd node on the fused kernels: encode (hash-grid gather or
    frequency) -> b2n_field_mlp_fw, and b2n_field_mlp_bw -> hash-grid scatter on the way back.  No eager
    normalise / cat / exp launches; the only saved activations are enc and h.  Gradients are computed at tcnn's
    loss scale (128) and returned unscaled in fp32; a gradient that leaves the fp16 range on the way is reported as
    inf in the returned parameter gradient, which is what a GradScaler (precision=16, train.py:265) keys on."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, d, xyz_params, rgb_params, model):
        L.require_cuda(x, d, xyz_params, rgb_params)
        # a sync-free RayMarcher hands over buffers sized with headroom and the true sample count on the device
        n_dev = getattr(x, "_b2n_n_dev", None)
        x, d = x.contiguous(), d.contiguous()
        p16, image = model._fused_state(x.device)
        n, dev = x.shape[0], x.device
        enc = model._encode(x, p16, n_dev=n_dev)
        sigmas = torch.empty(n, device=dev); rgbs = torch.empty(n, 3, device=dev)
        h = torch.empty(n, 16, dtype=torch.float16, device=dev)
        L.call("b2n_field_mlp_fw", L.ptr(enc), model.k1, L.ptr(d), L.ptr(image), n, L.ptr(n_dev), L.ptr(sigmas),
               L.ptr(rgbs), L.ptr(h))
        ctx.model, ctx.n_dev = model, n_dev
        ctx.save_for_backward(x, d, enc, h, rgbs, image)
        return sigmas, rgbs.to(torch.float16)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dL_dsigmas, dL_drgbs):
        model = ctx.model
        x, d, enc, h, rgbs, image = ctx.saved_tensors
        n, dev, S = x.shape[0], x.device, tcnn.LOSS_SCALE
        xe = model.xyz_encoder
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        ds = (dL_dsigmas.float() * S).contiguous() if dL_dsigmas is not None else z(n)
        dc = (dL_drgbs.float() * S).contiguous() if dL_drgbs is not None else z(n, 3)
        g_xyz, g_rgb = z(xe.params.numel()), z(model.rgb_net.params.numel())
        hashed = model.encoding == "HashGrid"
        denc = torch.empty(n, 32, dtype=torch.float16, device=dev) if hashed else None
        found = torch.zeros(1, dtype=torch.int32, device=dev)
        nd = L.ptr(ctx.n_dev)
        L.call("b2n_field_mlp_bw", L.ptr(ds), L.ptr(dc), L.ptr(enc), model.k1, L.ptr(d), L.ptr(image), n, nd,
               L.ptr(rgbs), L.ptr(h), 1.0 / S, L.ptr(denc), L.ptr(g_xyz), L.ptr(g_rgb), None, 0, L.ptr(found))
        if hashed:
            L.call("b2n_hashgrid_bw", L.ptr(x), L.ptr(denc), 32, model._layout, n, nd, 1.0 / S,
                   L.ptr(g_xyz[xe.mlp.n_params:]), None)
        g_rgb[:1] += torch.where(found > 0, float("inf"), 0.0)
        return None, None, g_xyz, g_rgb, None


class NGP(nn.Module):
    def __init__(self, scale, num_levels=16, encoding="HashGrid", log2_T=19, grid_size=128):
        super().__init__()
        self.scale = scale
        self.register_buffer("center", torch.zeros(1, 3))
        self.register_buffer("xyz_min", -torch.ones(1, 3) * scale)
        self.register_buffer("xyz_max", torch.ones(1, 3) * scale)
        self.register_buffer("half_size", (self.xyz_max - self.xyz_min) / 2)

        # cascade k of the occupancy grid covers [-2^(k-1), 2^(k-1)]^3 (clipped to the scene box)
        self.cascades = max(1 + int(np.ceil(np.log2(2 * scale))), 1)
        self.grid_size = grid_size
        self.register_buffer("density_bitfield", torch.zeros(self.cascades * grid_size ** 3 // 8, dtype=torch.uint8))

        L_, F_, N_min = num_levels, 2, 16
        b = np.exp(np.log(2048 * scale / N_min) / (L_ - 1))
        self.encoding = encoding
        if encoding == "HashGrid":
            enc_cfg = {"otype": "HashGrid", "n_levels": L_, "n_features_per_level": F_,
                       "log2_hashmap_size": log2_T, "base_resolution": N_min, "per_level_scale": b}
        elif encoding == "Frequency":
            enc_cfg = {"otype": "Frequency", "n_frequencies": 12}
        else:
            raise ValueError(f"unknown encoding {encoding!r}")
        self.xyz_encoder = tcnn.NetworkWithInputEncoding(
            n_input_dims=3, n_output_dims=16, encoding_config=enc_cfg,
            network_config={"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None",
                            "n_neurons": 64, "n_hidden_layers": 1})
        self.dir_encoder = tcnn.Encoding(n_input_dims=3, encoding_config={"otype": "SphericalHarmonics", "degree": 4})
        self.rgb_net = tcnn.Network(
            n_input_dims=32, n_output_dims=3,
            network_config={"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid",
                            "n_neurons": 64, "n_hidden_layers": 2})
        self.sigma_act = TruncExp.apply
        # first-layer width of the fused field kernels: 16 levels x 2 features, or Frequency-12 padded to 80
        self.k1 = self.xyz_encoder.mlp.in_width
        # the fused tcgen05 kernels are built for these two widths (num_levels = 16, the default everywhere in the
        # reference, and Frequency-12); any other `--num_levels` (train_scannet.py:72, opt.py:51) runs module by module on
        # the generic kernels of the tinycudann drop-in -- same results, one launch per module
        self.fused = self.k1 in (32, 80)
        self._image = self._image_key = self._layout = None

    # ------------------------------------------------------------------ field
    def density(self, x, return_feat=False):
        """x (N,3) in [-scale, scale] -> sigmas (N) fp32 [, h (N,16) fp16]."""
        if self.fused and x.is_cuda and not (torch.is_grad_enabled() and self.xyz_encoder.params.requires_grad):
            sigmas, h = self._density_fused(x.contiguous().float(), want_h=return_feat)
            return (sigmas, h) if return_feat else sigmas
        x = (x - self.xyz_min) / (self.xyz_max - self.xyz_min)
        h = self.xyz_encoder(x)
        sigmas = self.sigma_act(h[:, 0]).float()
        if return_feat:
            return sigmas, h
        return sigmas

    def forward(self, x, d):
        """x (N,3) positions, d (N,3) directions -> sigmas (N) fp32, rgbs (N,3) fp16.  One fused autograd node
        (or, without grad, the bare kernels).  The reference normalises d IN PLACE (networks.py:113); here the
        normalisation happens inside the kernel and d is left as it was."""
        if not x.is_cuda:
            raise RuntimeError("google-nerf_b200 kernels need CUDA tensors (there is no CPU fallback)")
        if not self.fused:       # networks.py:111-115 verbatim on the drop-in modules
            sigmas, h = self.density(x, return_feat=True)
            d = d / torch.norm(d, dim=-1, keepdim=True)
            d = self.dir_encoder((d + 1) / 2)
            return sigmas, self.rgb_net(torch.cat([d, h], 1))
        xe, rn = self.xyz_encoder.params, self.rgb_net.params
        if torch.is_grad_enabled() and (xe.requires_grad or rn.requires_grad):
            return _FieldFn.apply(x, d, xe, rn, self)
        return self._forward_fused(x, d)

    def _fused_state(self, device):
        """(fp16 parameter copy of xyz_encoder, canonical weight image).  The image lives in ONE buffer per model that
        is never reallocated (captured CUDA graphs keep pointing at it) and is re-packed in place when the fp16
        copies change; a trainer may adopt it (`adopt_image`) and keep it current itself."""
        xe, rn = self.xyz_encoder, self.rgb_net
        p16, r16 = xe.half_params(), rn.half_params()
        if self._image is None or self._image.device != device:
            self._image = torch.empty(64 * self.k1 + 8192, dtype=torch.float16, device=device)
            self._image_key = None
        if self._layout is None and self.encoding == "HashGrid":
            self._layout = L.GridLayout.from_buffer_copy(xe.enc.layout)      # NGP.density's box normalisation folded in
            self._layout.x_offset = -float(self.scale)
            self._layout.x_scale = 1.0 / (2.0 * float(self.scale))
        key = (xe._p16_key, rn._p16_key, p16.data_ptr(), r16.data_ptr())
        if self._image_key != key:
            L.call("b2n_field_pack_weights", L.ptr(p16), L.ptr(r16), L.ptr(self._image), self.k1, None)
            self._image_key = key
        return p16, self._image

    # render / training-graph states cached on the module (captured CUDA graphs, static buffers, ctypes structures):
    # rebuilt on demand, never part of a copy or a pickle of the model
    _TRANSIENT = ("_whole_rays", "_whole_rays_pool", "_device_loop", "_train_graphs", "_image", "_image_key", "_layout")

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in self._TRANSIENT:
            if k in state:
                state[k] = {} if isinstance(state[k], dict) else None
        return state

    def adopt_image(self, image):
        """A trainer that re-packs `image` after every optimiser step hands it over: the model's fused paths (and
        graphs captured over them) read that buffer from now on."""
        xe, rn = self.xyz_encoder, self.rgb_net
        p16, r16 = xe.half_params(), rn.half_params()
        self._image = image
        self._image_key = (xe._p16_key, rn._p16_key, p16.data_ptr(), r16.data_ptr())
        self._fused_state(image.device)

    def _encode(self, x, p16, out=None, n_dev=None, unit_cube=False):
        """positions (N,3) fp32 (world, or already in [0,1]^3 when unit_cube) -> encoded features (N,k1) fp16."""
        xe = self.xyz_encoder
        n = x.shape[0]
        if out is None:
            out = torch.empty(n, self.k1, dtype=torch.float16, device=x.device)
        if self.encoding == "HashGrid":
            L.call("b2n_hashgrid_fw", L.ptr(x), L.ptr(p16[xe.mlp.n_params:]), xe.enc.layout if unit_cube else self._layout,
                   n, L.ptr(n_dev), L.ptr(out), self.k1)
        else:
            lo, ext = (0.0, 1.0) if unit_cube else (-float(self.scale), 2.0 * float(self.scale))
            L.call("b2n_frequency_fw", L.ptr(x), 12, n, L.ptr(n_dev), L.ptr(out), self.k1, lo, ext)
        return out

    def _density_fused(self, x, want_h=False, unit_cube=False):
        """sigma (and optionally h) through the encode kernel + the density half of the fused tcgen05 kernel."""
        p16, image = self._fused_state(x.device)
        n = x.shape[0]
        enc = self._encode(x, p16, unit_cube=unit_cube)
        sigmas = torch.empty(n, device=x.device)
        h = torch.empty(n, 16, dtype=torch.float16, device=x.device) if want_h else None
        L.call("b2n_field_mlp_fw", L.ptr(enc), self.k1, None, L.ptr(image), n, None, L.ptr(sigmas), None, L.ptr(h))
        return sigmas, h

    def _density_fused01(self, x01):
        """sigma at positions already normalised to [0,1]^3 (occupancy-grid update)."""
        return self._density_fused(x01, unit_cube=True)[0]

    def _forward_fused(self, x, d, rgb_fp32=False):
        """Inference path (no autograd): encode + ONE fused tcgen05 kernel for both MLPs, SH and the activations.
        Same outputs as the reference (sigmas fp32, rgbs fp16); the box normalisation and the direction normalisation
        happen inside the kernels, so `d` is left untouched here."""
        p16, image = self._fused_state(x.device)
        x = x.contiguous().float(); d = d.contiguous().float()
        n = x.shape[0]
        enc = self._encode(x, p16)
        sigmas = torch.empty(n, device=x.device); rgbs = torch.empty(n, 3, device=x.device)
        L.call("b2n_field_mlp_fw", L.ptr(enc), self.k1, L.ptr(d), L.ptr(image), n, None, L.ptr(sigmas), L.ptr(rgbs),
               None)
        return sigmas, (rgbs if rgb_fp32 else rgbs.to(torch.float16))   # values are fp16-rounded either way

    # ------------------------------------------------------------------ occupancy grid
    def init_grid_buffers(self):
        """What ngp_pl/train.py:73-77 does from outside: density_grid zeros and the (G^3,3) cell coordinates."""
        G, dev = self.grid_size, self.center.device
        if not hasattr(self, "density_grid"):
            self.register_buffer("density_grid", torch.zeros(self.cascades, G ** 3, device=dev))
        if not hasattr(self, "grid_coords"):
            r = torch.arange(G, dtype=torch.int32, device=dev)
            z, y, x = torch.meshgrid(r, r, r, indexing="ij")
            self.register_buffer("grid_coords", torch.stack([x, y, z], -1).reshape(-1, 3).contiguous())
        return self

    @torch.no_grad()
    def get_all_cells(self):
        """[(morton indices (G^3) i64, coords (G^3,3) i32)] * cascades."""
        indices = vren.morton3D(self.grid_coords).long()
        return [(indices, self.grid_coords)] * self.cascades

    @torch.no_grad()
    def sample_uniform_and_occupied_cells(self, M):
        """Per cascade: M uniformly random cells and M cells drawn from the currently occupied ones."""
        cells = []
        dev = self.density_grid.device
        for c in range(self.cascades):
            coords1 = torch.randint(self.grid_size, (M, 3), dtype=torch.int32, device=dev)
            indices1 = vren.morton3D(coords1)
            # M draws (with replacement) from the occupied cells.  The reference materialises them with
            # torch.nonzero + randint (networks.py:149-152), which forces a host sync for the dynamic shape; the
            # same distribution is sampled here on the device: rank k ~ U{0..n_occ-1} -> k-th occupied cell by a
            # binary search in the inclusive prefix count.
            cs = torch.cumsum(self.density_grid[c] > 0, 0, dtype=torch.int32)
            k = (torch.rand(M, device=dev) * cs[-1]).to(torch.int32)
            k = torch.minimum(k, (cs[-1] - 1).clamp(min=0))
            # a cascade without any occupied cell (cs[-1] == 0; the reference's randint(0) raises there,
            # networks.py:151) would make the search return G^3: clamp, the draw then degenerates to a harmless repeat
            indices2 = torch.searchsorted(cs, k, right=True).clamp_(max=self.grid_size ** 3 - 1)
            # the order of the cells is immaterial (each one scatters into its own slot), so evaluate them in Morton
            # order: neighbouring cells share hash-grid corners, which turns most of the coarse-level gathers of the
            # 1M-cell density query into L1 hits
            indices = torch.sort(torch.cat([indices1.int(), indices2.int()]))[0]
            cells.append((indices.long(), vren.morton3D_invert(indices)))
        return cells

    @torch.no_grad()
    def mark_invisible_cells(self, K, poses, img_wh, chunk=64 ** 3):
        """Cells no training camera sees (or that sit closer than NEAR_DISTANCE to one) get density -1 and are
        never updated (networks.py:159-214).  One kernel launch: a thread per cell walks the cameras (`chunk`, the
        reference's bmm chunk size, is accepted and unused)."""
        L.require_cuda(self.density_grid)
        K = K.to(self.density_grid.device, torch.float32).contiguous()
        poses = poses.to(self.density_grid.device, torch.float32)[:, :3, :4].contiguous()
        if not self.density_grid.is_contiguous():
            self.density_grid = self.density_grid.contiguous()
        L.call("b2n_mark_invisible_cells", L.ptr(K), L.ptr(poses), poses.shape[0], int(img_wh[0]), int(img_wh[1]),
               NEAR_DISTANCE, self.grid_size, self.cascades, float(self.scale), L.ptr(self.density_grid))

    @torch.no_grad()
    def update_density_grid(self, density_threshold, warmup=False, decay=0.95, erode=False, shard=None, reduce_tmp=None,
                            cells=None, noise=None):
        """EMA-max update of the cascaded density grid from fresh field samples, then re-pack the bitfield
        (networks.py:216-252).  The mean/threshold stays on the device (no .item() sync).

        Data parallel (SURVEY 8e "occupancy update"): shard=(rank, world) makes this rank evaluate 1/world of the
        cells and reduce_tmp(tmp) must max-reduce the sampled densities over the ranks before they are merged, so
        that all ranks keep identical grids while the field evaluations are shared out.
        cells / noise (parity tests): the per-cascade (indices, coords) lists and (n,3) jitter tensors to use instead of
        drawing them here (the reference draws both with torch's global RNG, networks.py:221-231)."""
        self._grid_eval(warmup, shard, cells, noise)
        if reduce_tmp is not None:
            reduce_tmp(self._grid_tmp)
        self._grid_commit(density_threshold, decay, erode)

    @torch.no_grad()
    def _grid_eval(self, warmup, shard=None, cells=None, noise_in=None):
        """Densities at jittered positions inside the selected cells -> self._grid_tmp (0 where not sampled)."""
        G = self.grid_size
        rank, world = shard if shard is not None else (0, 1)
        if getattr(self, "_grid_tmp", None) is None or self._grid_tmp.shape != self.density_grid.shape \
                or self._grid_tmp.device != self.density_grid.device:
            self._grid_tmp = torch.zeros_like(self.density_grid)
        tmp = self._grid_tmp
        tmp.zero_()
        if cells is not None:
            pass
        elif warmup:
            cells = self.get_all_cells()
            if world > 1:                                    # a contiguous slice of all cells per rank
                n = G ** 3
                lo_i, hi_i = rank * n // world, (rank + 1) * n // world
                cells = [(i[lo_i:hi_i], c[lo_i:hi_i]) for i, c in cells]
        else:
            cells = self.sample_uniform_and_occupied_cells(-(-(G ** 3 // 4) // world))
        lo, hi = -float(self.scale), float(self.scale)
        for c in range(self.cascades):
            indices, coords = cells[c]
            s = min(2 ** (c - 1), self.scale)
            noise = torch.rand(coords.shape[0], 3, device=coords.device) if noise_in is None else noise_in[c].contiguous()
            xyz01 = torch.empty(coords.shape[0], 3, device=coords.device)
            L.call("b2n_grid_cell_positions", L.ptr(coords.contiguous()), L.ptr(noise), coords.shape[0], G, float(s),
                   lo, hi, 1, L.ptr(xyz01))
            sigmas = self._density_fused01(xyz01) if self.fused else torch.exp(self.xyz_encoder(xyz01)[:, 0].float())
            L.call("b2n_grid_scatter", L.ptr(indices.contiguous()), L.ptr(sigmas), indices.shape[0], L.ptr(tmp[c]),
                   tmp.shape[1])

    @torch.no_grad()
    def _grid_commit(self, density_threshold, decay=0.95, erode=False):
        """grid = max(decay * grid, tmp) where grid >= 0, optional erosion, threshold, bitfield."""
        G, tmp = self.grid_size, self._grid_tmp
        if not self.density_grid.is_contiguous():
            self.density_grid = self.density_grid.contiguous()
        L.call("b2n_grid_ema", L.ptr(self.density_grid), L.ptr(tmp), self.density_grid.numel(), float(decay))
        if erode:
            grid = self.density_grid.view(self.cascades, G, G, G)
            maxpool = torch.nn.functional.max_pool3d(grid, kernel_size=3, stride=1, padding=1)
            local_max = (grid == maxpool) & (maxpool > 0)
            self.density_grid[local_max.view(self.cascades, -1)] *= decay
        ws = torch.empty(3, dtype=torch.float64, device=tmp.device)
        stats = torch.empty(3, dtype=torch.float32, device=tmp.device)
        L.call("b2n_grid_threshold", L.ptr(self.density_grid), self.density_grid.numel(), float(density_threshold),
               L.ptr(ws), L.ptr(stats))
        vren.packbits(self.density_grid, float(density_threshold), self.density_bitfield, threshold_dev=stats)
        self._grid_stats = stats                                        # [threshold used, mean, #positive]
