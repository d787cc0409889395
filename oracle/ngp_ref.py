"""CPU restatement of NGP + render() + one training step -- TEST INFRASTRUCTURE and the CPU baseline.

Follows, line by line, the host logic of
  ngp_pl/models/networks.py:12-117,216-252   (NGP, density, forward, update_density_grid)
  ngp_pl/models/rendering.py:12-166          (render, __render_rays_train, __render_rays_test)
  ngp_pl/losses.py:26-40                     (NeRFLoss)
  ngp_pl/train.py:112,144-170                (FusedAdam eps=1e-15, training_step)
with the kernels replaced by oracle/vren_ref.py and oracle/tcnn_ref.py.  PARITY UNPINNED.
"""
import math

import numpy as np
import torch

from . import tcnn_ref as T
from . import vren_ref as V

MAX_SAMPLES = 1024
NEAR_DISTANCE = 0.05


class TruncExp(torch.autograd.Function):
    """ngp_pl/models/custom_functions.py:162-173."""
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        return g * torch.exp(ctx.saved_tensors[0].clamp(-15, 15))


class NGPRef:
    """State + field of ngp_pl/models/networks.py::NGP on CPU tensors."""

    def __init__(self, scale, encoding="HashGrid", num_levels=16, log2_T=19, grid_size=128, seed=1337,
                 dtype=torch.float32):
        self.scale = scale
        self.center = torch.zeros(1, 3)
        self.xyz_min = -torch.ones(1, 3) * scale
        self.xyz_max = torch.ones(1, 3) * scale
        self.half_size = (self.xyz_max - self.xyz_min) / 2
        self.cascades = max(1 + int(np.ceil(np.log2(2 * scale))), 1)          # networks.py:24
        self.grid_size = grid_size
        self.density_bitfield = torch.zeros(self.cascades * grid_size ** 3 // 8, dtype=torch.uint8)
        self.density_grid = torch.zeros(self.cascades, grid_size ** 3)
        self.encoding = encoding
        g = torch.Generator().manual_seed(seed)
        if encoding == "HashGrid":
            b = np.exp(np.log(2048 * scale / 16) / (num_levels - 1))              # networks.py:31
            self.layout = T.hashgrid_layout(num_levels, 2, log2_T, 16, b)
            enc_w, n_table = num_levels * 2, self.layout["n_params"]
        else:
            self.layout, enc_w, n_table = None, 80, 0
        self.sigma_shapes = T.mlp_layout(enc_w, 16, 64, 1)
        self.rgb_shapes = T.mlp_layout(32, 3, 64, 2)
        n_mlp = T.mlp_n_params(self.sigma_shapes)
        xyz = torch.zeros(n_mlp + n_table)
        T.xavier_uniform_(xyz, self.sigma_shapes, g)
        if n_table:
            xyz[n_mlp:] = (torch.rand(n_table, generator=g) * 2 - 1) * 1e-4
        rgb = T.xavier_uniform_(torch.zeros(T.mlp_n_params(self.rgb_shapes)), self.rgb_shapes, g)
        self.n_mlp = n_mlp
        self.xyz_params = xyz.to(dtype).requires_grad_(True)                      # "xyz_encoder.params"
        self.rgb_params = rgb.to(dtype).requires_grad_(True)                      # "rgb_net.params"

    # networks.py:87-100
    def density(self, x, return_feat=False):
        x = (x - self.xyz_min) / (self.xyz_max - self.xyz_min)
        if self.encoding == "HashGrid":
            enc = T.hashgrid_forward(x, self.xyz_params[self.n_mlp:], self.layout)
        else:
            enc = T.frequency_forward(x)
        h = T.mlp_forward(enc, self.xyz_params[:self.n_mlp], self.sigma_shapes, 16)
        sigmas = TruncExp.apply(h[:, 0].float())
        return (sigmas, h) if return_feat else sigmas

    # networks.py:102-117 (d is normalised IN PLACE, as in the reference)
    def forward(self, x, d):
        sigmas, h = self.density(x, return_feat=True)
        d /= torch.norm(d, dim=-1, keepdim=True)
        sh = T.sh4_forward((d + 1) / 2)
        rgbs = T.mlp_forward(torch.cat([sh.to(h.dtype), h], 1), self.rgb_params, self.rgb_shapes, 3, "Sigmoid")
        return sigmas, rgbs

    __call__ = forward

    # networks.py:216-252; rng supplies the cell jitter and the cell sampling
    @torch.no_grad()
    def update_density_grid(self, density_threshold, warmup=False, decay=0.95, erode=False, rng=None, cells=None):
        G = self.grid_size
        tmp = torch.zeros_like(self.density_grid)
        if cells is None:
            coords = _grid_coords(G)
            cells = [(V.morton3D(coords).long(), coords)] * self.cascades
        for c in range(self.cascades):
            indices, coords = cells[c]
            s = min(2 ** (c - 1), self.scale)
            half = s / G
            xyzs_w = (coords / (G - 1) * 2 - 1) * (s - half)
            noise = torch.rand(xyzs_w.shape, generator=rng) if not isinstance(rng, (list, tuple)) else rng[c]
            xyzs_w = xyzs_w + (noise * 2 - 1) * half
            tmp[c, indices] = self.density(xyzs_w)
        self.density_grid = torch.where(self.density_grid < 0, self.density_grid,
                                        torch.maximum(self.density_grid * decay, tmp))
        mean_density = self.density_grid[self.density_grid > 0].mean().item()
        V.packbits(self.density_grid, min(mean_density, density_threshold), self.density_bitfield)


@torch.no_grad()
def mark_invisible_cells(model, K, poses, img_wh, chunk=64 ** 3):
    """networks.py:159-214 on CPU tensors: density_grid <- 0 where some camera sees the cell centre at depth >=
    NEAR_DISTANCE and no camera has it in view closer than that, else -1."""
    G = model.grid_size
    w2c_R = poses[:, :3, :3].permute(0, 2, 1)
    w2c_T = torch.bmm(-w2c_R, poses[:, :3, 3:])
    coords = _grid_coords(G)
    indices = V.morton3D(coords).long()
    for c in range(model.cascades):
        for i in range(0, len(indices), chunk):
            xyzs = coords[i:i + chunk] / (G - 1) * 2 - 1
            s = min(2 ** (c - 1), model.scale)
            half_grid_size = s / G
            xyzs_w = (xyzs * (s - half_grid_size)).T
            xyzs_w = xyzs_w.unsqueeze(0).repeat(w2c_R.shape[0], 1, 1)
            xyzs_c = torch.bmm(w2c_R, xyzs_w) + w2c_T
            uvd = K @ xyzs_c
            uv = uvd[:, :2] / uvd[:, 2:]
            in_image = (uvd[:, 2] >= 0) & (uv[:, 0] >= 0) & (uv[:, 0] < img_wh[0]) & (uv[:, 1] >= 0) & (uv[:, 1] < img_wh[1])
            covered = ((uvd[:, 2] >= NEAR_DISTANCE) & in_image).any(0)
            too_near = ((uvd[:, 2] < NEAR_DISTANCE) & in_image).any(0)
            model.density_grid[c, indices[i:i + chunk]] = torch.where(covered & ~too_near, 0., -1.)


def _grid_coords(G):
    r = torch.arange(G, dtype=torch.int32)
    z, y, x = torch.meshgrid(r, r, r, indexing="ij")
    return torch.stack([x, y, z], -1).reshape(-1, 3)


def render(model, rays_o, rays_d, noise=None, **kwargs):
    """rendering.py:12-39.  ``noise`` is the per-ray jitter the reference draws with torch.rand_like
    (custom_functions.py:84); the oracle takes it as an input so CUDA and CPU share it."""
    rays_o = rays_o.contiguous(); rays_d = rays_d.contiguous()
    _, hits_t, _ = V.ray_aabb_intersect(rays_o, rays_d, model.center, model.half_size, 1)
    hits_t[(hits_t[:, 0, 0] >= 0) & (hits_t[:, 0, 0] < NEAR_DISTANCE), 0, 0] = NEAR_DISTANCE
    if kwargs.get("test_time", False):
        return _render_rays_test(model, rays_o, rays_d, hits_t, **kwargs)
    return _render_rays_train(model, rays_o, rays_d, hits_t, noise, **kwargs)


def _render_rays_train(model, rays_o, rays_d, hits_t, noise, **kwargs):
    esf = kwargs.get("exp_step_factor", 0.)
    if noise is None:
        noise = torch.rand(rays_o.shape[0])
    rays_a, xyzs, dirs, deltas, ts, counter = V.raymarching_train(
        rays_o, rays_d, hits_t[:, 0], model.density_bitfield, model.cascades, model.scale, esf, noise,
        model.grid_size, MAX_SAMPLES)
    results = {"total_samples": int(counter[0])}
    sigmas, rgbs = model(xyzs, dirs)
    opacity, depth, depth_sq, rgb = CompositeTrain.apply(sigmas, rgbs.float().contiguous(), deltas, ts, rays_a,
                                                         kwargs.get("T_threshold", 1e-4))
    bg = 1.0 if esf == 0 else 0.0
    results.update(opacity=opacity, depth=depth, depth_sq=depth_sq, rgb=rgb + bg * (1 - opacity)[:, None])
    results["_samples"] = (rays_a, xyzs, dirs, deltas, ts, sigmas, rgbs)
    return results


class CompositeTrain(torch.autograd.Function):
    """custom_functions.py:116-159 over the oracle kernels."""
    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, ts, rays_a, T_threshold):
        out = V.composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T_threshold)
        ctx.save_for_backward(sigmas, rgbs, deltas, ts, rays_a, *out)
        ctx.T_threshold = T_threshold
        return out

    @staticmethod
    def backward(ctx, gO, gD, gD2, gRGB):
        sigmas, rgbs, deltas, ts, rays_a, O, D, D2, RGB = ctx.saved_tensors
        ds, dc = V.composite_train_bw(gO, gD, gD2, gRGB, sigmas, rgbs, deltas, ts, rays_a, O, D, D2, RGB,
                                      ctx.T_threshold)
        return ds, dc, None, None, None, None


@torch.no_grad()
def _render_rays_test(model, rays_o, rays_d, hits_t, **kwargs):
    """rendering.py:42-114, the host loop restated verbatim."""
    esf = kwargs.get("exp_step_factor", 0.)
    N_rays = len(rays_o)
    opacity = torch.zeros(N_rays); depth = torch.zeros(N_rays); rgb = torch.zeros(N_rays, 3)
    samples = total_samples = 0
    alive = torch.arange(N_rays)
    min_samples = 1 if esf == 0 else 4
    hits = hits_t[:, 0]
    while samples < MAX_SAMPLES:
        N_alive = len(alive)
        if N_alive == 0:
            break
        N_samples = max(min(N_rays // N_alive, 64), min_samples)
        samples += N_samples
        xyzs, dirs, deltas, ts, N_eff = V.raymarching_test(rays_o, rays_d, hits, alive, model.density_bitfield,
                                                           model.cascades, model.scale, esf, model.grid_size,
                                                           MAX_SAMPLES, N_samples)
        total_samples += int(N_eff.sum())
        xyzs = xyzs.reshape(-1, 3); dirs = dirs.reshape(-1, 3)
        valid = ~torch.all(dirs == 0, dim=1)
        if valid.sum() == 0:
            break
        sigmas = torch.zeros(len(xyzs)); rgbs = torch.zeros(len(xyzs), 3)
        _s, _c = model(xyzs[valid], dirs[valid])
        sigmas[valid], rgbs[valid] = _s.float(), _c.float()
        V.composite_test_fw(sigmas.view(-1, N_samples), rgbs.view(-1, N_samples, 3), deltas, ts, hits, alive,
                            kwargs.get("T_threshold", 1e-4), N_eff, opacity, depth, rgb)
        alive = alive[alive >= 0]
    bg = 1.0 if esf == 0 else 0.0
    return {"opacity": opacity, "depth": depth, "rgb": rgb + bg * (1 - opacity)[:, None],
            "total_samples": total_samples}


def nerf_loss(results, target_rgb, lambda_opa=1e-3):
    """losses.py:32-40 followed by train.py:160 (sum of the means)."""
    d_rgb = (results["rgb"] - target_rgb) ** 2
    o = results["opacity"] + 1e-10
    d_opa = lambda_opa * (-o * torch.log(o))
    return d_rgb.mean() + d_opa.mean()


def shiftscale_inv_depthloss(disp_pred, disp_gt):
    """losses.py:5-23 verbatim."""
    t_pred = torch.median(disp_pred)
    s_pred = torch.mean(torch.abs(disp_pred - t_pred))
    t_gt = torch.median(disp_gt)
    s_gt = torch.mean(torch.abs(disp_gt - t_gt))
    disp_pred_n = (disp_pred - t_pred) / s_pred
    disp_gt_n = (disp_gt - t_gt) / s_gt
    return (disp_pred_n - disp_gt_n) ** 2


def depth_prior_loss(results, prior_disp, lambda_depth):
    """The SCADE-style depth-prior term on top of NeRFLoss: lambda * mean(shiftscale_inv_depthloss(1/depth, prior))
    over the rays that have a prior (> 0) and a rendered depth (> 1e-6)."""
    valid = (prior_disp > 0) & (results["depth"].detach() > 1e-6)
    return lambda_depth * shiftscale_inv_depthloss(1.0 / results["depth"][valid], prior_disp[valid]).mean()


class AdamRef:
    """apex FusedAdam(lr, eps=1e-15) semantics (train.py:112): betas (0.9,0.999), bias correction, no decay."""

    def __init__(self, params, lr=1e-2, eps=1e-15, betas=(0.9, 0.999)):
        self.params, self.lr, self.eps, self.betas, self.t = params, lr, eps, betas, 0
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]

    @torch.no_grad()
    def step(self):
        self.t += 1
        b1, b2 = self.betas
        c1, c2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        for p, m, v in zip(self.params, self.m, self.v):
            g = p.grad
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            p.sub_(self.lr * (m / c1) / ((v / c2).sqrt() + self.eps))
            p.grad = None


def train_step(model, opt, rays_o, rays_d, target_rgb, noise, **kwargs):
    """One training_step of train.py:144-170 (without the periodic grid update)."""
    res = render(model, rays_o, rays_d.clone(), noise=noise, **kwargs)
    loss = nerf_loss(res, target_rgb)
    loss.backward()
    opt.step()
    return loss.item(), res
