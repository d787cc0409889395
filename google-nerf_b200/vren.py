"""Drop-in for the reference's `vren` extension module: the same ten functions with the same positional
signatures, tensor shapes, dtypes and in-place contracts, over libb2n.so.

Call sites replaced (paths under /root/reference/ngp_pl):
  ray_aabb_intersect / ray_sphere_intersect   models/custom_functions.py:29,52
  raymarching_train                            models/custom_functions.py:86-90
  raymarching_test                             models/rendering.py:79-83
  composite_train_fw / composite_train_bw      models/custom_functions.py:140-142,153-158
  composite_test_fw                            models/rendering.py:97-100
  morton3D / morton3D_invert / packbits        models/networks.py:128,147,153,251-252

Differences from upstream that are visible but allowed by the reference's own use of the outputs:
  * raymarching_train packs samples deterministically (ray r owns row r of rays_a, start = exclusive prefix
    sum) instead of in atomic-arrival order, and returns buffers sized to the true total instead of
    N_rays*max_samples zero-filled rows (custom_functions.py:92-97 slices to counter[0] anyway).
"""
import torch

from . import _lib as L

_f32 = torch.float32


def _prep(t, dtype=_f32):
    L.require_cuda(t)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _intersect(fn, rays_o, rays_d, centers, sizes, max_hits):
    rays_o, rays_d, centers, sizes = (_prep(v) for v in (rays_o, rays_d, centers, sizes))
    n, dev = rays_o.shape[0], rays_o.device
    hits_cnt = torch.empty(n, dtype=torch.int32, device=dev)
    hits_t = torch.empty(n, max_hits, 2, dtype=_f32, device=dev)
    hits_idx = torch.empty(n, max_hits, dtype=torch.int64, device=dev)
    L.call(fn, L.ptr(rays_o), L.ptr(rays_d), L.ptr(centers), L.ptr(sizes), n, centers.shape[0], int(max_hits),
           L.ptr(hits_cnt), L.ptr(hits_t), L.ptr(hits_idx))
    return hits_cnt, hits_t, hits_idx


def ray_aabb_intersect(rays_o, rays_d, centers, half_sizes, max_hits):
    """-> hits_cnt (N_rays) i32, hits_t (N_rays,max_hits,2) f32 (-1 = miss), hits_voxel_idx (N_rays,max_hits) i64."""
    return _intersect("b2n_ray_aabb_intersect", rays_o, rays_d, centers, half_sizes, max_hits)


def ray_sphere_intersect(rays_o, rays_d, centers, radii, max_hits):
    return _intersect("b2n_ray_sphere_intersect", rays_o, rays_d, centers, radii.reshape(-1), max_hits)


def morton3D(coords):
    coords = _prep(coords, torch.int32)
    out = torch.empty(coords.shape[0], dtype=torch.int32, device=coords.device)
    L.call("b2n_morton3D", L.ptr(coords), coords.shape[0], L.ptr(out))
    return out


def morton3D_invert(indices):
    indices = _prep(indices, torch.int32)
    out = torch.empty(indices.shape[0], 3, dtype=torch.int32, device=indices.device)
    L.call("b2n_morton3D_invert", L.ptr(indices), indices.shape[0], L.ptr(out))
    return out


def packbits(density_grid, density_threshold, density_bitfield, threshold_dev=None):
    """In place on density_bitfield (uint8, one bit per cell, little bit order)."""
    L.require_cuda(density_grid, density_bitfield)
    assert density_grid.dtype == _f32 and density_grid.is_contiguous() and density_bitfield.is_contiguous()
    assert density_bitfield.numel() * 8 == density_grid.numel()
    L.call("b2n_packbits", L.ptr(density_grid), density_bitfield.numel(), float(density_threshold),
           L.ptr(threshold_dev), L.ptr(density_bitfield))


def raymarching_train_count(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise,
                            grid_size, max_samples, capacity=-1, serial=False):
    """First half of raymarching_train: -> rays_a (N_rays,3) i64, counter (4) i32 (device), workspace; no host sync.
    serial=True: the thread-per-ray count pass (same results; see b2n_raymarching_train_count_serial)."""
    n, dev = rays_o.shape[0], rays_o.device
    rays_a = torch.empty(n, 3, dtype=torch.int64, device=dev)
    counter = torch.empty(4, dtype=torch.int32, device=dev)
    workspace = torch.empty(n, 64, dtype=torch.int32, device=dev)        # per-chunk emission masks for the write pass
    L.call("b2n_raymarching_train_count_serial" if serial else "b2n_raymarching_train_count",
           L.ptr(rays_o), L.ptr(rays_d), L.ptr(hits_t), L.ptr(density_bitfield),
           int(cascades), float(scale), float(exp_step_factor), L.ptr(noise), int(grid_size), int(max_samples),
           n, int(capacity), L.ptr(rays_a), L.ptr(counter), L.ptr(workspace))
    return rays_a, counter, workspace


def raymarching_train_write(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise,
                            grid_size, max_samples, rays_a, n_rows, workspace=None):
    dev = rays_o.device
    xyzs = torch.empty(n_rows, 3, dtype=_f32, device=dev)
    dirs = torch.empty(n_rows, 3, dtype=_f32, device=dev)
    deltas = torch.empty(n_rows, dtype=_f32, device=dev)
    ts = torch.empty(n_rows, dtype=_f32, device=dev)
    L.call("b2n_raymarching_train_write", L.ptr(rays_o), L.ptr(rays_d), L.ptr(hits_t), L.ptr(density_bitfield),
           int(cascades), float(scale), float(exp_step_factor), L.ptr(noise), int(grid_size), int(max_samples),
           rays_o.shape[0], L.ptr(rays_a), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas), L.ptr(ts), L.ptr(workspace))
    return xyzs, dirs, deltas, ts


def raymarching_train(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise,
                      grid_size, max_samples):
    """-> rays_a (N_rays,3) i64, xyzs, dirs (N,3), deltas, ts (N), counter (i32; counter[0] = N)."""
    rays_o, rays_d, hits_t, noise = (_prep(v) for v in (rays_o, rays_d, hits_t, noise))
    L.require_cuda(density_bitfield)
    args = (rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise, grid_size, max_samples)
    rays_a, counter, workspace = raymarching_train_count(*args)
    total = int(counter[0].item())      # the one host sync the reference API forces (custom_functions.py:92)
    xyzs, dirs, deltas, ts = raymarching_train_write(*args, rays_a, total, workspace)
    return rays_a, xyzs, dirs, deltas, ts, counter


def raymarching_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale, exp_step_factor,
                     grid_size, max_samples, N_samples):
    """hits_t (N_rays,2) is advanced IN PLACE.  -> xyzs, dirs (N_alive,N_samples,3), deltas, ts
    (N_alive,N_samples), N_eff_samples (N_alive) i32; unused slots are zero."""
    L.require_cuda(rays_o, rays_d, hits_t, alive_indices, density_bitfield)
    assert hits_t.dtype == _f32 and hits_t.is_contiguous(), "hits_t must be a contiguous fp32 (N_rays,2) view"
    assert alive_indices.dtype == torch.int64
    rays_o, rays_d = _prep(rays_o), _prep(rays_d)
    alive_indices = alive_indices.contiguous()
    a, dev = alive_indices.shape[0], rays_o.device
    xyzs = torch.empty(a, N_samples, 3, dtype=_f32, device=dev)
    dirs = torch.empty(a, N_samples, 3, dtype=_f32, device=dev)
    deltas = torch.empty(a, N_samples, dtype=_f32, device=dev)
    ts = torch.empty(a, N_samples, dtype=_f32, device=dev)
    n_eff = torch.empty(a, dtype=torch.int32, device=dev)
    L.call("b2n_raymarching_test", L.ptr(rays_o), L.ptr(rays_d), L.ptr(hits_t), L.ptr(alive_indices),
           L.ptr(density_bitfield), int(cascades), float(scale), float(exp_step_factor), int(grid_size),
           int(max_samples), int(N_samples), a, L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas), L.ptr(ts), L.ptr(n_eff))
    return xyzs, dirs, deltas, ts, n_eff


def composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T_threshold):
    sigmas, rgbs, deltas, ts = (_prep(v) for v in (sigmas, rgbs, deltas, ts))
    rays_a = _prep(rays_a, torch.int64)
    n, dev = rays_a.shape[0], sigmas.device
    # zeros: rows of rays that are absent from rays_a keep 0 like upstream's zero-initialised outputs
    opacity = torch.zeros(n, dtype=_f32, device=dev)
    depth = torch.zeros(n, dtype=_f32, device=dev)
    depth_sq = torch.zeros(n, dtype=_f32, device=dev)
    rgb = torch.zeros(n, 3, dtype=_f32, device=dev)
    L.call("b2n_composite_train_fw", L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas), L.ptr(ts), L.ptr(rays_a),
           float(T_threshold), n, L.ptr(opacity), L.ptr(depth), L.ptr(depth_sq), L.ptr(rgb))
    return opacity, depth, depth_sq, rgb


def composite_train_bw(dL_dopacity, dL_ddepth, dL_ddepth_sq, dL_drgb, sigmas, rgbs, deltas, ts, rays_a,
                       opacity, depth, depth_sq, rgb, T_threshold):
    g = [_prep(v) for v in (dL_dopacity, dL_ddepth, dL_ddepth_sq, dL_drgb)]
    sigmas, rgbs, deltas, ts = (_prep(v) for v in (sigmas, rgbs, deltas, ts))
    rays_a = _prep(rays_a, torch.int64)
    outs = [_prep(v) for v in (opacity, depth, depth_sq, rgb)]
    N, dev = sigmas.shape[0], sigmas.device
    # the kernel writes every sample a ray owns (zeros after an early stop); rows no ray owns are never read downstream
    dL_dsigmas = torch.empty(N, dtype=_f32, device=dev)
    dL_drgbs = torch.empty(N, 3, dtype=_f32, device=dev)
    L.call("b2n_composite_train_bw", *[L.ptr(v) for v in g], L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas), L.ptr(ts),
           L.ptr(rays_a), *[L.ptr(v) for v in outs], float(T_threshold), rays_a.shape[0], L.ptr(dL_dsigmas),
           L.ptr(dL_drgbs), None, None)
    return dL_dsigmas, dL_drgbs


def composite_test_fw(sigmas, rgbs, deltas, ts, hits_t, alive_indices, T_threshold, N_eff_samples,
                      opacity, depth, rgb):
    """In place on alive_indices (-1 = converged), opacity, depth, rgb."""
    L.require_cuda(sigmas, rgbs, alive_indices, opacity, depth, rgb)
    sigmas, rgbs, deltas, ts = (_prep(v) for v in (sigmas, rgbs, deltas, ts))
    for t in (alive_indices, opacity, depth, rgb, N_eff_samples):
        assert t.is_contiguous()
    assert alive_indices.dtype == torch.int64 and N_eff_samples.dtype == torch.int32
    a, s = sigmas.shape
    L.call("b2n_composite_test_fw", L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas), L.ptr(ts), L.ptr(hits_t),
           L.ptr(alive_indices), float(T_threshold), L.ptr(N_eff_samples), int(s), a, L.ptr(opacity), L.ptr(depth),
           L.ptr(rgb))
