"""Vectorised torch-CPU restatement of the ``vren`` entry points -- TEST INFRASTRUCTURE.

This is the "same render math expressed as torch CPU ops" that BASELINE.json's
north_star names as the CPU baseline.  Every float op is a separate torch fp32
kernel (no fused multiply-add), mirroring oracle/c/vren_oracle.c op for op; the
two are cross-checked bit for bit in tests/test_oracle_consistency.py.

PARITY UNPINNED (see oracle/__init__.py).

Reference call sites restated here:
  ray_aabb_intersect   ngp_pl/models/custom_functions.py:8-29, rendering.py:27-29
  raymarching_train    ngp_pl/models/custom_functions.py:55-101
  raymarching_test     ngp_pl/models/rendering.py:79-83
  composite_train_fw   ngp_pl/models/custom_functions.py:116-146
  composite_train_bw   ngp_pl/models/custom_functions.py:148-159
  composite_test_fw    ngp_pl/models/rendering.py:97-101
  morton3D(_invert)    ngp_pl/models/networks.py:128,147,153
  packbits             ngp_pl/models/networks.py:251-252
"""
import numpy as np
import torch

SQRT3 = np.float32(1.73205080757)
F32 = torch.float32


def _f(v):
    """0-dim fp32 tensor so that scalar operands never promote to double."""
    return torch.tensor(v, dtype=F32)


# --------------------------------------------------------------------------- Morton / packbits
def _expand_bits(v):
    m = 0xFFFFFFFF
    v = ((v * 0x00010001) & m) & 0xFF0000FF
    v = ((v * 0x00000101) & m) & 0x0F00F00F
    v = ((v * 0x00000011) & m) & 0xC30C30C3
    v = ((v * 0x00000005) & m) & 0x49249249
    return v


def morton3D(coords):
    """coords (N,3) int32 -> (N) int32 Morton code  (networks.py:128)."""
    c = coords.to(torch.int64)
    out = _expand_bits(c[:, 0]) | (_expand_bits(c[:, 1]) << 1) | (_expand_bits(c[:, 2]) << 2)
    return out.to(torch.int32)


def _compact_bits(x):
    x = x & 0x49249249
    x = (x | (x >> 2)) & 0xC30C30C3
    x = (x | (x >> 4)) & 0x0F00F00F
    x = (x | (x >> 8)) & 0xFF0000FF
    x = (x | (x >> 16)) & 0x0000FFFF
    return x


def morton3D_invert(indices):
    """indices (N) int32 -> coords (N,3) int32  (networks.py:153)."""
    v = indices.to(torch.int64) & 0xFFFFFFFF
    return torch.stack([_compact_bits(v), _compact_bits(v >> 1), _compact_bits(v >> 2)], 1).to(torch.int32)


def packbits(density_grid, density_threshold, density_bitfield):
    """bit i of byte n = grid.flat[8n+i] > thr, written in place (networks.py:251-252)."""
    bits = (density_grid.reshape(-1, 8) > _f(density_threshold)).to(torch.uint8)
    weights = (1 << torch.arange(8)).to(torch.uint8)
    density_bitfield.copy_((bits * weights).sum(1).to(torch.uint8))
    return density_bitfield


# --------------------------------------------------------------------------- ray / AABB
def ray_aabb_intersect(rays_o, rays_d, centers, half_sizes, max_hits):
    """Slab test.  Returns hits_cnt (N) i32, hits_t (N,max_hits,2) f32, hits_voxel_idx (N,max_hits) i64."""
    rays_o = rays_o.to(F32); rays_d = rays_d.to(F32)
    N, V = rays_o.shape[0], centers.shape[0]
    inv = _f(1.0) / rays_d                                      # (N,3)
    lo = ((centers - half_sizes)[None] - rays_o[:, None]) * inv[:, None]   # (N,V,3)
    hi = ((centers + half_sizes)[None] - rays_o[:, None]) * inv[:, None]
    tmin = torch.minimum(lo, hi); tmax = torch.maximum(lo, hi)
    t1 = torch.maximum(torch.maximum(tmin[..., 0], tmin[..., 1]), tmin[..., 2])
    t2 = torch.minimum(torch.minimum(tmax[..., 0], tmax[..., 1]), tmax[..., 2])
    hit = ~(t1 > t2) & (t2 > 0)
    t1c = torch.maximum(t1, _f(0.0))
    hits_cnt = hit.sum(1).to(torch.int32)
    key = torch.where(hit, t1c, torch.full_like(t1c, float('inf')))
    order = torch.sort(key, dim=1, stable=True)[1][:, :max_hits]           # near to far
    if order.shape[1] < max_hits:
        pad = torch.zeros(N, max_hits - order.shape[1], dtype=torch.int64)
        valid = torch.cat([torch.gather(hit, 1, order), torch.zeros_like(pad, dtype=torch.bool)], 1)
        order = torch.cat([order, pad], 1)
    else:
        valid = torch.gather(hit, 1, order)
    hits_t = torch.stack([torch.gather(t1c, 1, order), torch.gather(t2, 1, order)], -1)
    hits_t = torch.where(valid[..., None], hits_t, torch.full_like(hits_t, -1.0))
    hits_idx = torch.where(valid, order, torch.full_like(order, -1))
    return hits_cnt, hits_t, hits_idx


# --------------------------------------------------------------------------- marcher
def _calc_dt(t, esf, max_samples, grid_size, scale):
    lo = SQRT3 / np.float32(max_samples)
    hi = np.float32(np.float32(SQRT3 * np.float32(2)) * np.float32(scale)) / np.float32(grid_size)
    return torch.clamp(t * _f(esf), _f(lo), _f(hi))


def _probe(o, d, d_inv, t, bitfield, cascades, scale, esf, grid_size, max_samples):
    """One DDA loop body for a batch of rays at parameter t."""
    G = grid_size
    x = o + t[:, None] * d                                        # mul then add, no fma
    dt = _calc_dt(t, esf, max_samples, G, scale)
    mx = torch.maximum(x[:, 0].abs(), torch.maximum(x[:, 1].abs(), x[:, 2].abs()))
    mip_pos = (torch.frexp(mx)[1].to(torch.int64) + 1).clamp(0, cascades - 1)
    mip_dt = torch.frexp(dt * _f(float(G)))[1].to(torch.int64).clamp(0, cascades - 1)
    mip = torch.maximum(mip_pos, mip_dt)
    mip_bound = torch.minimum(torch.ldexp(torch.ones_like(t), (mip - 1).to(torch.int32)), _f(scale))
    mip_bound_inv = _f(1.0) / mip_bound
    n = (_f(0.5) * (x * mip_bound_inv[:, None] + _f(1.0)) * _f(float(G))).clamp(_f(0.0), _f(G - 1.0)).to(torch.int32)
    idx = mip * (G ** 3) + morton3D(n).to(torch.int64)
    occ = (bitfield[idx // 8].to(torch.int64) >> (idx % 8)) & 1
    sgn = torch.copysign(torch.ones_like(d), d)
    G_inv = np.float32(1.0) / np.float32(G)
    tt = (((n.to(F32) + _f(0.5) + _f(0.5) * sgn) * _f(G_inv) * _f(2.0) - _f(1.0)) * mip_bound[:, None] - x) * d_inv
    t_target = t + torch.maximum(_f(0.0), torch.minimum(tt[:, 0], torch.minimum(tt[:, 1], tt[:, 2])))
    return x, dt, occ.bool(), t_target


def _march(rays_o, rays_d, t_start, t_end, bitfield, cascades, scale, esf, grid_size, max_samples,
           limit, check_nonneg):
    """Shared ladder-form marcher.

    The serial reference advances t with the SAME recurrence t <- t + calc_dt(t) whether it
    emits a sample or skips an empty cell, so the sequence of candidate parameters is
    path-independent ("the ladder").  A rung is probed iff no earlier probed-empty rung set a
    skip target beyond it.  This form vectorises over rays (and is what the warp-cooperative
    CUDA marcher parallelises over lanes); tests prove it bit-identical to the serial C loop.

    Returns per-emission arrays (ray_local_idx, x, t, dt) in emission order, the per-ray counts
    and, for the test marcher, the parameter after the last emitted sample.
    """
    N = rays_o.shape[0]
    d_inv = _f(1.0) / rays_d
    t = t_start.clone()
    cnt = torch.zeros(N, dtype=torch.int64)
    t_after = t_start.clone()
    skipping = torch.zeros(N, dtype=torch.bool)
    target = torch.zeros(N, dtype=F32)
    live = torch.arange(N)
    out_r, out_x, out_t, out_dt = [], [], [], []
    while live.numel() > 0:
        tl, sk = t[live], skipping[live]
        cond = (tl < t_end[live]) & (cnt[live] < limit)
        if check_nonneg:
            cond &= (tl >= 0)
        done = ~sk & ~cond
        live = live[~done]
        if live.numel() == 0:
            break
        tl, sk = t[live], skipping[live]
        pr = live[~sk]                                          # rays at the loop top: probe
        if pr.numel() > 0:
            x, dt, occ, tgt = _probe(rays_o[pr], rays_d[pr], d_inv[pr], t[pr], bitfield, cascades,
                                     scale, esf, grid_size, max_samples)
            e = pr[occ]
            out_r.append(e); out_x.append(x[occ]); out_t.append(t[pr][occ]); out_dt.append(dt[occ])
            cnt[e] += 1
            t_after[e] = t[e] + dt[occ]
            miss = pr[~occ]
            target[miss] = tgt[~occ]
            skipping[miss] = True
        # every live ray climbs one rung
        t[live] = t[live] + _calc_dt(t[live], esf, max_samples, grid_size, scale)
        skipping[live] = skipping[live] & (t[live] < target[live])
    if out_r:
        r = torch.cat(out_r); order = torch.sort(r, stable=True)[1]
        return r[order], torch.cat(out_x)[order], torch.cat(out_t)[order], torch.cat(out_dt)[order], cnt, t_after
    z = torch.zeros(0, dtype=F32)
    return torch.zeros(0, dtype=torch.int64), torch.zeros(0, 3), z, z, cnt, t_after


def raymarching_train(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor,
                      noise, grid_size, max_samples):
    """-> rays_a (N_rays,3) i64 [ray_idx,start,N], xyzs, dirs (N,3), deltas, ts (N), counter (2) i32.

    Deterministic packing (DESIGN.md "Sample order"): row r of rays_a is ray r and its samples
    start at the exclusive prefix sum of the counts in ray order.
    """
    rays_o = rays_o.to(F32); rays_d = rays_d.to(F32); hits_t = hits_t.to(F32)
    N = rays_o.shape[0]
    t1, t2 = hits_t[:, 0].clone(), hits_t[:, 1]
    hit = t1 >= 0
    dt0 = _calc_dt(t1, exp_step_factor, max_samples, grid_size, scale)
    t1 = torch.where(hit, t1 + dt0 * noise.to(F32), t1)
    r, x, t, dt, cnt, _ = _march(rays_o, rays_d, t1, t2, density_bitfield, cascades, scale,
                                 exp_step_factor, grid_size, max_samples, max_samples, True)
    start = torch.cumsum(cnt, 0) - cnt
    rays_a = torch.stack([torch.arange(N), start, cnt], 1)
    counter = torch.tensor([int(cnt.sum()), N], dtype=torch.int32)
    return rays_a, x, rays_d[r], dt, t, counter


def raymarching_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale,
                     exp_step_factor, grid_size, max_samples, N_samples):
    """hits_t (N_rays,2) is advanced IN PLACE (rendering.py:80).  Unused slots stay zero (rendering.py:87)."""
    A = alive_indices.shape[0]
    ro, rd = rays_o[alive_indices].to(F32), rays_d[alive_indices].to(F32)
    t1, t2 = hits_t[alive_indices, 0].clone(), hits_t[alive_indices, 1]
    r, x, t, dt, cnt, t_after = _march(ro, rd, t1, t2, density_bitfield, cascades, scale,
                                       exp_step_factor, grid_size, max_samples, N_samples, False)
    xyzs = torch.zeros(A, N_samples, 3); dirs = torch.zeros(A, N_samples, 3)
    deltas = torch.zeros(A, N_samples); ts = torch.zeros(A, N_samples)
    if r.numel() > 0:
        start = torch.cumsum(cnt, 0) - cnt
        slot = torch.arange(r.numel()) - start[r]
        xyzs[r, slot] = x; dirs[r, slot] = rd[r]; deltas[r, slot] = dt; ts[r, slot] = t
    hits_t[alive_indices, 0] = t_after
    return xyzs, dirs, deltas, ts, cnt.to(torch.int32)


# --------------------------------------------------------------------------- compositing
def _padded(rays_a, *arrs):
    start, N = rays_a[:, 1], rays_a[:, 2]
    maxN = int(N.max()) if N.numel() else 0
    k = torch.arange(max(maxN, 1))
    mask = k[None] < N[:, None]
    idx = torch.where(mask, start[:, None] + k[None], torch.zeros_like(k[None]))
    if arrs and arrs[0].shape[0] == 0:                              # no samples at all: gather from one dummy row
        arrs = [torch.zeros((1,) + tuple(a.shape[1:]), dtype=a.dtype) for a in arrs]
    return mask, idx, [a[idx] for a in arrs]


def _composite_terms(sigmas, rgbs, deltas, ts, rays_a, T_threshold):
    mask, idx, (s, c, dl, t) = _padded(rays_a, sigmas, rgbs, deltas, ts)
    a = _f(1.0) - torch.exp(-s * dl)
    a = torch.where(mask, a, torch.zeros_like(a))
    T_after = torch.cumprod(_f(1.0) - a, 1)                       # sequential: same order as the serial loop
    T_before = torch.cat([torch.ones_like(T_after[:, :1]), T_after[:, :-1]], 1)
    alive_after = torch.cumprod((T_after > _f(T_threshold)).to(torch.int64), 1).bool()
    incl = mask & torch.cat([torch.ones_like(alive_after[:, :1]), alive_after[:, :-1]], 1)
    w = torch.where(incl, a * T_before, torch.zeros_like(a))
    return mask, idx, incl, w, T_after, c, dl, t


def composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T_threshold):
    """-> opacity, depth, depth_sq (N_rays), rgb (N_rays,3), all fp32, indexed by rays_a[:,0]."""
    sigmas, rgbs, deltas, ts = (v.to(F32) for v in (sigmas, rgbs, deltas, ts))
    n_rays = rays_a.shape[0]
    _, _, _, w, _, c, _, t = _composite_terms(sigmas, rgbs, deltas, ts, rays_a, T_threshold)
    last = lambda v: torch.cumsum(v, 1)[:, -1]                    # sequential sum, serial order
    ray = rays_a[:, 0]
    opacity = torch.zeros(n_rays); depth = torch.zeros(n_rays); depth_sq = torch.zeros(n_rays)
    rgb = torch.zeros(n_rays, 3)
    opacity[ray] = last(w); depth[ray] = last(w * t); depth_sq[ray] = last(w * t * t)
    rgb[ray] = torch.stack([last(w * c[..., i]) for i in range(3)], 1)
    return opacity, depth, depth_sq, rgb


def composite_train_bw(dL_dopacity, dL_ddepth, dL_ddepth_sq, dL_drgb, sigmas, rgbs, deltas, ts, rays_a,
                       opacity, depth, depth_sq, rgb, T_threshold):
    """Analytic gradient exactly as SURVEY row a11 / A.5.  -> dL_dsigmas (N), dL_drgbs (N,3)."""
    sigmas, rgbs, deltas, ts = (v.to(F32) for v in (sigmas, rgbs, deltas, ts))
    mask, idx, incl, w, T_after, c, dl, t = _composite_terms(sigmas, rgbs, deltas, ts, rays_a, T_threshold)
    ray = rays_a[:, 0]
    cs = lambda v: torch.cumsum(v, 1)
    r = torch.stack([cs(w * c[..., i]) for i in range(3)], -1)
    d = cs(w * t); d2 = cs(w * t * t)
    g_rgb = dL_drgb[ray][:, None]                                  # (R,1,3)
    RGB = rgb[ray][:, None]
    dsig = dl * ((g_rgb * (c * T_after[..., None] - (RGB - r))).cumsum(-1)[..., -1]
                 + (dL_dopacity[ray] * (_f(1.0) - opacity[ray]))[:, None]
                 + dL_ddepth[ray][:, None] * (t * T_after - (depth[ray][:, None] - d))
                 + dL_ddepth_sq[ray][:, None] * (t * t * T_after - (depth_sq[ray][:, None] - d2)))
    drgb = g_rgb * w[..., None]
    N = sigmas.shape[0]
    dL_dsigmas = torch.zeros(N); dL_drgbs = torch.zeros(N, 3)
    dL_dsigmas[idx[incl]] = dsig[incl]
    dL_drgbs[idx[incl]] = drgb[incl]
    return dL_dsigmas, dL_drgbs


def composite_test_fw(sigmas, rgbs, deltas, ts, hits_t, alive_indices, T_threshold, N_eff_samples,
                      opacity, depth, rgb):
    """In place on alive_indices / opacity / depth / rgb (rendering.py:97-101)."""
    A, S = sigmas.shape
    k = torch.arange(S)
    mask = k[None] < N_eff_samples[:, None].to(torch.int64)
    r = alive_indices.clone()
    a = torch.where(mask, _f(1.0) - torch.exp(-sigmas.to(F32) * deltas.to(F32)), torch.zeros(A, S))
    T0 = _f(1.0) - opacity[r]
    T_after = torch.empty(A, S); T = T0.clone()                    # serial order: T = T0; T *= (1-a0); ...
    for s in range(S):
        T = T * (_f(1.0) - a[:, s]); T_after[:, s] = T
    T_before = torch.cat([T0[:, None], T_after[:, :-1]], 1)
    alive_after = torch.cumprod((T_after > _f(T_threshold)).to(torch.int64), 1).bool()
    incl = mask & torch.cat([torch.ones(A, 1, dtype=torch.bool), alive_after[:, :-1]], 1)
    w = torch.where(incl, a * T_before, torch.zeros(A, S))
    op, dp, col = opacity[r].clone(), depth[r].clone(), rgb[r].clone()
    for s in range(S):                                             # serial accumulation order
        col = col + w[:, s, None] * rgbs[:, s].to(F32)
        dp = dp + w[:, s] * ts[:, s]
        op = op + w[:, s]
    opacity[r] = op; depth[r] = dp; rgb[r] = col
    stopped = (incl & ~alive_after).any(1) | (N_eff_samples == 0)
    alive_indices[stopped] = -1
