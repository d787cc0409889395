"""Synthetic NeRF-synthetic- / ScanNet-shaped workloads (datasets are unavailable offline).

Shapes follow the reference datasets: ``directions (H*W,3)`` un-normalised ``((u-cx+.5)/fx, (v-cy+.5)/fy, 1)``
(ngp_pl/datasets/ray_utils.py:33-35), ``poses (N,3,4)`` camera-to-world with positions divided by
``2*scale`` (ngp_pl/datasets/nsvf.py:86-87), batches ``{img_idxs, pix_idxs, rgb}`` (ngp_pl/datasets/base.py:24-40).
The scene is an analytic union of three spheres and a box so that ground-truth colours and the
occupancy grid are both closed-form.
"""
import math

import numpy as np
import torch

# primitives in units of `scale` (so the same scene fits any bounding box)
_SPHERES = [((0.00, 0.05, -0.10), 0.42, (0.85, 0.25, 0.20)),
            ((0.45, -0.30, 0.25), 0.24, (0.20, 0.65, 0.30)),
            ((-0.40, 0.35, 0.30), 0.20, (0.25, 0.35, 0.85))]
_BOX = ((0.0, 0.0, -0.62), (0.75, 0.75, 0.08), (0.80, 0.75, 0.55))   # centre, half size, albedo


def inside(xyz, scale):
    """(N,3) world positions -> bool occupancy of the analytic scene."""
    p = xyz / scale
    occ = torch.zeros(p.shape[0], dtype=torch.bool, device=p.device)
    for c, r, _ in _SPHERES:
        occ |= ((p - torch.tensor(c, device=p.device)) ** 2).sum(-1) <= r * r
    c, h, _ = _BOX
    occ |= ((p - torch.tensor(c, device=p.device)).abs() <= torch.tensor(h, device=p.device)).all(-1)
    return occ


def _morton(coords):
    def expand(v):
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v
    c = coords.to(torch.int64)
    return expand(c[:, 0]) | (expand(c[:, 1]) << 1) | (expand(c[:, 2]) << 2)


def grid_coords(G, device="cpu"):
    """(G^3,3) int32 cell coordinates, same content as kornia.create_meshgrid3d(G,G,G,False).reshape(-1,3)
    as registered at ngp_pl/train.py:76-77 (x fastest)."""
    r = torch.arange(G, dtype=torch.int32, device=device)
    z, y, x = torch.meshgrid(r, r, r, indexing="ij")
    return torch.stack([x, y, z], -1).reshape(-1, 3)


def density_grid(scale, cascades, G=128, occupied_value=10.0, coarse=None, device="cpu"):
    """(cascades, G^3) fp32 grid in Morton order; cell centres as in ngp_pl/models/networks.py:227-229.
    ``coarse=64`` rasterises at 64^3 and replicates 2x per axis."""
    coords = grid_coords(G, device)
    idx = _morton(coords)
    out = torch.zeros(cascades, G ** 3, device=device)
    for c in range(cascades):
        s = min(2 ** (c - 1), scale)
        cc = coords if coarse is None else (coords // (G // coarse)) * (G // coarse) + (G // coarse - 1) / 2
        xyz = (cc.float() / (G - 1) * 2 - 1) * (s - s / G)
        out[c, idx] = inside(xyz, scale).float() * occupied_value
    return out


def bitfield_from_grid(grid, thr=0.0):
    bits = (grid.reshape(-1, 8) > thr).to(torch.uint8)
    w = (2 ** torch.arange(8, device=grid.device)).to(torch.uint8)
    return (bits * w).sum(1).to(torch.uint8)


def intrinsics(W, H, fx=None):
    fx = 1111.1 * (W / 800.0) if fx is None else fx
    K = torch.tensor([[fx, 0, W / 2], [0, fx, H / 2], [0, 0, 1]], dtype=torch.float32)
    return K


def directions(W, H, K):
    v, u = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    d = torch.stack([(u - K[0, 2] + 0.5) / K[0, 0], (v - K[1, 2] + 0.5) / K[1, 1], torch.ones_like(u)], -1)
    return d.reshape(-1, 3)


def hemisphere_poses(n, radius=4.0, dataset_scale=1.15, seed=0):
    """(n,3,4) c2w poses looking at the origin from the upper hemisphere ([right down front] camera frame).
    Positions are divided by 2*dataset_scale like ngp_pl/datasets/nsvf.py:86-87, which puts a radius-4 Lego rig
    at ~1.74 from the centre of a scene bounded by [-0.5,0.5]^3."""
    g = torch.Generator().manual_seed(seed)
    phi = torch.rand(n, generator=g) * 2 * math.pi
    theta = torch.rand(n, generator=g) * (0.45 * math.pi) + 0.05 * math.pi   # elevation above the horizon
    pos = radius * torch.stack([torch.cos(theta) * torch.cos(phi), torch.cos(theta) * torch.sin(phi), torch.sin(theta)], -1)
    fwd = -pos / pos.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
    right = torch.cross(fwd, up, dim=-1); right = right / right.norm(dim=-1, keepdim=True)
    down = torch.cross(fwd, right, dim=-1)
    return torch.stack([right, down, fwd, pos / (2 * dataset_scale)], -1).contiguous()   # (n,3,4)


def get_rays(dirs_cam, c2w):
    """ngp_pl/datasets/ray_utils.py:152-175 as plain torch (the per-step bmm)."""
    if c2w.ndim == 2:
        rays_d = dirs_cam @ c2w[:, :3].T
    else:
        rays_d = torch.bmm(dirs_cam[:, None], c2w[..., :3].transpose(1, 2))[:, 0]
    rays_o = c2w[..., 3].expand_as(rays_d)
    return rays_o.contiguous(), rays_d.contiguous()


def _shade_prims(o, d, spheres, boxes, bg, two_sided=False):
    """Nearest-hit Lambert + ambient shading of spheres [(centre, radius, albedo)] and axis-aligned boxes [(centre, half
    size, albedo)] -> colour (N,3), ray parameter t of the hit (inf = background)."""
    dev = o.device
    N = o.shape[0]
    best_t = torch.full((N,), float("inf"), device=dev)
    colour = torch.full((N, 3), float(bg), device=dev)
    light = torch.tensor([0.4, 0.3, 0.85], device=dev); light = light / light.norm()
    a = (d * d).sum(-1)
    for c, r, alb in spheres:
        oc = o - torch.tensor(c, device=dev)
        hb = (oc * d).sum(-1); cc = (oc * oc).sum(-1) - r * r
        disc = hb * hb - a * cc
        t = (-hb - disc.clamp(min=0).sqrt()) / a
        hit = (disc >= 0) & (t > 0) & (t < best_t)
        n = (oc + t[:, None] * d) / r
        lam = (n * light).sum(-1).clamp(min=0) * 0.7 + 0.3
        colour = torch.where(hit[:, None], torch.tensor(alb, device=dev)[None] * lam[:, None], colour)
        best_t = torch.where(hit, t, best_t)
    for c, h, alb in boxes:
        c = torch.tensor(c, device=dev); h = torch.tensor(h, device=dev)
        inv = 1.0 / d
        lo, hi = (c - h - o) * inv, (c + h - o) * inv
        tmin, tmax = torch.minimum(lo, hi), torch.maximum(lo, hi)
        t1, axis = tmin.max(-1); t2 = tmax.min(-1)[0]
        hit = (t1 <= t2) & (t1 > 0) & (t1 < best_t)
        n = torch.zeros(N, 3, device=dev); n.scatter_(1, axis[:, None], -torch.sign(d.gather(1, axis[:, None])))
        lam = (n * light).sum(-1)
        lam = (lam.abs() if two_sided else lam.clamp(min=0)) * 0.7 + 0.3
        colour = torch.where(hit[:, None], torch.tensor(alb, device=dev)[None] * lam[:, None], colour)
        best_t = torch.where(hit, t1, best_t)
    return colour, best_t


def shade(rays_o, rays_d, scale, bg=1.0):
    """Closed-form ground-truth colour of the analytic scene (nearest primitive, Lambert + ambient)."""
    o = rays_o / scale; d = rays_d / scale                           # primitives are in units of scale
    return _shade_prims(o, d, _SPHERES, [_BOX], bg)[0]


# ---- two more analytic scenes in WORLD units, for the ScanNet-shaped (C4) and the unbounded (C5) workloads ------------
_GREY, _WOOD = (0.75, 0.75, 0.72), (0.55, 0.40, 0.25)
# a room filling the scale-0.5 box: six thin wall slabs, a table, a cabinet and two spheres; cameras stand inside
ROOM = dict(
    spheres=[((0.10, -0.05, -0.20), 0.09, (0.85, 0.25, 0.20)), ((-0.22, 0.18, -0.05), 0.07, (0.20, 0.45, 0.85))],
    boxes=[((0.0, 0.0, -0.44), (0.46, 0.46, 0.02), _WOOD), ((0.0, 0.0, 0.44), (0.46, 0.46, 0.02), _GREY),
           ((-0.44, 0.0, 0.0), (0.02, 0.46, 0.46), (0.80, 0.78, 0.70)), ((0.44, 0.0, 0.0), (0.02, 0.46, 0.46), (0.70, 0.78, 0.80)),
           ((0.0, -0.44, 0.0), (0.46, 0.02, 0.46), (0.78, 0.70, 0.78)), ((0.0, 0.44, 0.0), (0.46, 0.02, 0.46), (0.72, 0.80, 0.72)),
           ((0.10, -0.05, -0.33), (0.16, 0.10, 0.04), _WOOD), ((-0.30, -0.30, -0.22), (0.08, 0.08, 0.20), (0.35, 0.30, 0.28))],
    bg=0.0)
# an unbounded 360-degree scene for scale 16: the small object scene at the centre, a ground slab out to the box and a
# few large far objects, so that all six occupancy cascades hold something; cameras circle the centre
UNBOUNDED = dict(
    spheres=[(tuple(0.5 * v for v in c), 0.5 * r, alb) for c, r, alb in _SPHERES] +
            [((5.0, 2.0, 1.2), 1.5, (0.8, 0.6, 0.2)), ((-6.0, -4.0, 1.8), 2.0, (0.3, 0.7, 0.7)),
             ((2.0, -9.0, 2.4), 2.5, (0.7, 0.3, 0.6)), ((-3.0, 11.0, 3.0), 3.0, (0.5, 0.5, 0.8))],
    boxes=[(tuple(0.5 * v for v in _BOX[0]), tuple(0.5 * v for v in _BOX[1]), _BOX[2]),
           ((0.0, 0.0, -0.40), (15.5, 15.5, 0.04), (0.45, 0.50, 0.40))],
    bg=0.0)


def scene_inside(xyz, sc):
    occ = torch.zeros(xyz.shape[0], dtype=torch.bool, device=xyz.device)
    for c, r, _ in sc["spheres"]:
        occ |= ((xyz - torch.tensor(c, device=xyz.device)) ** 2).sum(-1) <= r * r
    for c, h, _ in sc["boxes"]:
        occ |= ((xyz - torch.tensor(c, device=xyz.device)).abs() <= torch.tensor(h, device=xyz.device)).all(-1)
    return occ


def scene_shade(rays_o, rays_d, sc):
    """-> colour (N,3), ray parameter of the first hit (inf = nothing): the ground truth and the source of depth priors."""
    return _shade_prims(rays_o, rays_d, sc["spheres"], sc["boxes"], sc["bg"], two_sided=True)


def look_at_poses(positions, targets):
    """(n,3) camera positions and look-at points -> (n,3,4) c2w in the [right down front] camera frame."""
    fwd = targets - positions; fwd = fwd / fwd.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
    right = torch.cross(fwd, up, dim=-1); right = right / right.norm(dim=-1, keepdim=True)
    down = torch.cross(fwd, right, dim=-1)
    return torch.stack([right, down, fwd, positions], -1).contiguous()


def room_poses(n, seed=0):
    """ScanNet-shaped sparse views: cameras inside the room at standing height, looking across it."""
    g = torch.Generator().manual_seed(seed)
    pos = torch.stack([(torch.rand(n, generator=g) - 0.5) * 0.5, (torch.rand(n, generator=g) - 0.5) * 0.5,
                       (torch.rand(n, generator=g) - 0.5) * 0.2 + 0.05], -1)
    tgt = torch.stack([(torch.rand(n, generator=g) - 0.5) * 0.7, (torch.rand(n, generator=g) - 0.5) * 0.7,
                       (torch.rand(n, generator=g) - 0.5) * 0.3 - 0.15], -1)
    tgt = torch.where(((tgt - pos).norm(dim=-1, keepdim=True) < 0.15), -pos, tgt)
    return look_at_poses(pos, tgt)


def ring_poses(n, radius=1.3, seed=0):
    """360-degree capture: cameras on a ring around the centre object, slightly above the ground, looking inwards."""
    g = torch.Generator().manual_seed(seed)
    phi = torch.rand(n, generator=g) * 2 * math.pi
    r = radius * (0.85 + 0.3 * torch.rand(n, generator=g))
    pos = torch.stack([r * torch.cos(phi), r * torch.sin(phi), 0.15 + 0.5 * torch.rand(n, generator=g)], -1)
    tgt = torch.stack([0.2 * (torch.rand(n, generator=g) - 0.5), 0.2 * (torch.rand(n, generator=g) - 0.5),
                       -0.1 * torch.rand(n, generator=g)], -1)
    return look_at_poses(pos, tgt)


def write_nsvf_dataset(root, n_train=24, n_test=4, res=100, scale=0.5, seed=0):
    """Write the analytic scene as an NSVF-format "Synthetic" dataset that ngp_pl/datasets/nsvf.py reads unchanged:
    ``bbox.txt`` (xyz_min xyz_max), ``intrinsics.txt`` (fx of the 800-pixel image on the first line), ``pose/*.txt``
    (4x4 camera-to-world, [right down front]) and ``rgb/*.png`` with the split prefixes 0_ (train) / 2_ (test)
    (nsvf.py:17-35,66-75).  `root` must contain 'Synthetic' (that is how the loader picks this layout).  Images are
    res x res; load them with ``downsample = res / 800``.  The loader maps positions with (p - shift) / (2 * 1.05 *
    bbox_half), so raw positions are the model-space ones times 2.1 for the unit bbox written here."""
    import os
    from PIL import Image
    if "Synthetic" not in root:
        raise ValueError("the NSVF loader recognises this layout by 'Synthetic' in the path")
    os.makedirs(os.path.join(root, "rgb"), exist_ok=True); os.makedirs(os.path.join(root, "pose"), exist_ok=True)
    np.savetxt(os.path.join(root, "bbox.txt"), np.array([[-1.0, -1.0, -1.0, 1.0, 1.0, 1.0, 0.1]]), fmt="%.6f")
    with open(os.path.join(root, "intrinsics.txt"), "w") as f:
        f.write("1111.1 400.0 400.0 0.\n0. 0. 0.\n0.\n1.\n800 800\n")
    K = intrinsics(res, res)
    dirs = directions(res, res, K)
    poses = hemisphere_poses(n_train + n_test, seed=seed)
    for i in range(n_train + n_test):
        prefix = "0_" if i < n_train else "2_"
        c2w = poses[i]
        rays_o, rays_d = get_rays(dirs, c2w)
        img = shade(rays_o, rays_d, scale).reshape(res, res, 3)
        Image.fromarray((img.clamp(0, 1) * 255 + 0.5).to(torch.uint8).numpy()).save(
            os.path.join(root, "rgb", f"{prefix}{i:04d}.png"))
        raw = torch.eye(4)
        raw[:3, :3] = c2w[:, :3]
        raw[:3, 3] = c2w[:, 3] * 2.1
        np.savetxt(os.path.join(root, "pose", f"{prefix}{i:04d}.txt"), raw.numpy(), fmt="%.8f")
    return poses[:n_train], poses[n_train:]
