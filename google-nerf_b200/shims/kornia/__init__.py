"""The two kornia helpers ngp_pl uses: create_meshgrid (datasets/ray_utils.py:23) and create_meshgrid3d
(train.py:76-77), with kornia 0.6.5 semantics for normalized_coordinates=False."""
import torch


def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    gx, gy = torch.meshgrid(xs, ys, indexing="ij")                       # (W,H)
    return torch.stack([gx, gy], -1).permute(1, 0, 2).unsqueeze(0)        # (1,H,W,2), last dim = (x, y)


def create_meshgrid3d(depth, height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    zs = torch.linspace(0, depth - 1, depth, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
        zs = (zs / (depth - 1) - 0.5) * 2
    gz, gy, gx = torch.meshgrid(zs, ys, xs, indexing="ij")               # (D,H,W)
    return torch.stack([gx, gy, gz], -1).unsqueeze(0)                     # (1,D,H,W,3), last dim = (x, y, z)
