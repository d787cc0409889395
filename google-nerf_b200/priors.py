"""Depth priors for the SCADE-style path: the on-disk formats either side of the hot path.

* LeReS predictions: ``AdelaiDepth/LeReS/Minist_Test/tools/test_scannet.py:85`` writes, next to every input image,
  ``<name>-depth_raw.png`` = ``(pred_depth / pred_depth.max() * 60000).astype(uint16)`` -- an affine-invariant RELATIVE
  depth (scale and shift unknown), which is exactly what ``shiftscale_inv_depthloss`` (losses.py:5-23) is invariant to.
* A ScanNet-shaped scene directory as ``ngp_pl/datasets/scannet.py:85-257`` reads it: ``intrinsic_depth.txt`` (4x4, the
  3x3 part is K), ``test_step_<k>/<split>.txt`` (one frame name per line), ``pose/<name>.txt`` (4x4 camera-to-world),
  ``rgb/<name>.jpg``; positions are shifted and divided by ``2 * scale`` (scannet.py:146-147).  Here the same directory
  may also hold ``leres/<name>-depth_raw.png``; the dataset then hands every training batch a ``disp`` entry with the
  prior of its rays, which `NGPTrainer(lambda_depth > 0)` consumes.
"""
import os

import numpy as np
import torch

LERES_SCALE = 60000.0


def write_leres_prior(path, depth):
    """depth (H,W) float > 0 -> 16-bit PNG exactly like test_scannet.py:85."""
    from PIL import Image
    d = np.asarray(depth, dtype=np.float64)
    Image.fromarray((d / d.max() * LERES_SCALE).astype(np.uint16)).save(path)


def read_leres_prior(path, size=None, as_disparity=True, eps=1e-3):
    """``*-depth_raw.png`` -> (H,W) float32 tensor: relative depth in (0, 1], or (default) its reciprocal, the relative
    disparity the SSI loss takes.  Pixels stored as 0 (no prediction) come back as 0 = "no prior".  size = (W, H)
    resamples (nearest) to the training resolution like the images are (scannet.py:245)."""
    from PIL import Image
    im = Image.open(path)
    if size is not None and tuple(im.size) != tuple(size):
        im = im.resize(tuple(size), Image.NEAREST)
    rel = torch.from_numpy(np.asarray(im).astype(np.float32) / LERES_SCALE)
    if not as_disparity:
        return rel
    return torch.where(rel > 0, 1.0 / rel.clamp(min=eps), torch.zeros_like(rel))


class ScanNetShapedDataset(torch.utils.data.Dataset):
    """The reader of ngp_pl/datasets/scannet.py (+ base.py:24-40 sampling) for a ScanNet-shaped directory, with LeReS
    priors when ``leres/`` exists.  Attributes as the training scripts use them: K (3,3), directions (H*W,3),
    poses (N,3,4), rays (N,H*W,3), img_wh, batch_size; training items are {img_idxs, pix_idxs, rgb[, disp]}."""

    def __init__(self, root_dir, split="train", downsample=1.0, test_skip=8, rot_transpose=False, scale_flip=False,
                 shift=(0.0, 0.0, 0.0), scale=1.0, prior_dir="leres"):
        from PIL import Image
        self.root_dir, self.split, self.downsample = root_dir, split, downsample
        K = np.loadtxt(os.path.join(root_dir, "intrinsic_depth.txt"), dtype=np.float32)[:3, :3]
        w, h = int(640 * downsample), int(480 * downsample)
        self.K = torch.FloatTensor(K)
        self.img_wh = (w, h)
        v, u = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        self.directions = torch.stack([(u - K[0, 2] + 0.5) / K[0, 0], (v - K[1, 2] + 0.5) / K[1, 1],
                                       torch.ones_like(u)], -1).reshape(-1, 3)       # ray_utils.py:33-35
        with open(os.path.join(root_dir, f"test_step_{test_skip}", split + ".txt")) as f:
            names = [ln.rstrip() for ln in f if ln.strip()]
        poses, rays, priors = [], [], []
        has_prior = os.path.isdir(os.path.join(root_dir, prior_dir))
        for name in names:
            c2w = np.loadtxt(os.path.join(root_dir, "pose", name + ".txt"), dtype=np.float64).reshape(4, 4)[:3]
            if rot_transpose:
                c2w[:, :3] = c2w[:, :3].T
            if scale_flip:
                c2w[:3, 1] *= -1; c2w[:3, 2] *= -1
            c2w[:, 3] -= np.asarray(shift)
            c2w[:, 3] /= 2 * scale                                                   # scannet.py:146-147
            poses.append(c2w)
            img = Image.open(os.path.join(root_dir, "rgb", name + ".jpg")).convert("RGB").resize((w, h), Image.BILINEAR)
            rays.append(torch.from_numpy(np.asarray(img).astype(np.float32) / 255.0).reshape(-1, 3))
            if has_prior:
                priors.append(read_leres_prior(os.path.join(root_dir, prior_dir, name + "-depth_raw.png"), (w, h)).reshape(-1))
        self.poses = torch.FloatTensor(np.stack(poses))
        self.rays = torch.stack(rays)
        self.priors = torch.stack(priors) if priors else None
        self.batch_size = 8192

    def __len__(self):
        return 1000 if self.split.startswith("train") else len(self.poses)

    def __getitem__(self, idx):
        if self.split.startswith("train"):
            img_idxs = np.random.choice(len(self.poses), self.batch_size)
            pix_idxs = np.random.choice(self.img_wh[0] * self.img_wh[1], self.batch_size)
            sample = {"rgb": self.rays[img_idxs, pix_idxs], "img_idxs": img_idxs, "pix_idxs": pix_idxs}
            if self.priors is not None:
                sample["disp"] = self.priors[img_idxs, pix_idxs]
            return sample
        sample = {"pose": self.poses[idx], "img_idxs": idx}
        if len(self.rays) > 0:
            sample["rgb"] = self.rays[idx]
        if self.priors is not None:
            sample["disp"] = self.priors[idx]
        return sample


def write_scannet_dataset(root, n_views=20, test_skip=8, seed=0):
    """Write the analytic ROOM scene (synthetic.ROOM) as a ScanNet-shaped directory with LeReS-style priors, 640x480
    images (load with downsample = 624/640 for the 624x468 training resolution of BASELINE config 4).  The prior of a
    view is its true depth under that view's own random affine map (what LeReS leaves undetermined)."""
    from PIL import Image
    from . import synthetic as syn
    for d in ("rgb", "pose", "leres", f"test_step_{test_skip}"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    W, H = 640, 480
    K = syn.intrinsics(W, H, fx=577.870605)
    K[0, 2], K[1, 2] = 319.5, 239.5
    K4 = np.eye(4, dtype=np.float32); K4[:3, :3] = K.numpy()
    np.savetxt(os.path.join(root, "intrinsic_depth.txt"), K4, fmt="%.6f")
    dirs = syn.directions(W, H, K)
    poses = syn.room_poses(n_views, seed=seed)
    g = np.random.default_rng(seed)
    names = {"train": [], "test": []}
    for i in range(n_views):
        name = f"{i:06d}"
        names["test" if i % test_skip == test_skip - 1 else "train"].append(name)
        ro, rd = syn.get_rays(dirs, poses[i])
        col, t = syn.scene_shade(ro, rd, syn.ROOM)
        Image.fromarray((col.clamp(0, 1) * 255 + 0.5).to(torch.uint8).reshape(H, W, 3).numpy()).save(
            os.path.join(root, "rgb", name + ".jpg"), quality=95)
        c2w = torch.eye(4); c2w[:3] = poses[i]
        np.savetxt(os.path.join(root, "pose", name + ".txt"), c2w.numpy(), fmt="%.8f", delimiter=" ")
        depth = torch.where(torch.isfinite(t), t, torch.zeros_like(t)).reshape(H, W).numpy()
        a, b = 0.5 + g.random(), 0.2 * g.random()
        write_leres_prior(os.path.join(root, "leres", name + "-depth_raw.png"), a * depth + b * (depth > 0))
    for split, ns in names.items():
        with open(os.path.join(root, f"test_step_{test_skip}", split + ".txt"), "w") as f:
            f.write("\n".join(ns) + "\n")
    return poses
