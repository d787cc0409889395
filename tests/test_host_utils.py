"""losses / metrics / checkpoint helpers (CPU): behaviour of ngp_pl/losses.py, metrics.py, utils.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_ssi_depth_loss_is_shift_scale_invariant():
    from google_nerf_b200.losses import shiftscale_inv_depthloss
    g = torch.Generator().manual_seed(0)
    d = torch.rand(1001, generator=g) + 0.1
    # an affine map of the prediction (positive scale) leaves the loss at zero
    assert shiftscale_inv_depthloss(3.0 * d + 0.7, d).abs().max() < 1e-9 + 1e-5
    a = shiftscale_inv_depthloss(d, d.flip(0))
    b = shiftscale_inv_depthloss(5.0 * d - 2.0, 0.5 * d.flip(0) + 4.0)
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    # explicit formula (losses.py:15-23) on a tiny case: median of an odd-length vector is its middle element
    p, q = torch.tensor([1.0, 2.0, 6.0]), torch.tensor([0.0, 1.0, 2.0])
    pn = (p - 2.0) / ((p - 2.0).abs().mean()); qn = (q - 1.0) / ((q - 1.0).abs().mean())
    torch.testing.assert_close(shiftscale_inv_depthloss(p, q), (pn - qn) ** 2)


def test_nerf_loss_matches_oracle_and_fused_terms():
    from google_nerf_b200.losses import NeRFLoss
    from oracle import ngp_ref
    g = torch.Generator().manual_seed(1)
    res = {"rgb": torch.rand(64, 3, generator=g), "opacity": torch.rand(64, generator=g)}
    tgt = {"rgb": torch.rand(64, 3, generator=g)}
    d = NeRFLoss()(res, tgt)
    assert set(d) == {"rgb", "opacity"} and d["rgb"].shape == (64, 3) and d["opacity"].shape == (64,)
    total = sum(v.mean() for v in d.values())                 # train.py:160
    torch.testing.assert_close(total, ngp_ref.nerf_loss(res, tgt["rgb"]))


def test_psnr():
    from google_nerf_b200.metrics import mse, psnr
    a, b = torch.zeros(4, 3), torch.full((4, 3), 0.1)
    assert abs(float(mse(a, b)) - 0.01) < 1e-8 and abs(float(psnr(a, b)) - 20.0) < 1e-4
    m = torch.tensor([True, False, True, False])
    assert mse(a, b, m, reduction="none").shape == (2, 3)


def test_checkpoint_helpers(tmp_path):
    from google_nerf_b200.utils import extract_model_state_dict, load_ckpt, slim_ckpt
    lin = torch.nn.Linear(3, 2)
    sd = {"model." + k: v.clone() + 1 for k, v in lin.state_dict().items()}
    sd.update({"directions": torch.zeros(4, 3), "poses": torch.zeros(2, 3, 4), "model.density_grid": torch.zeros(8),
               "model.grid_coords": torch.zeros(8, 3), "val_lpips.net.w": torch.zeros(1), "model.skip.me": torch.ones(1)})
    path = os.path.join(tmp_path, "c.ckpt")
    torch.save({"state_dict": sd, "epoch": 3}, path)
    got = extract_model_state_dict(path, prefixes_to_ignore=["skip", "density_grid", "grid_coords"])
    assert set(got) == {"weight", "bias"}
    before = lin.weight.detach().clone()
    load_ckpt(lin, path, prefixes_to_ignore=["skip", "density_grid", "grid_coords"])
    torch.testing.assert_close(lin.weight.detach(), before + 1)
    slim = slim_ckpt(path)
    assert set(slim) == {"model.weight", "model.bias", "model.skip.me"}
    assert "poses" in slim_ckpt(path, save_poses=True)
    load_ckpt(lin, "")                                        # empty path is a no-op


def test_segment_csr_shim_matches_per_ray_sums():
    """The torch_scatter.segment_csr stand-in over the reference's CSR pointer (custom_functions.py:108-111) against a
    per-ray loop.  (RayMarcher.backward itself is one CUDA launch: tests/test_gpu_baseline.py.)"""
    sys.path.insert(0, os.path.join(ROOT, "google-nerf_b200", "shims"))
    from torch_scatter import segment_csr
    g = torch.Generator().manual_seed(2)
    counts = torch.tensor([3, 0, 5, 1, 0, 7])
    n = int(counts.sum())
    start = torch.cumsum(counts, 0) - counts
    rays_a = torch.stack([torch.arange(6), start, counts], 1)
    d_xyz = torch.randn(n, 3, generator=g)
    want_o = torch.stack([d_xyz[s:s + c].sum(0) for s, c in zip(start.tolist(), counts.tolist())])
    ptr = torch.cat([rays_a[:, 1], rays_a[-1:, 1] + rays_a[-1:, 2]])           # the reference's CSR pointer
    torch.testing.assert_close(segment_csr(d_xyz, ptr), want_o)


def test_leres_prior_and_scannet_shaped_dataset(tmp_path):
    """f3: the LeReS `*-depth_raw.png` encoding (test_scannet.py:85) round-trips, and a ScanNet-shaped directory written
    from the analytic room is read back with the reference dataset's attributes, normalisation and batch format."""
    import numpy as np
    from google_nerf_b200 import priors, synthetic as syn
    depth = np.random.default_rng(0).random((48, 64)) * 3 + 0.5
    depth[0, :5] = 0.0
    p = str(tmp_path / "a-depth_raw.png")
    priors.write_leres_prior(p, depth)
    rel = priors.read_leres_prior(p, as_disparity=False)
    assert rel.shape == (48, 64) and abs(float(rel.max()) - 1.0) < 1e-4
    np.testing.assert_allclose(rel.numpy(), depth / depth.max(), atol=1.0 / 60000 + 1e-7)
    disp = priors.read_leres_prior(p)
    assert float(disp[0, 0]) == 0.0 and abs(float(disp[5, 5]) - depth.max() / depth[5, 5]) < 2e-3 * depth.max() / depth[5, 5]
    root = str(tmp_path / "scene0000_00")
    poses = priors.write_scannet_dataset(root, n_views=9, test_skip=4)
    ds = priors.ScanNetShapedDataset(root, "train", downsample=624 / 640, test_skip=4, scale=1.0)
    assert ds.img_wh == (624, 468) and ds.K.shape == (3, 3) and ds.directions.shape == (624 * 468, 3)
    assert ds.poses.shape == (7, 3, 4) and ds.rays.shape == (7, 624 * 468, 3) and ds.priors.shape == (7, 624 * 468)
    torch.testing.assert_close(ds.poses[0, :, 3], poses[0, :, 3] / 2.0, rtol=1e-5, atol=1e-6)   # (x - shift) / (2 scale)
    ds.batch_size = 512
    b = ds[0]
    assert set(b) == {"rgb", "img_idxs", "pix_idxs", "disp"} and b["rgb"].shape == (512, 3) and b["disp"].shape == (512,)
    assert len(ds) == 1000 and float(b["disp"].min()) >= 0.0
    # the prior of a view is an affine image of its true depth: the SSI loss against the true disparity is ~0
    from google_nerf_b200.losses import shiftscale_inv_depthloss
    test = priors.ScanNetShapedDataset(root, "test", downsample=1.0, test_skip=4)
    ro, rd = syn.get_rays(syn.directions(640, 480, test.K), poses[3])
    t = syn.scene_shade(ro, rd, syn.ROOM)[1]
    rel_depth = 1.0 / test.priors[0]                                  # back from disparity
    assert float(shiftscale_inv_depthloss(rel_depth, t).mean()) < 1e-3


def test_param_layout_hook(tmp_path):
    from google_nerf_b200 import utils
    lin = torch.nn.Linear(2, 2)
    lin.params = torch.nn.Parameter(torch.arange(6.0))
    path = str(tmp_path / "c.ckpt")
    torch.save({"state_dict": {"model.params": torch.arange(6.0).flip(0), "model.weight": lin.weight.detach(),
                               "model.bias": lin.bias.detach()}}, path)
    utils.register_param_layout("flipped", lambda k, p: p.flip(0), lambda k, p: p.flip(0))
    utils.load_ckpt(lin, path, param_layout="flipped")
    assert torch.equal(lin.params.detach(), torch.arange(6.0))
    utils.load_ckpt(lin, path, param_layout="tcnn")                   # identity
    assert torch.equal(lin.params.detach(), torch.arange(6.0).flip(0))
