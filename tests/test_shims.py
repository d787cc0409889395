"""Import shims for the reference's third-party dependencies."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "google-nerf_b200", "shims")


@pytest.fixture()
def shim_path(built_lib):
    sys.path.insert(0, SHIMS)
    yield
    sys.path.remove(SHIMS)
    for m in ("kornia", "torch_scatter", "apex", "apex.optimizers", "vren", "tinycudann"):
        sys.modules.pop(m, None)


def test_kornia_and_torch_scatter_shims(shim_path):
    from kornia import create_meshgrid, create_meshgrid3d
    g = create_meshgrid(3, 4, False)[0]                        # (H,W,2) with u = x, v = y (ray_utils.py:23-24)
    assert g.shape == (3, 4, 2) and g[1, 2].tolist() == [2.0, 1.0]
    g3 = create_meshgrid3d(2, 3, 4, False, dtype=torch.int32).reshape(-1, 3)      # train.py:76-77
    assert g3.shape == (24, 3) and g3[1].tolist() == [1, 0, 0] and g3[-1].tolist() == [3, 2, 1]
    from google_nerf_b200 import synthetic as syn
    assert torch.equal(create_meshgrid3d(8, 8, 8, False, dtype=torch.int32).reshape(-1, 3), syn.grid_coords(8))
    from torch_scatter import segment_csr
    src = torch.arange(12.0).view(6, 2)
    out = segment_csr(src, torch.tensor([0, 2, 2, 6]))
    assert out.tolist() == [[2.0, 4.0], [0.0, 0.0], [28.0, 32.0]]
    import vren, tinycudann                                        # noqa: F401  (resolve to this implementation)
    assert hasattr(vren, "raymarching_train") and hasattr(tinycudann, "NetworkWithInputEncoding")


@pytest.mark.gpu
def test_fused_adam_shim_matches_reference_formula(shim_path):
    from apex.optimizers import FusedAdam
    from oracle.ngp_ref import AdamRef
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(1000, generator=g)
    p = torch.nn.Parameter(p0.clone().cuda()); ref = p0.clone().requires_grad_(True)
    opt = FusedAdam([p], 1e-2, eps=1e-15); oref = AdamRef([ref], lr=1e-2, eps=1e-15)
    for _ in range(3):
        gr = torch.randn(1000, generator=g)
        p.grad = gr.cuda(); ref.grad = gr.clone()
        opt.step(); oref.step()
        torch.testing.assert_close(p.detach().cpu(), ref.detach(), rtol=1e-5, atol=1e-6)
