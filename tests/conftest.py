import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    import __graft_entry__ as g
    g.build()
    from google_nerf_b200 import _lib
    return _lib


def make_scene(scale=0.5, n_rays=2048, seed=0, W=800, H=800, radius=4.0):
    """Synthetic analytic scene + Lego-shaped rays on the CPU (shared by oracle and CUDA tests)."""
    from google_nerf_b200 import synthetic as syn
    import numpy as np
    g = torch.Generator().manual_seed(seed)
    cascades = max(1 + int(np.ceil(np.log2(2 * scale))), 1)
    grid = syn.density_grid(scale, cascades)
    bitfield = syn.bitfield_from_grid(grid)
    K = syn.intrinsics(W, H)
    dirs = syn.directions(W, H, K)
    poses = syn.hemisphere_poses(16, radius=radius * max(1.0, scale), seed=seed)
    ii = torch.randint(poses.shape[0], (n_rays,), generator=g)
    pi = torch.randint(W * H, (n_rays,), generator=g)
    rays_o, rays_d = syn.get_rays(dirs[pi], poses[ii])
    noise = torch.rand(n_rays, generator=g)
    return dict(scale=scale, cascades=cascades, grid=grid, bitfield=bitfield, rays_o=rays_o, rays_d=rays_d,
                noise=noise, center=torch.zeros(1, 3), half_size=torch.ones(1, 3) * scale, K=K, poses=poses)


@pytest.fixture(scope="session")
def scene05():
    return make_scene(0.5, 2048, 0)


@pytest.fixture(scope="session")
def scene4():
    return make_scene(4.0, 1024, 1)
