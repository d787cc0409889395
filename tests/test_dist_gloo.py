"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: ray-band sharding + gather == unsharded render,
gradient averaging, occupancy-grid max-reduce."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_render(rays_o, rays_d, **kw):
    # deterministic per-ray function standing in for the renderer (pixels are independent)
    v = (rays_o * 3 + rays_d).sum(-1)
    return {"rgb": torch.stack([v, v * 2, v * 3], -1), "depth": v.abs(), "opacity": torch.sigmoid(v),
            "total_samples": int(rays_o.shape[0]) * 7}


def _worker(rank, world, port, n):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from google_nerf_b200 import dist_utils as D
    g = torch.Generator().manual_seed(0)
    ro, rd = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    full = _fake_render(ro, rd)
    out = D.render_sharded(_fake_render, ro, rd)
    for k in ("rgb", "depth", "opacity"):
        assert torch.equal(out[k], full[k]), k
    assert out["total_samples"] == full["total_samples"]
    # round-robin tiles (load-balanced partition): same pixels back in ray order, ragged last tile included
    for tile in (1, 7, 64, 5000):
        out = D.render_sharded(_fake_render, ro, rd, tile=tile)
        for k in ("rgb", "depth", "opacity"):
            # (sigmoid's vectorised and tail code paths differ in the last ulp, and the subsets change which is used)
            torch.testing.assert_close(out[k], full[k], rtol=1e-6, atol=1e-7, msg=lambda m: f"{k} tile {tile}: {m}")
        assert out["total_samples"] == full["total_samples"]
    # bands cover the rays exactly once
    bounds = [D.shard_bounds(n, world, r) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
    # gradient averaging == mean over ranks; grid max-reduce == elementwise max, -1 cells stay -1
    gr = torch.full((5,), float(rank + 1)); D.allreduce_mean_([gr])
    assert torch.allclose(gr, torch.full((5,), (1 + world) / 2))
    grid = torch.tensor([-1.0, 0.5 * (rank + 1), 3.0 - rank]); D.grid_max_reduce_(grid)
    assert torch.equal(grid, torch.tensor([-1.0, 0.5 * world, 3.0]))
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin.weight.fill_(float(rank))
    D.broadcast_model_(lin, src=0)
    assert float(lin.weight.abs().max()) == 0.0
    dist.destroy_process_group()


def test_world_size_2_gloo():
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, 1001), nprocs=2, join=True)


def test_shard_bounds_single_process():
    sys.path.insert(0, ROOT)
    from google_nerf_b200 import dist_utils as D
    assert D.shard_bounds(10, 3, 0) == (0, 4) and D.shard_bounds(10, 3, 1) == (4, 7) and D.shard_bounds(10, 3, 2) == (7, 10)
    out = D.render_sharded(_fake_render, torch.ones(4, 3), torch.ones(4, 3))
    assert out["rgb"].shape == (4, 3)
