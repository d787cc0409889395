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
    # a renderer that writes pixels and totals straight into the rank's block (what rendering.py::_WholeRays does on the
    # GPU): no packing on the host, totals from the gathered tails; a rank that reports a ray at the sample budget makes
    # EVERY rank render the frame again with whole_rays=False
    calls = []

    def packed_render(ro_, rd_, packed_out=None, tail_out=None, whole_rays=True, cut_on_rank=None, **kw):
        res = _fake_render(ro_, rd_)
        if packed_out is None or whole_rays is False:
            calls.append("generic")
            return res
        m = len(ro_)
        packed_out[:m, 0:3] = res["rgb"]; packed_out[:m, 3] = res["depth"]; packed_out[:m, 4] = res["opacity"]
        cnt = res["total_samples"]
        tail_out.copy_(torch.tensor([cnt & 0xffff, (cnt >> 16) & 0xffff, cnt >> 32, 1.0 if cut_on_rank == rank else 0.0]))
        calls.append("packed")
        return {"rgb": packed_out[:m, 0:3], "depth": packed_out[:m, 3], "opacity": packed_out[:m, 4],
                "total_samples": None, "tail": tail_out}
    for cut, want_calls in ((None, ["packed"]), (1, ["packed", "generic"])):
        calls.clear()
        out = D.render_sharded(packed_render, ro, rd, tile=64, cut_on_rank=cut)
        assert calls == want_calls, (cut, calls)
        for k in ("rgb", "depth", "opacity"):
            torch.testing.assert_close(out[k], full[k], rtol=1e-6, atol=1e-7)
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


def _pl_worker(rank, world, port, out_dir):
    """The Lightning stand-in with devices=2: DDP semantics over gloo (broadcast, own batches, averaged gradients)."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "google-nerf_b200", "shims"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from pytorch_lightning import LightningModule, Trainer
    from pytorch_lightning.callbacks import ModelCheckpoint
    from pytorch_lightning.plugins import DDPPlugin

    class Sys(LightningModule):
        def __init__(self):
            super().__init__()
            torch.manual_seed(100 + rank)                     # replicas start DIFFERENT: the broadcast must fix that
            self.lin = torch.nn.Linear(4, 1)

        def configure_optimizers(self):
            return torch.optim.SGD(self.parameters(), lr=0.1)

        def train_dataloader(self):
            g = torch.Generator().manual_seed(rank)           # every rank draws its own batches
            return [torch.randn(8, 4, generator=g) for _ in range(5)]

        def training_step(self, batch, batch_idx):
            return (self.lin(batch) - batch.sum(-1, keepdim=True)).pow(2).mean()

    m = Sys()
    ck = ModelCheckpoint(dirpath=out_dir, filename="{epoch:d}")
    Trainer(max_epochs=1, devices=world, strategy=DDPPlugin(find_unused_parameters=False), callbacks=[ck]).fit(m)
    # reference: the same thing by hand on rank 0's initial weights, gradient = mean over both ranks' batches
    torch.manual_seed(100)
    ref = torch.nn.Linear(4, 1)
    data = [[torch.randn(8, 4, generator=torch.Generator().manual_seed(r)) for _ in range(1)] for r in range(world)]
    gens = [torch.Generator().manual_seed(r) for r in range(world)]
    for _ in range(5):
        ref.zero_grad()
        for r in range(world):
            b = torch.randn(8, 4, generator=gens[r])
            ((ref(b) - b.sum(-1, keepdim=True)).pow(2).mean() / world).backward()
        with torch.no_grad():
            for p in ref.parameters():
                p -= 0.1 * p.grad
    torch.testing.assert_close(m.lin.weight.detach(), ref.weight.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(m.lin.bias.detach(), ref.bias.detach(), rtol=1e-5, atol=1e-6)
    assert os.path.exists(os.path.join(out_dir, "epoch=0.ckpt")) or rank != 0
    dist.barrier()
    if rank == 1:
        assert len(os.listdir(out_dir)) == 1                  # only rank 0 wrote a checkpoint
    dist.destroy_process_group()


def test_lightning_stand_in_ddp_world_2_gloo(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_pl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
