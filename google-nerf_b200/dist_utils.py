"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on the GPUs, gloo in the CPU tests).

The path shards only where it is naturally parallel:
  training   -- every rank draws its own ray batch (PL DDP semantics, ngp_pl/train.py:262-263); the two flat
                gradient buffers are all-reduced (sum) and averaged, the density grid is max-reduced;
  inference  -- the image's rays are split into contiguous row bands, no communication until the final gather.
"""
import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n, world, rank):
    """Contiguous, balanced split of n items: the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_mean_(tensors, group=None):
    """In-place average of gradient buffers over ranks (what DDP does implicitly)."""
    _, world = world_info(group)
    if world == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(world)
    return tensors


def grid_max_reduce_(density_grid, group=None):
    """All ranks end with the element-wise max of their density grids (new vs the reference, SURVEY.md F8:
    PL DDP has no explicit occupancy sync).  Cells marked invisible (-1) stay -1 because every rank marks the
    same cells."""
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(density_grid, op=dist.ReduceOp.MAX, group=group)
    return density_grid


def broadcast_model_(model, src=0, group=None):
    _, world = world_info(group)
    if world > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src, group=group)
    return model


def render_sharded(render_fn, rays_o, rays_d, group=None, tile=None, **kwargs):
    """Tile-sharded test-time render; the (rgb, depth, opacity) parts are all-gathered and put back in ray order.
    `render_fn(rays_o, rays_d, **kwargs)` -> dict with those three keys (+ total_samples).

    tile=None: rank r renders the contiguous band [start_r, end_r).  tile=T: rays are cut into tiles of T consecutive
    rays dealt round-robin to the ranks (tile t -> rank t % world) -- empty-space rays are cheap, so contiguous bands of
    an image are badly balanced; T = one image row keeps each rank's rays coherent."""
    rank, world = world_info(group)
    n = rays_o.shape[0]
    if tile is None:
        owners = None
        s, e = shard_bounds(n, world, rank)
        sel = slice(s, e)
        counts = [b - a for a, b in (shard_bounds(n, world, r) for r in range(world))]
    else:
        owners = (torch.arange(n, device=rays_o.device) // int(tile)) % world
        sel = (owners == rank).nonzero(as_tuple=True)[0]
        counts = [int(c) for c in torch.bincount(owners, minlength=world).tolist()]
    res = render_fn(rays_o[sel].contiguous(), rays_d[sel].contiguous(), **kwargs)
    if world == 1:
        return res
    out = {}
    longest = max(counts)
    for k in ("rgb", "depth", "opacity"):
        v = res[k].contiguous()
        if v.shape[0] < longest:                              # all_gather needs equal shapes: pad the short parts
            v = torch.cat([v, v.new_zeros((longest - v.shape[0],) + tuple(v.shape[1:]))], 0)
        parts = [torch.empty_like(v) for _ in counts]
        dist.all_gather(parts, v, group=group)
        if owners is None:
            out[k] = torch.cat([p[:c] for p, c in zip(parts, counts)], 0)
        else:
            full = v.new_empty((n,) + tuple(v.shape[1:]))
            for r, (p, c) in enumerate(zip(parts, counts)):
                full[owners == r] = p[:c]
            out[k] = full
    ts = torch.as_tensor(res["total_samples"], device=out["rgb"].device, dtype=torch.int64).reshape(1).clone()
    dist.all_reduce(ts, group=group)
    out["total_samples"] = int(ts.item())
    return out
