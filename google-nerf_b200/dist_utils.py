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


_PARTITIONS = {}


def _partition(n, world, rank, tile, device):
    """Index bookkeeping of a sharded render, cached per shape: which rays this rank renders, how many every rank
    renders, and the permutation that puts the gathered (rank-major, padded) rows back in ray order."""
    key = (n, world, rank, tile, str(device))
    part = _PARTITIONS.get(key)
    if part is None:
        idx = torch.arange(n, device=device)
        if tile is None:
            bounds = [shard_bounds(n, world, r) for r in range(world)]
            owners = torch.empty(n, dtype=torch.int64, device=device)
            for r, (a, b) in enumerate(bounds):
                owners[a:b] = r
        else:
            owners = (idx // int(tile)) % world
        sels = [(owners == r).nonzero(as_tuple=True)[0] for r in range(world)]
        counts = [int(x.numel()) for x in sels]
        longest = max(counts)
        # row j of rank r's padded part (longest rows + one row of bookkeeping) lands at ray sels[r][j]
        src = torch.cat([r * (longest + 1) + torch.arange(c, device=device) for r, c in enumerate(counts)])
        dst = torch.cat(sels)
        # the same as ONE gather: ray i of the frame (and, behind the rays, every rank's bookkeeping row) <- row inv[i]
        inv = torch.empty(n + world, dtype=torch.int64, device=device)
        inv[dst] = src
        inv[n:] = torch.arange(world, device=device) * (longest + 1) + longest
        part = dict(sel=sels[rank], counts=counts, longest=longest, src=src, dst=dst, inv=inv)
        if len(_PARTITIONS) > 16:
            _PARTITIONS.clear()
        _PARTITIONS[key] = part
    return part


def _takes_kwargs(fn):
    import inspect
    try:
        return any(p.kind == p.VAR_KEYWORD for p in inspect.signature(fn).parameters.values())
    except (TypeError, ValueError):
        return False


def render_sharded(render_fn, rays_o, rays_d, group=None, tile=None, **kwargs):
    """Tile-sharded test-time render; the (rgb, depth, opacity) parts are gathered with ONE collective and put back in
    ray order.  `render_fn(rays_o, rays_d, **kwargs)` -> dict with those three keys (+ total_samples).

    tile=None: rank r renders the contiguous band [start_r, end_r).  tile=T: rays are cut into tiles of T consecutive
    rays dealt round-robin to the ranks (tile t -> rank t % world) -- empty-space rays are cheap, so contiguous bands of
    an image are badly balanced; T = one image row keeps each rank's rays coherent.  The index bookkeeping
    is cached per (n_rays, world, tile), so repeated frames cost no host synchronisation here."""
    rank, world = world_info(group)
    n = rays_o.shape[0]
    if world == 1:
        return render_fn(rays_o, rays_d, **kwargs)
    part = _partition(n, world, rank, tile, rays_o.device)
    sel = part["sel"]
    # one (longest + 1, 5) fp32 block per rank: rgb | depth | opacity, and in the last row the rank's totals (samples
    # marched as three exact 16-bit digits, rays cut at the sample budget) -- ONE collective and one host sync per frame
    L1 = part["longest"] + 1
    c = sel.numel()
    mine = part.get("mine")
    if mine is None:
        mine = part["mine"] = torch.zeros(L1, 5, dtype=torch.float32, device=rays_o.device)
    # a renderer that can write its pixels and totals straight into the block does (models/rendering.py::_WholeRays)
    extra = dict(packed_out=mine, tail_out=mine[L1 - 1, :4]) if _takes_kwargs(render_fn) else {}
    res = render_fn(rays_o.index_select(0, sel), rays_d.index_select(0, sel), **extra, **kwargs)
    if res.get("tail") is None:
        mine[:c, 0:3] = res["rgb"]; mine[:c, 3] = res["depth"]; mine[:c, 4] = res["opacity"]
        cnt = int(res["total_samples"])
        mine[L1 - 1].copy_(torch.tensor([cnt & 0xffff, (cnt >> 16) & 0xffff, cnt >> 32, 0, 0], dtype=torch.float32))
    gathered = torch.empty(world * L1, 5, dtype=torch.float32, device=rays_o.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    full = gathered.index_select(0, part["inv"])             # ray order, then the ranks' bookkeeping rows
    digits = full[n:, :4].cpu().to(torch.int64).sum(0)       # the frame's one host synchronisation
    full = full[:n]
    if int(digits[3]) > 0 and kwargs.get("whole_rays", True) is not False:
        # some rank had a ray at the per-call sample budget: every rank sees that and renders the frame with the round loop
        return render_sharded(render_fn, rays_o, rays_d, group=group, tile=tile, **{**kwargs, "whole_rays": False})
    total = int(digits[0]) + (int(digits[1]) << 16) + (int(digits[2]) << 32)
    return {"rgb": full[:, 0:3], "depth": full[:, 3], "opacity": full[:, 4], "total_samples": total}
