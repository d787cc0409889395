"""train.py:39,229-238."""
import torch
import torch.distributed as dist


def all_gather_ddp_if_available(tensor, group=None, sync_grads=False):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        parts = [torch.empty_like(tensor) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, tensor.contiguous(), group=group)
        return torch.stack(parts)
    return tensor
