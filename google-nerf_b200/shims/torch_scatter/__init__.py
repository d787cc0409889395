"""torch_scatter.segment_csr(src, indptr) with reduce='sum' (ngp_pl/models/custom_functions.py:109-111)."""
import torch


def segment_csr(src, indptr, out=None, reduce="sum"):
    if reduce not in ("sum", "add"):
        raise NotImplementedError("only reduce='sum' is used by ngp_pl")
    n_seg = indptr.numel() - 1
    counts = (indptr[1:] - indptr[:-1]).to(torch.int64)
    seg = torch.repeat_interleave(torch.arange(n_seg, device=src.device), counts)
    res = torch.zeros((n_seg,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    lo, hi = int(indptr[0]), int(indptr[-1])
    res.index_add_(0, seg, src[lo:hi])
    if out is not None:
        out.copy_(res); return out
    return res
