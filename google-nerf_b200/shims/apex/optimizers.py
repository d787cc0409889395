"""apex.optimizers.FusedAdam as used at ngp_pl/train.py:112 (`FusedAdam(net_params, lr, eps=1e-15)`): Adam with bias
correction and no weight decay, one fused b2n_adam_step launch per parameter tensor.  Parameters must be CUDA fp32
tensors (apex's own requirement); there is no CPU path."""
import torch

from google_nerf_b200 import _lib as L


def _bump_version(t):
    try:
        torch._C._increment_version([t])
    except TypeError:
        torch._C._increment_version(t)
    except AttributeError:                                   # very old torch: an in-place no-op does the same
        t.add_(0)


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, adam_w_mode=True,
                 weight_decay=0.0, amsgrad=False, set_grad_none=True):
        if amsgrad or weight_decay != 0.0 or not bias_correction:
            raise NotImplementedError("only the configuration ngp_pl uses is built (bias-corrected Adam, no decay)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.set_grad_none = set_grad_none

    def zero_grad(self, set_to_none=None):
        super().zero_grad(set_to_none=self.set_grad_none if set_to_none is None else set_to_none)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None or p.numel() == 0:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                    st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
                st["step"] += 1
                g = p.grad
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdam: parameters must be contiguous CUDA fp32 tensors (no CPU fallback)")
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.to(torch.float32).contiguous()
                n4 = p.numel() // 4 * 4
                if n4:
                    L.call("b2n_adam_step", L.ptr(p), L.ptr(g), L.ptr(st["exp_avg"]), L.ptr(st["exp_avg_sq"]), None, n4,
                           float(group["lr"]), float(b1), float(b2), float(group["eps"]), 1.0, st["step"], None, 0)
                if n4 < p.numel():                        # a tail shorter than one 16-byte vector: same formula, torch ops on the GPU
                    sl = slice(n4, p.numel())
                    pf, gf, m, v = p.view(-1)[sl], g.view(-1)[sl], st["exp_avg"].view(-1)[sl], st["exp_avg_sq"].view(-1)[sl]
                    m.mul_(b1).add_(gf, alpha=1 - b1); v.mul_(b2).addcmul_(gf, gf, value=1 - b2)
                    c1, c2 = 1 - b1 ** st["step"], 1 - b2 ** st["step"]
                    pf.sub_(group["lr"] * (m / c1) / ((v / c2).sqrt() + group["eps"]))
                _bump_version(p)          # the kernel wrote through the raw pointer: consumers that cache derived data
                                          # by (pointer, version) -- the tinycudann modules' fp16 copy -- must see a change
        return loss
