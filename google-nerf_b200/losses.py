"""Loss terms of the training scripts (ngp_pl/losses.py), same names and return conventions.

`NeRFLoss` returns per-element tensors keyed 'rgb' / 'opacity'; the training step sums their means
(ngp_pl/train.py:159-160).  `NGPTrainer` evaluates the same two terms and their gradients inside the compositing
kernel (csrc/composite.cu: composite_loss_fwbw_kernel); this module is the autograd form for code written against
the reference API (render() + loss + backward), and what the tests compare the fused kernel with.
"""
import torch
from torch import nn


def shiftscale_inv_depthloss(disp_pred, disp_gt):
    """Shift- and scale-invariant disparity loss (ngp_pl/losses.py:5-23; MiDaS, arXiv:1907.01341): both inputs (N) are
    centred on their median and divided by their mean absolute deviation; returns the (N) squared differences."""
    def normalise(d):
        shift = torch.median(d)
        scale = (d - shift).abs().mean()
        return (d - shift) / scale
    return (normalise(disp_pred) - normalise(disp_gt)) ** 2


class NeRFLoss(nn.Module):
    """ngp_pl/losses.py:26-40: squared colour error and the opacity entropy regulariser lambda_opa * (-o log o)."""

    def __init__(self, lambda_opa=1e-3):
        super().__init__()
        self.lambda_opa = lambda_opa

    def forward(self, results, target, **kwargs):
        o = results["opacity"] + 1e-10
        return {"rgb": (results["rgb"] - target["rgb"]) ** 2,
                "opacity": self.lambda_opa * (-o * torch.log(o))}   # pushes opacity to 0 or 1 (no floaters)
