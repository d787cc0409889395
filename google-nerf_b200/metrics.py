"""Image metrics used by the evaluation code (ngp_pl/metrics.py:4-16)."""
import torch


def mse(image_pred, image_gt, valid_mask=None, reduction="mean"):
    err = (image_pred - image_gt) ** 2
    if valid_mask is not None:
        err = err[valid_mask]
    return err.mean() if reduction == "mean" else err


@torch.no_grad()
def psnr(image_pred, image_gt, valid_mask=None, reduction="mean"):
    return -10.0 * torch.log10(mse(image_pred, image_gt, valid_mask, reduction))
