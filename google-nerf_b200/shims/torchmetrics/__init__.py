"""Minimal stand-in for the torchmetrics classes the reference uses (train.py:28-32,60-66,164,194-205):
PeakSignalNoiseRatio and StructuralSimilarityIndexMeasure with the call / compute / reset protocol.
Formulas: PSNR = 10 log10(data_range^2 / MSE) over everything seen since reset; SSIM = mean over pixels of the
standard index with an 11x11 Gaussian window (sigma 1.5), K1 = 0.01, K2 = 0.03, computed per image and averaged."""
import torch
from torch import nn
import torch.nn.functional as F


class Metric(nn.Module):
    def __init__(self):
        super().__init__()
        self.reset()

    def reset(self):
        raise NotImplementedError

    def update(self, preds, target):
        raise NotImplementedError

    def compute(self):
        raise NotImplementedError

    def forward(self, preds, target):
        """Accumulate, and return the value of this batch alone (torchmetrics' forward semantics).  The states are
        additive sums / counts."""
        saved = {k: getattr(self, k) for k in self._state}
        self.reset(); self.update(preds, target)
        value = self.compute()
        for k in self._state:
            setattr(self, k, saved[k] + getattr(self, k))
        return value


class PeakSignalNoiseRatio(Metric):
    _state = ("sum_sq", "count")

    def __init__(self, data_range=None, **kwargs):
        self.data_range = data_range
        super().__init__()

    def reset(self):
        self.sum_sq, self.count = 0.0, 0

    def update(self, preds, target):
        d = (preds.detach().float() - target.detach().float())
        self.sum_sq = self.sum_sq + (d * d).sum()
        self.count += d.numel()
        if self.data_range is None:
            self._range = float(target.max() - target.min())

    def compute(self):
        r = float(self.data_range) if self.data_range is not None else self._range
        mse = self.sum_sq / max(self.count, 1)
        return 10.0 * torch.log10(torch.as_tensor(r * r, dtype=torch.float32, device=getattr(mse, "device", None)) / mse)


def _gaussian_window(size, sigma, device, dtype):
    x = torch.arange(size, device=device, dtype=dtype) - (size - 1) / 2
    g = torch.exp(-(x * x) / (2 * sigma * sigma))
    g = g / g.sum()
    return g[:, None] * g[None, :]


def structural_similarity(preds, target, data_range=1.0, kernel_size=11, sigma=1.5, k1=0.01, k2=0.03):
    """(B,C,H,W) -> mean SSIM per image (B)."""
    preds, target = preds.float(), target.float()
    c = preds.shape[1]
    w = _gaussian_window(kernel_size, sigma, preds.device, preds.dtype).expand(c, 1, kernel_size, kernel_size).contiguous()
    pad = kernel_size // 2
    p = F.pad(preds, (pad,) * 4, mode="reflect"); t = F.pad(target, (pad,) * 4, mode="reflect")
    mu_p, mu_t = F.conv2d(p, w, groups=c), F.conv2d(t, w, groups=c)
    s_pp = F.conv2d(p * p, w, groups=c) - mu_p * mu_p
    s_tt = F.conv2d(t * t, w, groups=c) - mu_t * mu_t
    s_pt = F.conv2d(p * t, w, groups=c) - mu_p * mu_t
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    ssim = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p * mu_p + mu_t * mu_t + c1) * (s_pp + s_tt + c2))
    return ssim.flatten(1).mean(1)


class StructuralSimilarityIndexMeasure(Metric):
    _state = ("total", "count")

    def __init__(self, data_range=None, kernel_size=11, sigma=1.5, k1=0.01, k2=0.03, **kwargs):
        self.data_range, self.kernel_size, self.sigma, self.k1, self.k2 = data_range, kernel_size, sigma, k1, k2
        super().__init__()

    def reset(self):
        self.total, self.count = 0.0, 0

    def update(self, preds, target):
        r = float(self.data_range) if self.data_range is not None else float(target.max() - target.min())
        v = structural_similarity(preds.detach(), target.detach(), r, self.kernel_size, self.sigma, self.k1, self.k2)
        self.total = self.total + v.sum()
        self.count += v.numel()

    def compute(self):
        return self.total / max(self.count, 1)
