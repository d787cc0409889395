"""LPIPS needs pretrained VGG weights, which are not available offline: the class exists so that
`from torchmetrics.image.lpip import LearnedPerceptualImagePatchSimilarity` (train.py:32) resolves, and raises when it
is actually requested (`--eval_lpips`)."""
from torch import nn


class LearnedPerceptualImagePatchSimilarity(nn.Module):
    def __init__(self, net_type="vgg", **kwargs):
        super().__init__()
        raise RuntimeError("LPIPS is unavailable in this environment (no pretrained VGG weights offline); "
                           "run without --eval_lpips")
