"""Minimal stand-in for the `imageio` calls of the reference (train.py:209-214,291-298; datasets/color_utils.py):
imread / imsave / imwrite through PIL, mimsave as an animated GIF next to the requested path (no ffmpeg offline)."""
import os
import warnings

import numpy as np
from PIL import Image


def imread(path):
    return np.asarray(Image.open(path))


def imsave(path, array):
    a = np.asarray(array)
    if a.dtype != np.uint8:
        a = (np.clip(a, 0.0, 1.0) * 255).astype(np.uint8) if a.dtype.kind == "f" else a.astype(np.uint8)
    Image.fromarray(a).save(path)


imwrite = imsave


def mimsave(path, frames, fps=30, **kwargs):
    frames = [Image.fromarray(np.asarray(f).astype(np.uint8)) for f in frames]
    if not frames:
        return
    base, ext = os.path.splitext(path)
    if ext.lower() != ".gif":
        warnings.warn(f"imageio shim: no video encoder offline, writing {base}.gif instead of {path}")
        path = base + ".gif"
    frames[0].save(path, save_all=True, append_images=frames[1:], duration=max(1, int(1000 / fps)), loop=0)
