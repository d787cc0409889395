"""Import-only stub: ngp_pl/datasets/scannet.py imports matplotlib (and trimesh, plyfile) at module level, and
datasets/__init__.py imports every dataset, so `train.py` cannot even start without them.  None of them is used on the
NSVF / synthetic path; any attribute access raises with an explanation."""


class _Unavailable:
    def __init__(self, name):
        self._name = name

    def __getattr__(self, item):
        raise RuntimeError(f"{self._name}.{item} is not available offline (import-only stub of google-nerf_b200/shims)")

    def __call__(self, *a, **k):
        raise RuntimeError(f"{self._name} is not available offline (import-only stub of google-nerf_b200/shims)")


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Unavailable(f"matplotlib.{name}")
