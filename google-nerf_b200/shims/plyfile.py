"""Import-only stub (see shims/matplotlib/__init__.py)."""


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("plyfile is not available offline (import-only stub of google-nerf_b200/shims)")


class PlyData(_Unavailable):
    pass


class PlyElement(_Unavailable):
    pass
