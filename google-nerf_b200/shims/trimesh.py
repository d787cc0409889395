"""Import-only stub (see shims/matplotlib/__init__.py)."""


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    raise RuntimeError(f"trimesh.{name} is not available offline (import-only stub of google-nerf_b200/shims)")
