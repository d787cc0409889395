"""Importable alias of the hyphenated package directory ``google-nerf_b200/`` (a hyphen is not a valid
module name).  ``import google_nerf_b200`` executes ``google-nerf_b200/__init__.py`` with its
directory as the package path, so ``google_nerf_b200.models.rendering`` etc. resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "google-nerf_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
