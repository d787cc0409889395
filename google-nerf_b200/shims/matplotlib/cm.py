from . import _Unavailable


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Unavailable(f"matplotlib.cm.{name}")
