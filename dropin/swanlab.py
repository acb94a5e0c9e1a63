"""Stand-in for the `swanlab` logging client the reference's utils/trainer.py:14 imports unconditionally.

`dropin/` sits first on sys.path, so this module shadows an installed swanlab: when the real package is importable from
any OTHER sys.path entry it is loaded and re-exported (logging then works as upstream, `use_swanlab=True` included);
otherwise the four entry points the trainer calls (utils/trainer.py:117-125,181-186,300-305,330) are no-ops, so a run
with `use_swanlab=True` trains instead of dying with AttributeError."""
import importlib.machinery
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load_real():
    paths = [p for p in sys.path if os.path.abspath(p or ".") != _HERE]
    spec = importlib.machinery.PathFinder.find_spec("swanlab", paths)
    if spec is None or spec.loader is None or os.path.abspath(os.path.dirname(spec.origin or "")) == _HERE:
        return None
    mod = importlib.util.module_from_spec(spec)
    saved = sys.modules.get("swanlab")
    sys.modules["swanlab"] = mod  # the package's own relative imports must find it under its real name
    try:
        spec.loader.exec_module(mod)
    except Exception:
        if saved is not None:
            sys.modules["swanlab"] = saved
        else:
            sys.modules.pop("swanlab", None)
        return None
    return mod


_real = _load_real()
if _real is not None:
    globals().update({k: v for k, v in vars(_real).items() if not k.startswith("__")})
else:
    class Image:  # swanlab.Image(path_or_array, caption=...)
        def __init__(self, *args, **kwargs):
            self.args, self.kwargs = args, kwargs

    def init(*args, **kwargs):
        return None

    def log(*args, **kwargs):
        return None

    def finish(*args, **kwargs):
        return None
