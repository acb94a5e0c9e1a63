"""TEST INFRASTRUCTURE -- locates the reference itself (not the restatement): the checkout at $DMC_REFERENCE_DIR or
/root/reference when present (build container), else the archive oracle/_ref/reference_src.zip written by
oracle/make_ref.py (the GPU box).  Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may
import this module; the product package never does."""
import atexit
import importlib
import os
import shutil
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
ZIP = os.path.join(HERE, "_ref", "reference_src.zip")
_extracted = None


def available():
    return os.path.exists(ZIP) or os.path.isdir(_checkout() or "")


def _checkout():
    for c in (os.environ.get("DMC_REFERENCE_DIR"), "/root/reference"):
        if c and os.path.exists(os.path.join(c, "sample.py")):
            return c
    return None


def reference_dir():
    """a directory holding the reference's files as a checkout would (for running its scripts by path)"""
    global _extracted
    c = _checkout()
    if c:
        return c
    if _extracted is None:
        if not os.path.exists(ZIP):
            raise FileNotFoundError("the reference is not available: no checkout and no oracle/_ref/reference_src.zip "
                                    "(python oracle/make_ref.py in the build container)")
        _extracted = tempfile.mkdtemp(prefix="dmc_reference_")
        with zipfile.ZipFile(ZIP) as z:
            z.extractall(_extracted)
        atexit.register(shutil.rmtree, _extracted, ignore_errors=True)
    return _extracted


def source_kind():
    return "checkout" if _checkout() else "oracle/_ref/reference_src.zip"


def import_reference():
    """-> dict(UNet, DiT, DDPM, DDIM, config_unet, config_dit, set_seed): the reference's OWN classes.  The reference's
    top-level package names (`models`, `diffusion`, `configs`) are generic, so they are imported with the reference first
    on sys.path and then REMOVED from sys.modules / sys.path again: the caller keeps the class objects, later imports of
    same-named packages (dropin/) are unaffected."""
    ref = reference_dir()
    names = ("models", "diffusion", "configs", "utils")
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in names}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, ref)
    try:
        unet = importlib.import_module("models.unet")
        dit = importlib.import_module("models.dit")
        ddpm = importlib.import_module("diffusion.ddpm")
        ddim = importlib.import_module("diffusion.ddim")
        cu = importlib.import_module("configs.cifar10_unet")
        cd = importlib.import_module("configs.cifar10_dit")
        out = dict(UNet=unet.UNet, DiT=dit.DiT, DDPM=ddpm.DDPM, DDIM=ddim.DDIM, config_unet=cu.config, config_dit=cd.config,
                   dir=ref, kind=source_kind())
        try:  # models/__init__ pulls dim.py (mamba_ssm is optional there)
            out["DiM"] = importlib.import_module("models.dim").DiM
        except Exception:  # pragma: no cover
            out["DiM"] = None
    finally:
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k.split(".")[0] in names]:
            del sys.modules[k]
        sys.modules.update(saved)
    return out
