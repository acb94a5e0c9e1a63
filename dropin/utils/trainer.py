"""`from utils.trainer import DiffusionTrainer` (reference: train.py:23).

The trainer is a CALLER of the hot path (DDP wrap, EMA, checkpoints, logging: SURVEY.md section 2 row 8) and stays the reference's
own, unmodified code: this shim loads `utils/trainer.py` from the reference checkout by file path -- the checkout that holds the
script being run (dropin/run.py), or $DMC_REFERENCE_DIR -- and re-exports it.  Its `self.diffusion.p_losses(self.model, ...)` /
`loss.backward()` then run on the native UNet (models/unet_train.py)."""
import importlib.util
import os
import sys


def _reference_dir():
    cands = [os.environ.get("DMC_REFERENCE_DIR"), os.path.dirname(os.path.abspath(sys.argv[0])) if sys.argv and sys.argv[0] else None]
    for c in cands:
        if c and os.path.exists(os.path.join(c, "utils", "trainer.py")):
            return c
    raise ImportError("utils.trainer: the reference checkout (utils/trainer.py) was not found; set DMC_REFERENCE_DIR")


_spec = importlib.util.spec_from_file_location("_dmc_reference_trainer", os.path.join(_reference_dir(), "utils", "trainer.py"))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
DiffusionTrainer = _mod.DiffusionTrainer
