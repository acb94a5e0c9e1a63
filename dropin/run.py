#!/usr/bin/env python
"""Runs one of the reference's own scripts (sample.py, ...) UNMODIFIED against the B200 hot path:

    python dropin/run.py /path/to/reference/sample.py --checkpoint ckpt.pth --sampling_method ddim --cfg_scale 3 \
        --labels 0,1 --device cuda

Python puts a script's own directory first on sys.path, so running `python /path/to/reference/sample.py` with
PYTHONPATH alone would still import the reference's `models` / `diffusion` packages.  This launcher puts `dropin/` (the
packages named like the reference's, re-exporting the native classes) and the repo root in front, then executes the
script with runpy -- the script's directory comes AFTER the shims, so `from models import UNet` resolves here while
`from datasets import ...` (train.py:22) still finds the reference's package."""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    sys.argv = [script] + sys.argv[2:]
    # drop the launcher's own directory entry and the script's directory if present, then put the shims first and the script's
    # directory right behind them: `models` / `diffusion` / `utils` / `configs` resolve to the shims, the packages only the
    # reference has (`datasets`, `metrics`) to the reference's own -- ahead of same-named packages in site-packages
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (HERE, os.path.dirname(script))]
    sys.path[:0] = [HERE, ROOT, os.path.dirname(script)]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
