#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- packs the reference checkout's own Python sources into oracle/_ref/reference_src.zip.

    python oracle/make_ref.py [--reference /root/reference]

The reference (sunyzhi55/Diffusion_Models_Collection) is pure Python: there is nothing to compile, so the "build" of the
`oracle/_ref` checker is an archive of its `*.py` files, byte for byte, taken from where they lie under /root/reference.
The archive is a build artefact: `oracle/_ref/` is git-ignored (never committed, no reference source enters the history)
but not gpurun-ignored, so it travels to the GPU box, where /root/reference does not exist.  It is used ONLY by

  * `bench.py --impl reference` and the `cpu_baseline` leg: the reference's own `DDIM.sample(UNet, (16, 3, 32, 32))`
    (BASELINE.json configs[0]) timed on the host cores, imported straight from the zip (zipimport);
  * `tests/test_gpu_dropin_sample.py`: the reference's UNMODIFIED sample.py executed through dropin/run.py on the B200;
  * `tests/` checks that pin the oracle restatement against the live reference.

Nothing under `diffusion_models_collection_b200/` imports it (tests/test_cabi.py enforces that).
`__graft_entry__.build()` runs this script whenever /root/reference is present."""
import argparse
import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
ZIP = os.path.join(OUT_DIR, "reference_src.zip")
MANIFEST = os.path.join(OUT_DIR, "MANIFEST.json")


def collect(ref):
    files = []
    for root, dirs, names in os.walk(ref):
        dirs[:] = sorted(d for d in dirs if d not in (".git", "__pycache__", "assets", "docs"))
        for n in sorted(names):
            if n.endswith(".py") or n in ("requirements.txt", "LICENSE"):
                p = os.path.join(root, n)
                files.append((os.path.relpath(p, ref), p))
    return files


def build(ref="/root/reference", force=False):
    """-> path of the archive (built / refreshed when the reference is present), or None when there is neither"""
    if not os.path.isdir(ref):
        return ZIP if os.path.exists(ZIP) else None
    files = collect(ref)
    man = {rel: hashlib.sha256(open(p, "rb").read()).hexdigest() for rel, p in files}
    if not force and os.path.exists(ZIP) and os.path.exists(MANIFEST):
        try:
            if json.load(open(MANIFEST)).get("files") == man:
                return ZIP
        except Exception:
            pass
    os.makedirs(OUT_DIR, exist_ok=True)
    with zipfile.ZipFile(ZIP, "w", zipfile.ZIP_DEFLATED) as z:
        for rel, p in files:
            zi = zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0))  # reproducible archive
            zi.compress_type = zipfile.ZIP_DEFLATED
            z.writestr(zi, open(p, "rb").read())
    json.dump({"source": ref, "files": man}, open(MANIFEST, "w"), indent=1, sort_keys=True)
    return ZIP


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    z = build(a.reference, a.force)
    print(z if z else "no reference checkout and no archive")
    sys.exit(0 if z else 1)
