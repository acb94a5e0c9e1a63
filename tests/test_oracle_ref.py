"""CPU: oracle/_ref -- the reference itself, packed by oracle/make_ref.py so that it can travel to the GPU box (where
/root/reference does not exist).  The archive must hold exactly the checkout's files, import WITHOUT the checkout, and
reproduce the committed golden eps bit for bit (i.e. what ships to the box is the code that wrote the fixtures)."""

import hashlib
import json
import os
import zipfile

import numpy as np
import pytest
import torch

from diffusion_models_collection_b200 import synth
from oracle import make_ref, ref_loader
from tests.golden_cases import SMALL_UNET, UNET_CASES, case_inputs

REF = "/root/reference"


def _archive():
    z = make_ref.build()
    if z is None:
        pytest.skip("neither /root/reference nor oracle/_ref/reference_src.zip is present")
    return z


def test_archive_is_the_checkout_byte_for_byte():
    z = _archive()
    if not os.path.isdir(REF):
        pytest.skip("no checkout to compare against (GPU box)")
    man = json.load(open(make_ref.MANIFEST))["files"]
    with zipfile.ZipFile(z) as zf:
        names = set(zf.namelist())
        assert names == set(man)
        for rel in ("models/unet.py", "models/dit.py", "diffusion/ddim.py", "diffusion/ddpm.py", "sample.py", "utils/trainer.py"):
            assert rel in names
            data = zf.read(rel)
            assert data == open(os.path.join(REF, rel), "rb").read()
            assert hashlib.sha256(data).hexdigest() == man[rel]


def test_archive_is_not_tracked_by_git_but_travels():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ignore = open(os.path.join(root, ".gitignore")).read().split()
    assert "oracle/_ref/" in ignore
    gri = os.path.join(root, ".gpurunignore")
    if os.path.exists(gri):
        assert not any(l.strip().startswith("oracle") for l in open(gri))


def test_reference_imported_from_the_archive_reproduces_the_golden_eps(monkeypatch, golden):
    _archive()
    monkeypatch.setattr(ref_loader, "_checkout", lambda: None)  # what the GPU box sees: the archive only
    monkeypatch.setattr(ref_loader, "_extracted", None)
    ref = ref_loader.import_reference()
    assert ref["kind"] == "oracle/_ref/reference_src.zip" and not ref["dir"].startswith(REF)
    c = UNET_CASES["small_cond"]
    net = ref["UNet"](**SMALL_UNET, num_classes=10).eval()
    net.load_state_dict(synth.make_unet_state_dict(SMALL_UNET, 10, seed=c["wseed"]), strict=True)
    x, t, y = case_inputs(c)
    with torch.no_grad():
        eps = net(x, t, y)
    assert np.array_equal(eps.numpy(), golden["unet"]["small_cond"])
    # the loader leaves no generic package names behind (dropin/ uses the same ones)
    import sys

    assert not any(k.split(".")[0] in ("models", "diffusion", "configs") and "reference" in str(getattr(m, "__file__", ""))
                   for k, m in sys.modules.items())
