"""CPU: the C-ABI shared library builds, loads without a GPU and exports every symbol include/dmc.h declares
(no compute calls here); the ctypes mirror in _lib.py covers exactly the same set; struct sizes agree with the C
compiler's view of the header."""

import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dmc.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"DMC_API\s+[\w\s\*]+?\b(dmc_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from diffusion_models_collection_b200 import build

    return build.build()


def test_header_declares_something():
    syms = declared_symbols()
    assert "dmc_ddim_step" in syms and "dmc_plan_run" in syms and len(syms) >= 25


def test_library_exports_every_declared_symbol(lib_path):
    lib = C.CDLL(lib_path)
    for s in declared_symbols():
        assert getattr(lib, s, None) is not None, s
    lib.dmc_abi_version.restype = C.c_int
    assert lib.dmc_abi_version() == 1


def test_ctypes_mirror_matches_header():
    from diffusion_models_collection_b200 import _lib

    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_struct_layouts_match_the_c_compiler():
    """sizeof() of every descriptor as gcc sees include/dmc.h == ctypes.sizeof of the Python mirror."""
    from diffusion_models_collection_b200 import _lib

    names = {
        "dmc_ddim_coef": _lib.DdimCoef, "dmc_ddpm_coef": _lib.DdpmCoef, "dmc_guidance": _lib.Guidance,
        "dmc_cond_desc": _lib.CondDesc, "dmc_stem_desc": _lib.StemDesc, "dmc_gn_stats_desc": _lib.GnStatsDesc,
        "dmc_gn_apply_desc": _lib.GnApplyDesc, "dmc_conv_desc": _lib.ConvDesc, "dmc_attn_desc": _lib.AttnDesc,
        "dmc_upsample_desc": _lib.UpsampleDesc, "dmc_step_desc": _lib.StepDesc, "dmc_dit_cond_desc": _lib.DitCondDesc,
        "dmc_patch_embed_desc": _lib.PatchEmbedDesc, "dmc_ln_mod_desc": _lib.LnModDesc, "dmc_head_desc": _lib.HeadDesc, "dmc_wgrad_desc": _lib.WgradDesc, "dmc_gn_bwd_desc": _lib.GnBwdDesc, "dmc_pack_item": _lib.PackItem, "dmc_opt_item": _lib.OptItem, "dmc_opt_chunk": _lib.OptChunk, "dmc_adamw_desc": _lib.AdamWDesc,
        "dmc_attn_bwd_desc": _lib.AttnBwdDesc, "dmc_dit_glm_desc": _lib.DitGlmDesc, "dmc_dit_glm_bwd_desc": _lib.DitGlmBwdDesc,
    }
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in names)
    prog = f'#include <stdio.h>\n#include "dmc.h"\nint main(void){{{body}return 0;}}\n'
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    for line in out.strip().splitlines():
        n, sz = line.split()
        assert C.sizeof(names[n]) == int(sz), n


def test_product_path_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from diffusion_models_collection_b200 import _lib
    from diffusion_models_collection_b200.models import UNet

    net = UNet()
    with pytest.raises(_lib.DmcError):
        net(torch.zeros(1, 3, 32, 32), torch.zeros(1, dtype=torch.long))
    with pytest.raises(_lib.DmcError):
        _lib.load(require_device=True)


def test_oracle_is_not_imported_by_the_product():
    code = ("import sys; import diffusion_models_collection_b200.models, diffusion_models_collection_b200.diffusion;"
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
