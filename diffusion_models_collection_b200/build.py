"""In-tree build of libdmc_b200.so with plain nvcc (no torch headers, no JIT cache).

    python -m diffusion_models_collection_b200.build [--force]

nvcc cross-compiles sm_100a without a GPU, so this runs in the build container; the resulting .so is
git-ignored but travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libdmc_b200.so")
OBJ = os.path.join(PKG, "csrc", "_build")

SOURCES = ["plan.cu", "sched.cu", "elementwise.cu", "conv_umma.cu", "conv_ref.cu", "attention.cu", "attention_umma.cu", "dit_ops.cu", "conv_wgrad.cu", "train_ops.cu", "attention_bwd_mma.cu", "optim.cu", "dit_train_ops.cu",
           "conv_inst_256_1_2.cu", "conv_inst_128_2_2.cu", "conv_inst_256_1_1.cu", "conv_inst_128_2_1.cu", "conv_inst_128_1_1.cu",
           "conv_inst_64_1_1.cu", "conv_inst_32_1_1.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libdmc_b200.so")
    return exe


def _signature():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(ROOT, "include")):
        for name in sorted(os.listdir(root)):
            path = os.path.join(root, name)
            if os.path.isfile(path) and name.split(".")[-1] in ("cu", "cuh", "h"):
                h.update(name.encode())
                h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    sig = _signature()
    stamp = os.path.join(OBJ, "signature.txt")
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == sig:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OBJ, src.replace(".cu", ".ptxas.log"))
        with open(log, "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 4, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(sig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
