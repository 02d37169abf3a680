"""In-tree build of the CUDA extension (sm_100a only).

``nvcc`` cross-compiles without a GPU, so this runs in the CPU container and the
resulting ``csrc/liblm_bev.so`` travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# LM_BEV_LIB: load / build another file name inside csrc/ (A/B tuning builds side by side); the default is the product
LIB_PATH = os.path.join(CSRC, os.path.basename(os.environ.get("LM_BEV_LIB", "") or "liblm_bev.so"))
SOURCES = [os.path.join(CSRC, "lm_bev.cu"), os.path.join(CSRC, "lm_post.cu")]
HEADERS = [os.path.join(ROOT, "include", h) for h in ("lm_bev.h", "lm_las.h", "lm_post.h")] + \
          [os.path.join(CSRC, h) for h in ("lm_dev.cuh", "lm_sweep.cuh", "lm_host.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",              # no FMA contraction: float keys must match the spec bit for bit
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build liblm_bev.so")
    return exe


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in SOURCES + HEADERS)


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into csrc/liblm_bev.so; returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    extra = os.environ.get("LM_BEV_NVCC_EXTRA", "").split()      # tuning builds, e.g. -DLM_BIN_THREADS=256
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-I", os.path.join(ROOT, "include"), "-o", LIB_PATH, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
