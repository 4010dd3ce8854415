"""Builds ``libmapdamage_b200.so`` in-tree with nvcc for sm_100a.

``python -m mapdamage_b200.build`` or ``build_library()``; nvcc cross-compiles
without a GPU.  The library is kept next to this file so that it travels with
the source tree (it is git-ignored, not installed into site-packages).
"""
import os
import shutil
import subprocess
from pathlib import Path

PACKAGE = Path(__file__).resolve().parent
CSRC = PACKAGE / "csrc"
LIBRARY = PACKAGE / "libmapdamage_b200.so"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).is_file():
        raise RuntimeError("nvcc not found; cannot build libmapdamage_b200.so")
    return nvcc


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.cpp")) + [PACKAGE.parent / "include" / "mapdamage_b200.h"]


def is_stale():
    if not LIBRARY.is_file():
        return True
    built = LIBRARY.stat().st_mtime
    return any(src.stat().st_mtime > built for src in sources())


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIBRARY
    cmd = [find_nvcc(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("MDG_NVCC_FLAGS", "").split()  # e.g. -DMDG_PHASE_CLOCKS: per-phase cycle counts, printed
    cmd += ["-o", str(LIBRARY), str(CSRC / "mdg_api.cu"), str(CSRC / "mdg_bamio.cpp"), str(CSRC / "mdg_sampler.cpp"), str(CSRC / "mdg_inflate.cpp"), "-lcudart", "-ldl", "-lz", "-lpthread"]
    env = dict(os.environ)
    result = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if result.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), result.stderr))
    if verbose:
        print(result.stderr)
    return LIBRARY


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
