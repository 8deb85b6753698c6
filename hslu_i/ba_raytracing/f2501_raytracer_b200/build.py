"""In-tree build of librt_b200.so: hand-written CUDA for sm_100a + the C++ host layer.

    python -m hslu_i.ba_raytracing.f2501_raytracer_b200.build [--force] [--verbose]

-fmad=false / -ffp-contract=off: no implicit FMA contraction anywhere; the explicit fmaf() calls
are the only fused operations (they mirror the reference's `mul_add` call sites), which is what
makes hit/miss decisions bit-identical to the CPU oracle.  Never --use_fast_math.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parents[2]
CSRC = PKG / "csrc"
OUT = Path(os.environ.get("RT_LIB_PATH", PKG / "librt_b200.so"))  # RT_LIB_PATH / RT_NVCC_EXTRA: tuning experiments only

SOURCES = [CSRC / "rt_api.cu", CSRC / "rt_bvh_device.cu", CSRC / "rt_bvh.cpp", CSRC / "rt_sampler.cpp", CSRC / "rt_error.cpp", CSRC / "host" / "rt_host.cpp"]
HEADERS = [CSRC / "rt_device.cuh", CSRC / "rt_kernels.cuh", CSRC / "rt_bvh.h", CSRC / "rt_bvh_device.h", CSRC / "rt_internal.h", CSRC / "host" / "rt_host.hpp",
           ROOT / "include" / "rt_api.h", ROOT / "include" / "rt_host.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
    "-shared", "-cudart", "static", "-lpthread",
]


def nvcc_path() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS + [Path(__file__)])


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    cmd = [nvcc_path(), *NVCC_FLAGS, f"-I{ROOT / 'include'}", f"-I{CSRC}", f"-I{CSRC / 'host'}"]
    cmd += os.environ.get("RT_NVCC_EXTRA", "").split()
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", str(OUT), *map(str, SOURCES)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)
