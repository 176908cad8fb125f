"""Builds simd-raytracer_b200/librt_b200.so (the C-ABI library of include/rt_b200.h) in-tree, for sm_100a only.

    python simd-raytracer_b200/build.py [--force]

nvcc cross-compiles without a GPU.  Flags that are part of the numerical contract (DESIGN.md "Numerics"):
  device: -fmad=false            no FMA contraction; the FAST variants spell fmaf() explicitly
  host  : -ffp-contract=off      triangle / vertex normals feed shading and must equal the canonical reference build
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIB = os.path.join(HERE, "librt_b200.so")
BUILD = os.path.join(HERE, "build")

CU = [os.path.join(HERE, "csrc", "rt_api.cu")]
CPP = [os.path.join(HERE, "host", f) for f in ("kd_build.cpp", "bvh_build.cpp", "scene_io.cpp", "jpeg_decode.cpp")]
import glob  # noqa: E402

DEPS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(HERE, "host", "*.hpp"))) + \
       [os.path.join(REPO, "include", "rt_b200.h"), os.path.abspath(__file__)]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-std=c++20", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "-Xptxas", "-v"]
CXX_FLAGS = ["-std=c++20", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-Wall", "-Wextra"]


def _stale(out: str, srcs) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """defines / out: developer sweeps only - a variant library with -D overrides of the tuning knobs (csrc/rt_stream.cuh),
    loaded through the RT_B200_LIB environment variable; the product is always the default build."""
    os.makedirs(BUILD, exist_ok=True)
    objs = []
    lib = out or LIB
    tag = "" if not out else "." + os.path.splitext(os.path.basename(out))[0]
    for src in CU:
        obj = os.path.join(BUILD, os.path.basename(src) + tag + ".o")
        if force or _stale(obj, [src] + DEPS):
            cmd = [NVCC] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            with open(os.path.join(BUILD, os.path.basename(src) + tag + ".ptxas.log"), "w") as fh:
                fh.write(r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
            if verbose:
                sys.stderr.write(r.stderr)
        objs.append(obj)
    for src in CPP:
        obj = os.path.join(BUILD, os.path.basename(src) + ".o")
        if force or _stale(obj, [src] + DEPS):
            subprocess.check_call(["g++"] + CXX_FLAGS + ["-c", src, "-o", obj])
        objs.append(obj)
    if force or _stale(lib, objs):
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs +
                              ["-cudart", "static", "-Xlinker", "--exclude-libs,ALL"])
    return lib


if __name__ == "__main__":
    defs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--define=")]
    outs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
