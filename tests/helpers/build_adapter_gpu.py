"""Builds tests/helpers/_bin/adapter_gpu (tests/helpers/adapter_gpu.cpp: the reference's main.cpp flow over b200_accel, against
the UNMODIFIED reference headers).  Only possible where /root/reference exists (the build container); the binary travels to the
GPU box with the snapshot.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RT_REFERENCE", "/root/reference") + "/include"
OUT = os.path.join(HERE, "_bin", "adapter_gpu")


def build(force: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(REF, "raytracer")):
        return OUT if os.path.exists(OUT) else None
    src = os.path.join(HERE, "adapter_gpu.cpp")
    deps = [src, os.path.join(REPO, "include", "b200_accel.hpp"), os.path.join(REPO, "include", "rt_b200.h"),
            os.path.join(REPO, "oracle", "rtsc_ref_scene.hpp")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    libdir = os.path.join(REPO, "simd-raytracer_b200")
    # -ffp-contract=off -fstack-reuse=none: the canonical flags of the reference build; $ORIGIN-relative rpath
    subprocess.check_call(["g++", "-std=c++23", "-O2", "-ffp-contract=off", "-fstack-reuse=none", "-Wno-dangling-reference", "-pthread",
                           "-I", os.path.join(REPO, "oracle", "stub"), "-I", os.path.join(REPO, "oracle", "cfg"), "-I", REF,
                           "-I", os.path.join(REPO, "include"), src, "-o", OUT, "-L", libdir, "-lrt_b200",
                           "-Wl,-rpath,$ORIGIN/../../../simd-raytracer_b200"])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
