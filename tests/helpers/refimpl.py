"""ctypes binding to oracle/_ref/libref_*.so - the UNMODIFIED reference compiled by oracle/Makefile.

Test infrastructure only (tests/, bench.py's reference / cpu_baseline legs).  Never imported by the product.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF_DIR = os.path.join(REPO, "oracle", "_ref")

QUERY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("t", "<f4"), ("u", "<f4"), ("v", "<f4"),
                        ("tri", "<i4"), ("cull", "<u4")])


def cpu_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    fl = line.split()
                    return all(f in fl for f in ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"))
    except OSError:
        pass
    return False


def variant_path(fp="canon", isa="v3", spp=1, depth=5, gi=0, kd_depth=8, kd_leaf=64) -> str:
    return os.path.join(REF_DIR, f"libref_{fp}_{isa}_s{spp}d{depth}g{gi}_k{kd_depth}x{kd_leaf}.so")


def available(**kw) -> bool:
    if kw.get("isa") == "v4" and not cpu_has_avx512():
        return False
    return os.path.exists(variant_path(**kw))


class RefImpl:
    """One loaded reference variant + one scene."""

    def __init__(self, rtsc_path: str, **variant):
        self.path = variant_path(**variant)
        self.lib = C.CDLL(self.path)
        L = self.lib
        L.ref_scene_load.restype = C.c_void_p
        L.ref_scene_load.argtypes = [C.c_char_p]
        L.ref_scene_free.argtypes = [C.c_void_p]
        L.ref_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_render.restype = C.c_double
        L.ref_render.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_count.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_record.restype = C.c_uint64
        L.ref_record.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_occluded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.h = L.ref_scene_load(rtsc_path.encode())
        if not self.h:
            raise RuntimeError(f"reference failed to load {rtsc_path}")
        info = np.zeros(12, np.uint64)
        bs = C.c_double(0)
        L.ref_info(self.h, info.ctypes.data, C.byref(bs))
        (self.width, self.height, self.n_tris, self.n_nodes, self.n_packs, self.W, self.spp, self.max_ray_depth,
         self.gi_rays, self.kd_max_depth, self.kd_max_leaf, self.threads) = (int(x) for x in info)
        self.build_seconds = bs.value

    def close(self):
        if self.h:
            self.lib.ref_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, want_image=True):
        img = np.zeros((self.height, self.width, 3), np.float32) if want_image else None
        sec = self.lib.ref_render(self.h, img.ctypes.data if want_image else None)
        return img, sec

    def count(self):
        c = np.zeros(4, np.uint64)
        self.lib.ref_count(self.h, c.ctypes.data)
        return dict(cull=int(c[0]), cull_hit=int(c[1]), nocull=int(c[2]), nocull_hit=int(c[3]))

    def record(self, cap: int):
        out = np.zeros(cap, QUERY_DTYPE)
        img = np.zeros((self.height, self.width, 3), np.float32)
        n = self.lib.ref_record(self.h, out.ctypes.data, cap, img.ctypes.data)
        return out[:n], img

    def trace(self, rays: np.ndarray, cull: bool):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        tuv = np.zeros((n, 3), np.float32)
        tri = np.zeros(n, np.int32)
        self.lib.ref_trace(self.h, rays.ctypes.data, n, 1 if cull else 0, tuv.ctypes.data, tri.ctypes.data)
        return tuv, tri

    def occluded(self, rays: np.ndarray, max_t: np.ndarray):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        max_t = np.ascontiguousarray(max_t, np.float32)
        out = np.zeros(len(rays), np.uint8)
        self.lib.ref_occluded(self.h, rays.ctypes.data, max_t.ctypes.data, len(rays), out.ctypes.data)
        return out

    def tree(self):
        node5 = np.zeros((self.n_nodes, 5), np.uint64)
        boxes = np.zeros((self.n_nodes, 6), np.float32)
        packs = np.zeros((self.n_packs, self.W), np.uint64)
        self.lib.ref_tree(self.h, node5.ctypes.data, boxes.ctypes.data, packs.ctypes.data)
        return node5, boxes, packs


def quantise(rgb: np.ndarray) -> np.ndarray:
    """uint8(255.999 * clamp(c, 0, 1)) with the product computed in double - io/image/ppm.hpp:17-19."""
    c = np.clip(rgb.astype(np.float32), np.float32(0), np.float32(1)).astype(np.float64)
    return (255.999 * c).astype(np.uint8)
