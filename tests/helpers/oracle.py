"""ctypes binding to oracle/librt_oracle.so (the C restatement).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
ORACLE_DIR = os.path.join(REPO, "oracle")
LIB = os.path.join(ORACLE_DIR, "librt_oracle.so")

RNG_MINSTD, RNG_PHILOX = 0, 1
KIND_PRIMARY, KIND_SHADOW, KIND_REFLECT, KIND_REFRACT, KIND_GI = range(5)
N_COUNTS = 12

RECORD_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("t", "<f4"), ("u", "<f4"), ("v", "<f4"),
                         ("tri", "<i4"), ("cull", "<u4"), ("kind", "<u4")])


class Params(C.Structure):
    _fields_ = [("fov_degrees", C.c_double), ("eps", C.c_float), ("shadow_bias", C.c_float),
                ("reflection_bias", C.c_float), ("refraction_bias", C.c_float), ("spp", C.c_uint32),
                ("max_ray_depth", C.c_uint32), ("gi_rays", C.c_uint32), ("seed", C.c_uint32), ("rng", C.c_uint32),
                ("sample_offset", C.c_uint32), ("spp_total", C.c_uint32), ("raw_sum", C.c_uint32)]


def build(force: bool = False) -> str:
    src = [os.path.join(ORACLE_DIR, f) for f in ("rt_oracle.c", "rt_oracle.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ro_default_params.argtypes = [C.POINTER(Params)]
        L.ro_scene_load_rtsc.restype = C.c_void_p
        L.ro_scene_load_rtsc.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32]
        L.ro_scene_from_rtsc_bytes.restype = C.c_void_p
        L.ro_scene_from_rtsc_bytes.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ro_scene_free.argtypes = [C.c_void_p]
        L.ro_scene_info.argtypes = [C.c_void_p, C.c_void_p]
        L.ro_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ro_geometry.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.ro_trace.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ro_occluded.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ro_primary_rays.argtypes = [C.c_void_p, C.POINTER(Params)] + [C.c_uint32] * 4 + [C.c_void_p]
        L.ro_render.argtypes = [C.c_void_p, C.POINTER(Params)] + [C.c_uint32] * 4 + [C.c_void_p, C.c_int, C.c_void_p]
        L.ro_record_frame.restype = C.c_uint64
        L.ro_record_frame.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_uint64, C.c_void_p]
        L.ro_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def default_params(**kw) -> Params:
    p = Params()
    lib().ro_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().ro_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


class Oracle:
    def __init__(self, rtsc: bytes | str, kd_max_depth: int = 8, kd_max_leaf: int = 64):
        L = lib()
        if isinstance(rtsc, (bytes, bytearray)):
            self.h = L.ro_scene_from_rtsc_bytes(bytes(rtsc), len(rtsc), kd_max_depth, kd_max_leaf)
        else:
            self.h = L.ro_scene_load_rtsc(rtsc.encode(), kd_max_depth, kd_max_leaf)
        if not self.h:
            raise RuntimeError("oracle failed to load scene")
        info = np.zeros(8, np.uint64)
        L.ro_scene_info(self.h, info.ctypes.data)
        (self.width, self.height, self.n_tris, self.n_nodes, self.n_refs, self.n_leaves, self.max_leaf_refs,
         self.tree_depth) = (int(x) for x in info)

    def close(self):
        if getattr(self, "h", None):
            lib().ro_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tree(self):
        node5 = np.zeros((self.n_nodes, 5), np.uint64)
        boxes = np.zeros((self.n_nodes, 6), np.float32)
        refs = np.zeros(self.n_refs, np.uint32)
        lib().ro_tree(self.h, node5.ctypes.data, boxes.ctypes.data, refs.ctypes.data)
        return node5, boxes, refs

    def geometry(self, n_verts: int):
        tri9 = np.zeros((self.n_tris, 9), np.float32)
        fn = np.zeros((self.n_tris, 3), np.float32)
        vn = np.zeros((n_verts, 3), np.float32)
        vidx = np.zeros((self.n_tris, 3), np.uint32)
        mesh = np.zeros(self.n_tris, np.uint32)
        lib().ro_geometry(self.h, tri9.ctypes.data, fn.ctypes.data, vn.ctypes.data, vidx.ctypes.data, mesh.ctypes.data)
        return tri9, fn, vn, vidx, mesh

    def trace(self, rays, cull: bool, eps: float = np.float32(1e-6), want_counts=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        tuv = np.zeros((n, 3), np.float32)
        tri = np.zeros(n, np.int32)
        counts = np.zeros(N_COUNTS, np.uint64)
        lib().ro_trace(self.h, C.c_float(eps), rays.ctypes.data, n, 1 if cull else 0, tuv.ctypes.data,
                       tri.ctypes.data, counts.ctypes.data)
        return (tuv, tri, counts) if want_counts else (tuv, tri)

    def occluded(self, rays, max_t, params: Params | None = None):
        p = params or default_params()
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        max_t = np.ascontiguousarray(max_t, np.float32)
        out = np.zeros(len(rays), np.uint8)
        counts = np.zeros(N_COUNTS, np.uint64)
        lib().ro_occluded(self.h, C.byref(p), rays.ctypes.data, max_t.ctypes.data, len(rays), out.ctypes.data,
                          counts.ctypes.data)
        return out, counts

    def primary_rays(self, params: Params | None = None, rect=None):
        p = params or default_params()
        x0, y0, x1, y1 = rect or (0, 0, self.width, self.height)
        rays = np.zeros(((y1 - y0) * (x1 - x0), 6), np.float32)
        lib().ro_primary_rays(self.h, C.byref(p), x0, y0, x1, y1, rays.ctypes.data)
        return rays

    def render(self, params: Params | None = None, rect=None, threads: int = 0, out=None):
        p = params or default_params()
        x0, y0, x1, y1 = rect or (0, 0, self.width, self.height)
        img = out if out is not None else np.zeros((self.height, self.width, 3), np.float32)
        counts = np.zeros(N_COUNTS, np.uint64)
        lib().ro_render(self.h, C.byref(p), x0, y0, x1, y1, img.ctypes.data, threads, counts.ctypes.data)
        return img, counts

    def record_frame(self, params: Params | None = None, cap: int = 1 << 24):
        p = params or default_params()
        out = np.zeros(cap, RECORD_DTYPE)
        img = np.zeros((self.height, self.width, 3), np.float32)
        n = lib().ro_record_frame(self.h, C.byref(p), out.ctypes.data, cap, img.ctypes.data)
        return out[:n], img
