"""Python host binding of the B200 ray-tracing backend (ctypes over the C ABI of include/rt_b200.h).

The product is the shared library ``librt_b200.so`` next to this file (CUDA kernels for sm_100a + the extern "C" layer);
this module only marshals numpy arrays / raw device pointers into it, for the tests, the benchmark and Python callers.
There is no CPU fallback anywhere: if the library is missing this module raises at import, and every compute call on a
machine without an sm_100 device raises ``RtError(RT_ERR_NO_DEVICE)``.

The directory name contains a hyphen, so import it with
``importlib.import_module("simd-raytracer_b200")`` (the repo root must be on ``sys.path``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# RT_B200_LIB: developer override for parameter sweeps (a variant built by build.py --define=... --out=...)
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(HERE, "librt_b200.so")

RT_OK, RT_ERR_BAD_ARG, RT_ERR_NO_DEVICE, RT_ERR_CUDA, RT_ERR_IO, RT_ERR_PARSE, RT_ERR_OOM, RT_ERR_UNSUPPORTED = range(8)
RT_FRAME_RERENDERED = 8
RT_ERR_TIMEOUT = 9
FLAG_RAW_SUM, FLAG_FAST_MATH, FLAG_ORDERED = 1, 2, 4
DEVICE_HOST_ONLY = -1
TEX_ALBEDO, TEX_EDGES, TEX_CHECKER, TEX_BITMAP = range(4)
MAT_DIFFUSE, MAT_REFLECTIVE, MAT_REFRACTIVE, MAT_CONSTANT, MAT_TEXTURE = range(5)

HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("tri", "<i4")])

# every symbol include/rt_b200.h declares
ABI_SYMBOLS = (
    "rt_abi_version", "rt_status_string", "rt_last_error", "rt_default_build_opts", "rt_default_params",
    "rt_scene_create", "rt_scene_create_from_crtscene", "rt_scene_create_from_rtsc", "rt_scene_export_rtsc", "rt_scene_destroy",
    "rt_scene_get_info", "rt_scene_get_tree", "rt_scene_get_device_layout", "rt_scene_get_bvh_layout",
    "rt_scene_get_geometry",
    "rt_trace_closest", "rt_trace_occluded", "rt_trace_closest_device", "rt_trace_occluded_device",
    "rt_render_frame", "rt_render_frame_rgb8", "rt_render_frame_begin", "rt_render_frame_rgb8_begin", "rt_frame_wait", "rt_alloc_pinned", "rt_free_pinned", "rt_render_frame_device_begin", "rt_render_frame_device", "rt_trace_primary",
    "rt_get_counters",
    "rt_resolve_sum_device",
    "rt_peer_group_create", "rt_peer_group_connect", "rt_peer_group_connect_local", "rt_peer_framebuffer", "rt_peer_result_rgb",
    "rt_peer_result_rgb8", "rt_peer_combine", "rt_peer_signal_ready", "rt_peer_reduce_resolve", "rt_peer_wait_done",
    "rt_peer_download_result", "rt_peer_read_result", "rt_peer_group_destroy", "rt_peer_host_result_attach",
)


class RtError(RuntimeError):
    def __init__(self, status: int, detail: str):
        super().__init__(f"rt_b200 status {status}: {detail}")
        self.status = status


class LightDesc(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("intensity", C.c_float)]


class TextureDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("c0", C.c_float * 3), ("c1", C.c_float * 3), ("scalar", C.c_float),
                ("bmp_w", C.c_uint32), ("bmp_h", C.c_uint32), ("bmp_off", C.c_uint32)]


class MaterialDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("albedo", C.c_float * 3), ("ior", C.c_float), ("smooth_shading", C.c_uint32),
                ("texture", C.c_int32)]


class MeshDesc(C.Structure):
    _fields_ = [("material", C.c_uint32), ("n_vertices", C.c_uint32), ("n_uvs", C.c_uint32), ("n_triangles", C.c_uint32),
                ("vertices", C.c_void_p), ("uvs", C.c_void_p), ("triangles", C.c_void_p)]


class SceneDesc(C.Structure):
    _fields_ = [("background", C.c_float * 3), ("width", C.c_uint32), ("height", C.c_uint32), ("bucket_size", C.c_uint32),
                ("camera_position", C.c_float * 3), ("camera_matrix", C.c_float * 9),
                ("n_lights", C.c_uint32), ("lights", C.POINTER(LightDesc)),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(TextureDesc)),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(MaterialDesc)),
                ("n_meshes", C.c_uint32), ("meshes", C.POINTER(MeshDesc)),
                ("n_texel_bytes", C.c_uint64), ("texels", C.c_void_p)]


class BuildOpts(C.Structure):
    _fields_ = [("kd_max_depth", C.c_uint32), ("kd_max_leaf_size", C.c_uint32), ("device", C.c_int32),
                ("accel_width", C.c_uint32), ("accel_build", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("fov_degrees", C.c_double), ("epsilon", C.c_float), ("shadow_bias", C.c_float),
                ("reflection_bias", C.c_float), ("refraction_bias", C.c_float), ("samples_per_pixel", C.c_uint32),
                ("max_ray_depth", C.c_uint32), ("diffuse_reflection_ray_count", C.c_uint32), ("seed", C.c_uint32),
                ("sample_offset", C.c_uint32), ("spp_total", C.c_uint32),
                ("x0", C.c_uint32), ("y0", C.c_uint32), ("x1", C.c_uint32), ("y1", C.c_uint32), ("flags", C.c_uint32),
                ("band_rows", C.c_uint32), ("band_period", C.c_uint32), ("band_phase", C.c_uint32)]


class SceneInfo(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32),
                ("n_triangles", C.c_uint64), ("n_vertices", C.c_uint64), ("n_nodes", C.c_uint64), ("n_leaves", C.c_uint64),
                ("n_leaf_refs", C.c_uint64), ("n_packets", C.c_uint64), ("max_leaf_refs", C.c_uint64), ("tree_depth", C.c_uint64),
                ("device_bytes", C.c_uint64), ("build_seconds", C.c_double), ("flatten_seconds", C.c_double),
                ("upload_seconds", C.c_double), ("device", C.c_int32),
                ("bvh_leaf_size", C.c_uint32),
                ("bvh_n_nodes", C.c_uint64), ("bvh_n_refs", C.c_uint64), ("bvh_n_leaves", C.c_uint64), ("bvh_depth", C.c_uint64),
                ("accel_width", C.c_uint32), ("accel_build", C.c_uint32), ("bvh4_n_nodes", C.c_uint64), ("bvh4_stack_need", C.c_uint64),
                ("accel_build_seconds", C.c_double)]


class Counters(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("primary_hits", C.c_uint64), ("shadow", C.c_uint64), ("shadow_hits", C.c_uint64),
                ("secondary", C.c_uint64), ("secondary_hits", C.c_uint64), ("nodes_pool", C.c_uint64), ("shadow_pool", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("passes", C.c_uint32), ("ms_total", C.c_float), ("ms_primary", C.c_float),
                ("ms_secondary", C.c_float), ("ms_shadow", C.c_float), ("ms_shade", C.c_float), ("ms_resolve", C.c_float)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing - build it with `python simd-raytracer_b200/build.py` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32, f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_float
    L.rt_abi_version.restype = C.c_int
    L.rt_status_string.restype = C.c_char_p
    L.rt_status_string.argtypes = [C.c_int]
    L.rt_last_error.restype = C.c_char_p
    L.rt_default_build_opts.argtypes = [C.POINTER(BuildOpts)]
    L.rt_default_params.argtypes = [C.POINTER(Params)]
    L.rt_scene_create.argtypes = [C.POINTER(SceneDesc), C.POINTER(BuildOpts), C.POINTER(vp)]
    L.rt_scene_create_from_crtscene.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(BuildOpts), C.POINTER(vp)]
    L.rt_scene_create_from_rtsc.argtypes = [vp, u64, C.POINTER(BuildOpts), C.POINTER(vp)]
    L.rt_scene_export_rtsc.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.rt_scene_destroy.argtypes = [vp]
    L.rt_scene_destroy.restype = None
    L.rt_scene_get_info.argtypes = [vp, C.POINTER(SceneInfo)]
    L.rt_scene_get_tree.argtypes = [vp, vp, vp, vp]
    L.rt_scene_get_device_layout.argtypes = [vp, vp, vp]
    L.rt_scene_get_geometry.argtypes = [vp, vp, vp, vp]
    L.rt_scene_get_bvh_layout.argtypes = [vp, vp, vp, vp]
    L.rt_trace_closest.argtypes = [vp, vp, u64, i32, f32, u32, vp]
    L.rt_trace_occluded.argtypes = [vp, vp, vp, u64, f32, f32, u32, vp]
    L.rt_trace_closest_device.argtypes = [vp, vp, u64, i32, f32, u32, vp, vp]
    L.rt_trace_occluded_device.argtypes = [vp, vp, vp, u64, f32, f32, u32, vp, vp]
    L.rt_render_frame.argtypes = [vp, C.POINTER(Params), vp]
    L.rt_render_frame_rgb8.argtypes = [vp, C.POINTER(Params), vp]
    L.rt_render_frame_device.argtypes = [vp, C.POINTER(Params), vp, vp]
    L.rt_render_frame_begin.argtypes = [vp, C.POINTER(Params), vp, C.POINTER(u64)]
    L.rt_render_frame_rgb8_begin.argtypes = [vp, C.POINTER(Params), vp, C.POINTER(u64)]
    L.rt_frame_wait.argtypes = [vp, u64]
    L.rt_alloc_pinned.argtypes = [u64]
    L.rt_alloc_pinned.restype = vp
    L.rt_free_pinned.argtypes = [vp]
    L.rt_free_pinned.restype = None
    L.rt_render_frame_device_begin.argtypes = [vp, C.POINTER(Params), vp, vp, C.POINTER(u64)]
    L.rt_trace_primary.argtypes = [vp, C.POINTER(Params), vp]
    L.rt_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.rt_resolve_sum_device.argtypes = [vp, vp, u32, vp, vp, vp]
    L.rt_peer_group_create.argtypes = [u32, u32, i32, u32, u32, C.POINTER(vp), vp]
    L.rt_peer_group_connect.argtypes = [vp, vp]
    L.rt_peer_group_connect_local.argtypes = [C.POINTER(vp), u32]
    for fn in ("rt_peer_framebuffer", "rt_peer_result_rgb", "rt_peer_result_rgb8"):
        getattr(L, fn).argtypes = [vp]
        getattr(L, fn).restype = vp
    L.rt_peer_combine.argtypes = [vp, u32, u32, vp]
    L.rt_peer_signal_ready.argtypes = [vp, vp]
    L.rt_peer_reduce_resolve.argtypes = [vp, u32, u32, vp]
    L.rt_peer_wait_done.argtypes = [vp, vp]
    L.rt_peer_read_result.argtypes = [vp, vp, vp, vp]
    L.rt_peer_download_result.argtypes = [vp, vp, vp, vp]
    L.rt_peer_host_result_attach.argtypes = [vp, C.c_char_p, i32, C.POINTER(vp)]
    L.rt_peer_group_destroy.argtypes = [vp]
    L.rt_peer_group_destroy.restype = None
    return L


lib = _load()


def _stream(stream):
    """None -> the scene's own stream (C ABI: null).  An integer is a cudaStream_t; 0 is CUDA's legacy default stream,
    which the C ABI cannot tell from "none", so it is passed as cudaStreamLegacy (0x1)."""
    if stream is None:
        return None
    return C.c_void_p(1 if int(stream) == 0 else int(stream))


def _check(status: int) -> None:
    if status != RT_OK:
        raise RtError(status, (lib.rt_last_error() or b"").decode(errors="replace"))


def default_params(**kw) -> Params:
    """config.hpp:6-17 defaults; keyword arguments override fields of rt_params."""
    p = Params()
    lib.rt_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


ACCEL_BUILD_HOST, ACCEL_BUILD_DEVICE = 0, 1


def build_opts(kd_max_depth: int = 8, kd_max_leaf_size: int = 64, device: int = 0, accel_width: int = 0,
               accel_build: int = ACCEL_BUILD_HOST) -> BuildOpts:
    o = BuildOpts()
    lib.rt_default_build_opts(C.byref(o))
    o.kd_max_depth, o.kd_max_leaf_size, o.device = kd_max_depth, kd_max_leaf_size, device
    o.accel_width, o.accel_build = accel_width, accel_build
    return o


class Scene:
    """Owns an rt_scene handle: the role of kd_tree_simd_accel (render/accel/kd_tree_simd.hpp:63-303) + its scene_ptr."""

    def __init__(self, handle: int):
        self.h = C.c_void_p(handle)
        info = SceneInfo()
        _check(lib.rt_scene_get_info(self.h, C.byref(info)))
        self.info = info
        self.width, self.height = info.width, info.height

    # ---- construction -------------------------------------------------------------------------------------------
    @classmethod
    def from_rtsc(cls, data: bytes, kd_max_depth: int = 8, kd_max_leaf_size: int = 64, device: int = 0, accel_width: int = 0,
                  accel_build: int = ACCEL_BUILD_HOST) -> "Scene":
        h = C.c_void_p()
        o = build_opts(kd_max_depth, kd_max_leaf_size, device, accel_width, accel_build)
        buf = C.create_string_buffer(data, len(data))
        _check(lib.rt_scene_create_from_rtsc(C.cast(buf, C.c_void_p), len(data), C.byref(o), C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_crtscene(cls, path: str, asset_root: str | None = None, kd_max_depth: int = 8, kd_max_leaf_size: int = 64,
                      device: int = 0) -> "Scene":
        h = C.c_void_p()
        o = build_opts(kd_max_depth, kd_max_leaf_size, device)
        _check(lib.rt_scene_create_from_crtscene(path.encode(), asset_root.encode() if asset_root else None, C.byref(o), C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_arrays(cls, *, width: int, height: int, background=(0, 0, 0), bucket_size: int = 64, camera_position=(0, 0, 0),
                    camera_matrix=(1, 0, 0, 0, 1, 0, 0, 0, 1), lights=(), textures=(), materials=(), meshes=(), texels=b"",
                    kd_max_depth: int = 8, kd_max_leaf_size: int = 64, device: int = 0) -> "Scene":
        """meshes: iterable of (material_index, vertices (nv,3) f32, uvs (nuv,2) f32 or None, triangles (nt,3) u32);
        lights: (x, y, z, intensity); materials: dicts with kind/albedo/ior/smooth_shading/texture;
        textures: dicts with kind/c0/c1/scalar/bmp_w/bmp_h/bmp_off."""
        d = SceneDesc()
        d.background[:] = background
        d.width, d.height, d.bucket_size = width, height, bucket_size
        d.camera_position[:] = camera_position
        d.camera_matrix[:] = camera_matrix
        keep = []
        larr = (LightDesc * max(len(lights), 1))()
        for i, l in enumerate(lights):
            larr[i].position[:] = l[:3]
            larr[i].intensity = l[3]
        tarr = (TextureDesc * max(len(textures), 1))()
        for i, t in enumerate(textures):
            tarr[i].kind = t["kind"]
            tarr[i].c0[:] = t.get("c0", (0, 0, 0))
            tarr[i].c1[:] = t.get("c1", (0, 0, 0))
            tarr[i].scalar = t.get("scalar", 0.0)
            tarr[i].bmp_w, tarr[i].bmp_h, tarr[i].bmp_off = t.get("bmp_w", 0), t.get("bmp_h", 0), t.get("bmp_off", 0)
        marr = (MaterialDesc * max(len(materials), 1))()
        for i, m in enumerate(materials):
            marr[i].kind = m["kind"]
            marr[i].albedo[:] = m.get("albedo", (0, 0, 0))
            marr[i].ior = m.get("ior", 1.0)
            marr[i].smooth_shading = int(m.get("smooth_shading", 0))
            marr[i].texture = m.get("texture", -1)
        meshes = list(meshes)
        xarr = (MeshDesc * max(len(meshes), 1))()
        for i, (mat, v, uv, tri) in enumerate(meshes):
            v = np.ascontiguousarray(v, np.float32).reshape(-1, 3)
            tri = np.ascontiguousarray(tri, np.uint32).reshape(-1, 3)
            uv = np.zeros((0, 2), np.float32) if uv is None else np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
            keep += [v, uv, tri]
            xarr[i].material, xarr[i].n_vertices, xarr[i].n_uvs, xarr[i].n_triangles = mat, len(v), len(uv), len(tri)
            xarr[i].vertices, xarr[i].uvs, xarr[i].triangles = v.ctypes.data, (uv.ctypes.data if len(uv) else None), tri.ctypes.data
        tex = np.frombuffer(bytes(texels), np.uint8)
        d.n_lights, d.lights = len(lights), larr
        d.n_textures, d.textures = len(textures), tarr
        d.n_materials, d.materials = len(materials), marr
        d.n_meshes, d.meshes = len(meshes), xarr
        d.n_texel_bytes, d.texels = len(tex), (tex.ctypes.data if len(tex) else None)
        h = C.c_void_p()
        o = build_opts(kd_max_depth, kd_max_leaf_size, device)
        _check(lib.rt_scene_create(C.byref(d), C.byref(o), C.byref(h)))
        del keep
        return cls(h.value)

    def close(self) -> None:
        if getattr(self, "h", None) is not None and self.h.value:
            lib.rt_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- host-side structures (no GPU needed) -----------------------------------------------------------------------
    def export_rtsc(self) -> bytes:
        """the scene as loaded, in the flat RTSC container (rt_scene_export_rtsc)"""
        n = C.c_uint64(0)
        _check(lib.rt_scene_export_rtsc(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _check(lib.rt_scene_export_rtsc(self.h, C.cast(buf, C.c_void_p), n.value, C.byref(n)))
        return buf.raw[:n.value]

    def tree(self):
        n, r = self.info.n_nodes, self.info.n_leaf_refs
        node5 = np.zeros((n, 5), np.uint64)
        boxes = np.zeros((n, 6), np.float32)
        refs = np.zeros(r, np.uint32)
        _check(lib.rt_scene_get_tree(self.h, node5.ctypes.data, boxes.ctypes.data, refs.ctypes.data))
        return node5, boxes, refs

    def device_layout(self):
        nodes8 = np.zeros((self.info.n_nodes, 2), np.uint32)
        packets = np.zeros((self.info.n_packets, 10, 4), np.uint32)
        _check(lib.rt_scene_get_device_layout(self.h, nodes8.ctypes.data, packets.ctypes.data))
        return nodes8, packets

    def bvh_layout(self):
        """64-byte two-child nodes / 48-byte triangle records / root box of the bounding-volume hierarchy (RT_FLAG_ORDERED)"""
        nodes = np.zeros((max(self.info.bvh_n_nodes, 1), 16), np.uint32)
        tris = np.zeros((max(self.info.bvh_n_refs, 1), 12), np.uint32)
        root = np.zeros(6, np.float32)
        _check(lib.rt_scene_get_bvh_layout(self.h, nodes.ctypes.data, tris.ctypes.data, root.ctypes.data))
        return nodes, tris, root

    def geometry(self):
        tri9 = np.zeros((self.info.n_triangles, 9), np.float32)
        fn = np.zeros((self.info.n_triangles, 3), np.float32)
        vn = np.zeros((self.info.n_vertices, 3), np.float32)
        _check(lib.rt_scene_get_geometry(self.h, tri9.ctypes.data, fn.ctypes.data, vn.ctypes.data))
        return tri9, fn, vn

    # ---- queries: accel.intersect<cull>(ray) / is_occluded for batches -----------------------------------------------
    def trace_closest(self, rays, cull: bool, eps: float = np.float32(1e-6), flags: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        hits = np.zeros(len(rays), HIT_DTYPE)
        _check(lib.rt_trace_closest(self.h, rays.ctypes.data, len(rays), 1 if cull else 0, C.c_float(eps), flags, hits.ctypes.data))
        return hits

    def trace_occluded(self, rays, max_t, eps: float = np.float32(1e-6), shadow_bias: float = np.float32(1e-4),
                       flags: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        max_t = np.ascontiguousarray(max_t, np.float32).reshape(-1)
        if len(max_t) != len(rays):
            raise ValueError("max_t length")
        out = np.zeros(len(rays), np.uint8)
        _check(lib.rt_trace_occluded(self.h, rays.ctypes.data, max_t.ctypes.data, len(rays), C.c_float(eps), C.c_float(shadow_bias),
                                     flags, out.ctypes.data))
        return out

    def trace_closest_device(self, d_rays: int, n: int, cull: bool, d_hits: int, eps: float = np.float32(1e-6), flags: int = 0,
                             stream: int | None = None) -> None:
        _check(lib.rt_trace_closest_device(self.h, d_rays, n, 1 if cull else 0, C.c_float(eps), flags, d_hits, _stream(stream)))

    def trace_occluded_device(self, d_rays: int, d_max_t: int, n: int, d_out: int, eps: float = np.float32(1e-6),
                              shadow_bias: float = np.float32(1e-4), flags: int = 0, stream: int | None = None) -> None:
        _check(lib.rt_trace_occluded_device(self.h, d_rays, d_max_t, n, C.c_float(eps), C.c_float(shadow_bias), flags, d_out,
                                            _stream(stream)))

    def trace_primary(self, params: Params | None = None) -> np.ndarray:
        p = params or default_params()
        x1, y1 = (p.x1 or self.width), (p.y1 or self.height)
        hits = np.zeros((y1 - p.y0, x1 - p.x0), HIT_DTYPE)
        _check(lib.rt_trace_primary(self.h, C.byref(p), hits.ctypes.data))
        return hits

    # ---- frames: render_frame<A,F> ----------------------------------------------------------------------------------------
    def render_frame(self, params: Params | None = None, out: np.ndarray | None = None) -> np.ndarray:
        p = params or default_params()
        img = out if out is not None else np.zeros((self.height, self.width, 3), np.float32)
        assert img.dtype == np.float32 and img.flags.c_contiguous and img.shape == (self.height, self.width, 3)
        _check(lib.rt_render_frame(self.h, C.byref(p), img.ctypes.data))
        return img

    def render_frame_rgb8(self, params: Params | None = None, out: np.ndarray | None = None) -> np.ndarray:
        p = params or default_params()
        img = out if out is not None else np.zeros((self.height, self.width, 3), np.uint8)
        _check(lib.rt_render_frame_rgb8(self.h, C.byref(p), img.ctypes.data))
        return img

    def render_frame_begin(self, params: Params, out: np.ndarray) -> int:
        """Frame sequences (rt_render_frame_begin): renders now, downloads into `out` (pinned host memory) behind the next
        frame's render; returns the ticket for frame_wait."""
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (self.height, self.width, 3)
        t = C.c_uint64(0)
        _check(lib.rt_render_frame_begin(self.h, C.byref(params), out.ctypes.data, C.byref(t)))
        return int(t.value)

    def render_frame_rgb8_begin(self, params: Params, out: np.ndarray) -> int:
        """rt_render_frame_rgb8_begin: the frame sequence with the 8-bit frame (io/image/ppm.hpp:17-19 on the device) downloaded"""
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape == (self.height, self.width, 3)
        t = C.c_uint64(0)
        _check(lib.rt_render_frame_rgb8_begin(self.h, C.byref(params), out.ctypes.data, C.byref(t)))
        return int(t.value)

    def frame_wait(self, ticket: int) -> bool:
        """True when the frame had to be rendered again (only reported for render_frame_device_begin tickets)"""
        st = lib.rt_frame_wait(self.h, ticket)
        if st == RT_FRAME_RERENDERED:
            return True
        _check(st)
        return False

    def render_frame_device_begin(self, params: Params, d_rgb: int, stream: int | None = None) -> int:
        t = C.c_uint64(0)
        _check(lib.rt_render_frame_device_begin(self.h, C.byref(params), d_rgb, _stream(stream), C.byref(t)))
        return int(t.value)

    def render_frame_device(self, params: Params, d_rgb: int, stream: int | None = None) -> None:
        _check(lib.rt_render_frame_device(self.h, C.byref(params), d_rgb, _stream(stream)))

    def resolve_sum_device(self, d_sum: int, spp_total: int, d_rgb: int = 0, d_rgb8: int = 0, stream: int | None = None) -> None:
        _check(lib.rt_resolve_sum_device(self.h, d_sum, spp_total, d_rgb or None, d_rgb8 or None, _stream(stream)))

    def counters(self) -> Counters:
        c = Counters()
        _check(lib.rt_get_counters(self.h, C.byref(c)))
        return c


PEER_HANDLE_BYTES = 64
PEER_OUT_RGB, PEER_OUT_RGB8, PEER_OUT_HOST_RGB = 1, 2, 4


class PeerGroup:
    """One rank's end of the NVLink peer-memory combine (include/rt_b200.h "multi-GPU combine"): owns the rank's raw-sum
    framebuffer, result buffers and flag block; `handle` is the cudaIpc handle to all-gather across ranks."""

    def __init__(self, world: int, rank: int, device: int, width: int, height: int):
        self.world, self.rank, self.width, self.height = world, rank, width, height
        h = C.c_void_p()
        buf = (C.c_uint8 * PEER_HANDLE_BYTES)()
        _check(lib.rt_peer_group_create(world, rank, device, width, height, C.byref(h), C.cast(buf, C.c_void_p)))
        self.h = h
        self.handle = bytes(buf)

    def connect(self, handles) -> None:
        """handles: the `handle` of every rank, in rank order (multi-process: one rank per process)."""
        blob = b"".join(handles)
        assert len(blob) == self.world * PEER_HANDLE_BYTES
        _check(lib.rt_peer_group_connect(self.h, C.cast(C.create_string_buffer(blob, len(blob)), C.c_void_p)))

    @staticmethod
    def connect_local(groups) -> None:
        """all ranks live in this process (single-GPU emulation of N ranks, or N peer GPUs driven by one thread)"""
        arr = (C.c_void_p * len(groups))(*[g.h for g in groups])
        _check(lib.rt_peer_group_connect_local(arr, len(groups)))

    @property
    def framebuffer(self) -> int:
        return int(lib.rt_peer_framebuffer(self.h))

    @property
    def result_rgb(self) -> int:
        return int(lib.rt_peer_result_rgb(self.h))

    @property
    def result_rgb8(self) -> int:
        return int(lib.rt_peer_result_rgb8(self.h))

    def combine(self, spp_total: int, outputs: int = PEER_OUT_RGB | PEER_OUT_RGB8, stream: int | None = None) -> None:
        _check(lib.rt_peer_combine(self.h, spp_total, outputs, _stream(stream)))

    def signal_ready(self, stream: int | None = None) -> None:
        _check(lib.rt_peer_signal_ready(self.h, _stream(stream)))

    def reduce_resolve(self, spp_total: int, outputs: int = PEER_OUT_RGB | PEER_OUT_RGB8, stream: int | None = None) -> None:
        _check(lib.rt_peer_reduce_resolve(self.h, spp_total, outputs, _stream(stream)))

    def wait_done(self, stream: int | None = None) -> None:
        _check(lib.rt_peer_wait_done(self.h, _stream(stream)))

    def read_result(self, stream: int | None = None, rgb: np.ndarray | None = None, want_rgb8: bool = True):
        """rank 0: (float frame, 8-bit frame) on the host; `rgb` may be a caller-provided (pinned) float32 buffer"""
        if rgb is None:
            rgb = np.zeros((self.height, self.width, 3), np.float32)
        assert rgb.dtype == np.float32 and rgb.size == self.height * self.width * 3 and rgb.flags["C_CONTIGUOUS"]
        rgb8 = np.zeros((self.height, self.width, 3), np.uint8) if want_rgb8 else None
        _check(lib.rt_peer_read_result(self.h, rgb.ctypes.data, rgb8.ctypes.data if want_rgb8 else None, _stream(stream)))
        return rgb, rgb8

    def download_result(self, rgb: np.ndarray, stream: int | None = None) -> None:
        """rank 0: queue the copy of the combined float frame (of the frame signalled last) into pinned `rgb`; no sync"""
        assert rgb.dtype == np.float32 and rgb.size == self.height * self.width * 3 and rgb.flags["C_CONTIGUOUS"]
        _check(lib.rt_peer_download_result(self.h, rgb.ctypes.data, None, _stream(stream)))

    def attach_host_result(self, shm_name: str, create: bool) -> np.ndarray:
        """map and pin the shared host frame (two slots of (H, W, 3) float32); frame e lands in slot (e - 1) & 1 when
        reduce_resolve is called with PEER_OUT_HOST_RGB.  The array is a view of the mapping: drop it before close()."""
        p = C.c_void_p()
        _check(lib.rt_peer_host_result_attach(self.h, shm_name.encode(), int(create), C.byref(p)))
        n = 2 * self.height * self.width * 3
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).reshape(2, self.height, self.width, 3)

    def close(self) -> None:
        if self.h:
            lib.rt_peer_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


# ---- multi-GPU partitioning: pure functions, shared by bench.py and the CLI ------------------------------
def spp_slice(spp_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous sample slice [first, first+count) of rank `rank`; slices differ by at most one sample."""
    base, rem = divmod(spp_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def row_bands(height: int, bucket: int, rank: int, world: int) -> list[tuple[int, int]]:
    """Interleaved horizontal bands of `bucket` rows (the reference's bucket decomposition, render/tile/bucket.hpp:7-21,
    dealt round-robin): [(y0, y1), ...] of rank `rank`."""
    out = []
    for k, y0 in enumerate(range(0, height, bucket)):
        if k % world == rank:
            out.append((y0, min(height, y0 + bucket)))
    return out


def sharded_frame(render_slice, spp_total: int, rank: int, world: int, reduce_to_root, resolve):
    """One frame over `world` ranks with a replicated scene: every rank renders the RAW SAMPLE SUM of its
    own sample slice, ONE reduce(sum) brings the framebuffers to rank 0, which divides by spp_total (and quantises).

    render_slice(first_sample, n_samples) -> framebuffer (whatever tensor type the caller's collective takes)
    reduce_to_root(fb) -> None            in-place sum-reduction to rank 0 (torch.distributed.reduce over NCCL / gloo)
    resolve(fb) -> frame                  rank 0 only: sum / spp_total
    A rank whose slice is empty (world > spp_total) contributes zeros."""
    first, count = spp_slice(spp_total, rank, world)
    fb = render_slice(first, count)
    if world > 1:
        reduce_to_root(fb)
    return resolve(fb) if rank == 0 else None
