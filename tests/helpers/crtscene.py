"""Test-infrastructure scene loader: .crtscene JSON -> flat numpy arrays, and the RTSC flat binary format.

This is NOT the product loader (that one is C++: simd-raytracer_b200/host/crtscene_loader.cpp).  It exists so the
oracle (oracle/rt_oracle.c), the compiled reference (oracle/_ref) and the product can all be fed byte-identical
scene data in tests.  Semantics follow the reference loader, /root/reference/include/raytracer/io/json/loader.hpp:
  * every number is float(double_from_json)                               (loader.hpp:9-17)
  * bucket_size optional, default 64                                      (loader.hpp:46-60)
  * uvs are 3-component in the file, third component dropped              (loader.hpp:174-193)
  * material "diffuse" with a string albedo becomes a texture material    (loader.hpp:108-128)
  * textures live in a map keyed by name (first insertion wins)           (loader.hpp:250-254)
  * mesh_idx of every triangle = position of the object in the array      (loader.hpp:260-262)

RTSC layout (little endian), version 1:
  "RTSC" u32 version
  f32 bg[3]; u32 width, height, bucket
  f32 cam_pos[3]; f32 cam_matrix[9]
  u32 n_lights;    n_lights  x { f32 pos[3], f32 intensity }
  u32 n_textures;  n_textures x { u32 kind, f32 c0[3], f32 c1[3], f32 scalar, u32 bmp_w, u32 bmp_h, u32 bmp_off }
  u32 n_materials; n_materials x { u32 kind, f32 albedo[3], f32 ior, u32 smooth, i32 texture }
  u32 n_meshes;    n_meshes x { u32 material, u32 n_vertices, u32 n_uvs, u32 n_tris }
  then per mesh: f32 vertices[3*nv]; f32 uvs[2*nuv]; u32 tris[3*nt]
  u32 n_texel_bytes; u8 texels[n] (RGB8, row major, all bitmaps concatenated)
"""
from __future__ import annotations

import gzip
import json
import os
import struct
from dataclasses import dataclass, field

import numpy as np

TEX_ALBEDO, TEX_EDGES, TEX_CHECKER, TEX_BITMAP = 0, 1, 2, 3
MAT_DIFFUSE, MAT_REFLECTIVE, MAT_REFRACTIVE, MAT_CONSTANT, MAT_TEXTURE = 0, 1, 2, 3, 4

TEXTURE_DTYPE = np.dtype([("kind", "<u4"), ("c0", "<f4", 3), ("c1", "<f4", 3), ("scalar", "<f4"),
                          ("bmp_w", "<u4"), ("bmp_h", "<u4"), ("bmp_off", "<u4")])
MATERIAL_DTYPE = np.dtype([("kind", "<u4"), ("albedo", "<f4", 3), ("ior", "<f4"), ("smooth", "<u4"),
                           ("texture", "<i4")])
LIGHT_DTYPE = np.dtype([("pos", "<f4", 3), ("intensity", "<f4")])


@dataclass
class Mesh:
    material: int
    vertices: np.ndarray  # (nv,3) f32
    uvs: np.ndarray       # (nuv,2) f32
    tris: np.ndarray      # (nt,3) u32


@dataclass
class Scene:
    bg: np.ndarray
    width: int
    height: int
    bucket: int
    cam_pos: np.ndarray
    cam_matrix: np.ndarray
    lights: np.ndarray
    textures: np.ndarray
    materials: np.ndarray
    meshes: list = field(default_factory=list)
    texels: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint8))

    @property
    def n_triangles(self) -> int:
        return int(sum(len(m.tris) for m in self.meshes))


def _f32(x):
    return np.asarray(x, dtype=np.float64).astype(np.float32)


def decode_bitmap(path: str) -> np.ndarray:
    """Decode an image file to (h,w,3) uint8.

    The reference decodes with stb_image (scene/texture/bitmap.hpp:15), which is not in this image.  Baseline 4:4:4 JPEG - the
    one bitmap the reference ships - goes through tests/helpers/jpeg_stb.py, a restatement of the fixed-point arithmetic of
    that decoder's JPEG path; with its texels the bitmap quadrant of the reference's published outputs/textures.png is
    reproduced exactly (tests/test_oracle_golden.py), which is the pin.  Any other file falls back to PIL (libjpeg differs from
    that arithmetic by up to 3/255 on the shipped file), and is then 'parity unpinned' at the texel level.
    """
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:2] == b"\xff\xd8":
        from . import jpeg_stb
        try:
            return jpeg_stb.decode(data)
        except NotImplementedError:
            pass
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("RGB"), dtype=np.uint8)


def load_crtscene(path: str, root: str | None = None) -> Scene:
    with open(path, "r") as fh:
        doc = json.load(fh)
    st = doc["settings"]
    ims = st["image_settings"]
    bucket = int(ims.get("bucket_size", 64))
    cam = doc["camera"]
    lights = np.zeros(len(doc["lights"]), LIGHT_DTYPE)
    for i, l in enumerate(doc["lights"]):
        lights[i]["pos"] = _f32(l["position"])
        lights[i]["intensity"] = _f32(l["intensity"])

    tex_names: dict[str, int] = {}
    tex_list = []
    texel_chunks = []
    texel_off = 0
    for t in doc.get("textures", []):
        rec = np.zeros((), TEXTURE_DTYPE)
        kind = t["type"]
        if kind == "albedo":
            rec["kind"] = TEX_ALBEDO
            rec["c0"] = _f32(t["albedo"])
        elif kind == "edges":
            rec["kind"] = TEX_EDGES
            rec["c0"] = _f32(t["edge_color"])
            rec["c1"] = _f32(t["inner_color"])
            rec["scalar"] = _f32(t["edge_width"])
        elif kind == "checker":
            rec["kind"] = TEX_CHECKER
            rec["c0"] = _f32(t["color_A"])
            rec["c1"] = _f32(t["color_B"])
            rec["scalar"] = _f32(t["square_size"])
        elif kind == "bitmap":
            rec["kind"] = TEX_BITMAP
            fp = t["file_path"]
            if root is not None and not os.path.isabs(fp):
                fp = os.path.join(root, fp)
            img = decode_bitmap(fp)
            rec["bmp_h"], rec["bmp_w"] = img.shape[0], img.shape[1]
            rec["bmp_off"] = texel_off
            texel_chunks.append(img.reshape(-1))
            texel_off += img.size
        else:
            raise ValueError("texture type unknown")
        if t["name"] not in tex_names:  # unordered_map::emplace keeps the first
            tex_names[t["name"]] = len(tex_list)
        tex_list.append(rec)
    textures = np.array(tex_list, TEXTURE_DTYPE) if tex_list else np.zeros(0, TEXTURE_DTYPE)

    materials = np.zeros(len(doc["materials"]), MATERIAL_DTYPE)
    for i, m in enumerate(doc["materials"]):
        kind = m["type"]
        materials[i]["texture"] = -1
        materials[i]["ior"] = 1.0
        if kind == "diffuse":
            if isinstance(m["albedo"], list):
                materials[i]["kind"] = MAT_DIFFUSE
                materials[i]["albedo"] = _f32(m["albedo"])
            elif isinstance(m["albedo"], str):
                materials[i]["kind"] = MAT_TEXTURE
                materials[i]["texture"] = tex_names[m["albedo"]]
            else:
                raise ValueError("albedo neither array nor string")
        elif kind == "reflective":
            materials[i]["kind"] = MAT_REFLECTIVE
            materials[i]["albedo"] = _f32(m["albedo"])
        elif kind == "refractive":
            materials[i]["kind"] = MAT_REFRACTIVE
            materials[i]["ior"] = _f32(m["ior"])
        elif kind == "constant":
            materials[i]["kind"] = MAT_CONSTANT
            materials[i]["albedo"] = _f32(m["albedo"])
        else:
            raise ValueError("material type unknown")
        materials[i]["smooth"] = 1 if m["smooth_shading"] else 0

    meshes = []
    for o in doc["objects"]:
        v = _f32(o["vertices"])
        if v.size % 3:
            raise ValueError("vertex coordinates not multiple of 3")
        uv = _f32(o.get("uvs", []))
        if uv.size % 3:
            raise ValueError("uv coordinates not multiple of 3")
        tr = np.asarray(o["triangles"], dtype=np.uint32)
        if tr.size % 3:
            raise ValueError("triangle indices not multiple of 3")
        meshes.append(Mesh(int(o["material_index"]), v.reshape(-1, 3).copy(),
                           uv.reshape(-1, 3)[:, :2].copy() if uv.size else np.zeros((0, 2), np.float32),
                           tr.reshape(-1, 3).copy()))

    return Scene(bg=_f32(st["background_color"]), width=int(ims["width"]), height=int(ims["height"]), bucket=bucket,
                 cam_pos=_f32(cam["position"]), cam_matrix=_f32(cam["matrix"]), lights=lights, textures=textures,
                 materials=materials, meshes=meshes,
                 texels=np.concatenate(texel_chunks) if texel_chunks else np.zeros(0, np.uint8))


def to_rtsc_bytes(s: Scene) -> bytes:
    out = [b"RTSC", struct.pack("<I", 1)]
    out.append(np.asarray(s.bg, "<f4").tobytes())
    out.append(struct.pack("<III", s.width, s.height, s.bucket))
    out.append(np.asarray(s.cam_pos, "<f4").tobytes())
    out.append(np.asarray(s.cam_matrix, "<f4").tobytes())
    out.append(struct.pack("<I", len(s.lights)))
    out.append(s.lights.tobytes())
    out.append(struct.pack("<I", len(s.textures)))
    out.append(s.textures.tobytes())
    out.append(struct.pack("<I", len(s.materials)))
    out.append(s.materials.tobytes())
    out.append(struct.pack("<I", len(s.meshes)))
    for m in s.meshes:
        out.append(struct.pack("<IIII", m.material, len(m.vertices), len(m.uvs), len(m.tris)))
    for m in s.meshes:
        out.append(np.ascontiguousarray(m.vertices, "<f4").tobytes())
        out.append(np.ascontiguousarray(m.uvs, "<f4").tobytes())
        out.append(np.ascontiguousarray(m.tris, "<u4").tobytes())
    out.append(struct.pack("<I", int(s.texels.size)))
    out.append(np.ascontiguousarray(s.texels, np.uint8).tobytes())
    return b"".join(out)


def from_rtsc_bytes(buf: bytes) -> Scene:
    assert buf[:4] == b"RTSC", "bad magic"
    off = 4
    (ver,) = struct.unpack_from("<I", buf, off)
    off += 4
    assert ver == 1

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=off).copy()
        off += a.nbytes
        return a

    bg = take("<f4", 3)
    width, height, bucket = (int(x) for x in take("<u4", 3))
    cam_pos = take("<f4", 3)
    cam_matrix = take("<f4", 9)
    lights = take(LIGHT_DTYPE, int(take("<u4", 1)[0]))
    textures = take(TEXTURE_DTYPE, int(take("<u4", 1)[0]))
    materials = take(MATERIAL_DTYPE, int(take("<u4", 1)[0]))
    n_meshes = int(take("<u4", 1)[0])
    heads = [take("<u4", 4) for _ in range(n_meshes)]
    meshes = []
    for h in heads:
        v = take("<f4", 3 * int(h[1])).reshape(-1, 3)
        uv = take("<f4", 2 * int(h[2])).reshape(-1, 2)
        tr = take("<u4", 3 * int(h[3])).reshape(-1, 3)
        meshes.append(Mesh(int(h[0]), v, uv, tr))
    texels = take(np.uint8, int(take("<u4", 1)[0]))
    assert off == len(buf), "trailing bytes in RTSC"
    return Scene(bg, width, height, bucket, cam_pos, cam_matrix, lights, textures, materials, meshes, texels)


def save_rtsc(s: Scene, path: str) -> None:
    data = to_rtsc_bytes(s)
    if path.endswith(".gz"):
        with open(path, "wb") as raw:  # mtime=0 keeps the fixture bytes reproducible
            with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, compresslevel=9) as fh:
                fh.write(data)
    else:
        with open(path, "wb") as fh:
            fh.write(data)


def load_rtsc(path: str) -> Scene:
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as fh:
        return from_rtsc_bytes(fh.read())


# ---------------------------------------------------------------------------------------------------------------
# synthetic scenes (no reference data needed)
# ---------------------------------------------------------------------------------------------------------------

def synthetic_scene(n_tris: int = 2000, seed: int = 1234, width: int = 256, height: int = 192,
                    with_box: bool = True) -> Scene:
    """Random-triangle soup in [-1,1]^3 inside a diffuse box, in the spirit of SURVEY.md section 8d config 5.

    centre ~ U(-1,1)^3, each vertex = centre + e*U(-1,1)^3 with e = 2*cbrt(1/n).  numpy's PCG64 replaces the
    survey's mt19937_64 (the scene is synthetic on both sides of every comparison, only determinism matters).
    """
    rng = np.random.default_rng(seed)
    e = 2.0 * (1.0 / max(n_tris, 1)) ** (1.0 / 3.0)
    c = rng.uniform(-1, 1, (n_tris, 1, 3))
    v = (c + e * rng.uniform(-1, 1, (n_tris, 3, 3))).astype(np.float32).reshape(-1, 3)
    tris = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    meshes = [Mesh(0, v, np.zeros((0, 2), np.float32), tris)]
    mats = np.zeros(7, MATERIAL_DTYPE)
    mats["texture"] = -1
    mats["ior"] = 1.0
    mats[0]["kind"] = MAT_DIFFUSE
    mats[0]["albedo"] = (0.8, 0.8, 0.8)
    mats[0]["smooth"] = 0
    cols = [(0.9, 0.4, 0.2), (0.4, 0.9, 0.2), (0.2, 0.4, 0.9), (0.9, 0.2, 0.9), (0.2, 0.9, 0.9), (0.7, 0.7, 0.7)]
    for i, col in enumerate(cols):
        mats[1 + i]["kind"] = MAT_DIFFUSE
        mats[1 + i]["albedo"] = col
    if with_box:
        b = 1.5
        # five walls (open towards the camera at +z), two triangles each, facing inwards
        quads = [
            ([-b, -b, -b], [b, -b, -b], [b, b, -b], [-b, b, -b]),   # back  (z=-b)
            ([-b, -b, b], [-b, -b, -b], [-b, b, -b], [-b, b, b]),   # left  (x=-b)
            ([b, -b, -b], [b, -b, b], [b, b, b], [b, b, -b]),       # right (x=+b)
            ([-b, -b, b], [b, -b, b], [b, -b, -b], [-b, -b, -b]),   # floor (y=-b)
            ([-b, b, -b], [b, b, -b], [b, b, b], [-b, b, b]),       # ceil  (y=+b)
        ]
        for i, q in enumerate(quads):
            meshes.append(Mesh(1 + i, np.asarray(q, np.float32), np.zeros((0, 2), np.float32),
                               np.asarray([[0, 1, 2], [0, 2, 3]], np.uint32)))
    lights = np.zeros(2, LIGHT_DTYPE)
    lights[0]["pos"] = (0.5, 1.2, 1.4)
    lights[0]["intensity"] = 40
    lights[1]["pos"] = (-1.0, 0.3, 1.2)
    lights[1]["intensity"] = 25
    return Scene(bg=np.asarray([0.05, 0.1, 0.2], np.float32), width=width, height=height, bucket=64,
                 cam_pos=np.asarray([0, 0, 3], np.float32),
                 cam_matrix=np.asarray([1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32), lights=lights,
                 textures=np.zeros(0, TEXTURE_DTYPE), materials=mats, meshes=meshes)
