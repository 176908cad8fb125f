"""Host side of the backend without a GPU: ABI surface, struct layout, builder / flattener parity with the oracle, scene
loaders, error behaviour, sharding arithmetic.  No compute entry point runs here - they must refuse (no CPU fallback)."""
from __future__ import annotations

import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

from .conftest import REPO, SCENES, scene_bytes
from .helpers import crtscene

HEADER = os.path.join(REPO, "include", "rt_b200.h")


def header_functions() -> list[str]:
    with open(HEADER) as fh:
        return re.findall(r"^RT_API\s+[\w\s\*]+?\b(rt_[a-z0-9_]+)\(", fh.read(), flags=re.M)


def test_library_exports_every_declared_symbol(rt):
    names = header_functions()
    assert len(names) >= 20
    assert sorted(names) == sorted(rt.ABI_SYMBOLS)
    for n in names:
        assert hasattr(rt.lib, n), n
    assert rt.lib.rt_abi_version() == 3
    assert rt.lib.rt_status_string(2) == b"no usable sm_100 CUDA device"


def test_header_is_plain_c_and_matches_ctypes_layout(rt, tmp_path):
    """include/rt_b200.h compiles as C99 and every struct has the size the Python binding assumes."""
    structs = {"rt_light_desc": rt.LightDesc, "rt_texture_desc": rt.TextureDesc, "rt_material_desc": rt.MaterialDesc,
               "rt_mesh_desc": rt.MeshDesc, "rt_scene_desc": rt.SceneDesc, "rt_build_opts": rt.BuildOpts, "rt_params": rt.Params,
               "rt_scene_info": rt.SceneInfo, "rt_counters": rt.Counters}
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "rt_b200.h"\nint main(void){\n' +
                   "".join(f'printf("{n} %zu\\n", sizeof({n}));\n' for n in structs) +
                   'printf("rt_hit %zu\\n", sizeof(rt_hit));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)])
    sizes = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n, cls in structs.items():
        assert int(sizes[n]) == C.sizeof(cls), n
    assert int(sizes["rt_hit"]) == rt.HIT_DTYPE.itemsize == 16


def test_default_params_are_config_hpp(rt):
    p = rt.default_params()
    assert (p.fov_degrees, p.samples_per_pixel, p.max_ray_depth, p.diffuse_reflection_ray_count, p.seed) == (90.0, 1, 5, 0, 42)
    assert p.epsilon == np.float32(1e-6) and p.shadow_bias == p.reflection_bias == p.refraction_bias == np.float32(1e-4)
    o = rt.build_opts()
    assert (o.kd_max_depth, o.kd_max_leaf_size) == (8, 64)


@pytest.mark.parametrize("name", SCENES)
def test_builder_equals_oracle(rt, oracle_mod, golden, name):
    """same tree (node links, boxes, leaf lists), same derived geometry, bit for bit; sizes == the compiled reference's."""
    data = scene_bytes(name)
    s = rt.Scene.from_rtsc(data, device=rt.DEVICE_HOST_ONLY)
    o = oracle_mod.Oracle(data)
    assert s.info.n_nodes == o.n_nodes == golden["scenes"][name]["configs"]["s1d5g0"]["n_nodes"]
    assert (s.info.n_leaves, s.info.n_leaf_refs, s.info.max_leaf_refs, s.info.tree_depth) == (o.n_leaves, o.n_refs, o.max_leaf_refs, o.tree_depth)
    n5, bx, rf = s.tree()
    on5, obx, orf = o.tree()
    assert np.array_equal(n5, on5) and np.array_equal(bx.view(np.uint32), obx.view(np.uint32)) and np.array_equal(rf, orf)
    t9, fn, vn = s.geometry()
    ot9, ofn, ovn, _, _ = o.geometry(s.info.n_vertices)
    for a, b in ((t9, ot9), (fn, ofn), (vn, ovn)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # W=8 pack count of the reference (kd_tree_simd.hpp:117-144) follows from the leaf sizes
    leaf_sizes = n5[n5[:, 3] != np.uint64(2**64 - 1), 4].astype(np.int64)
    assert int(((leaf_sizes + 7) // 8).sum()) == golden["scenes"][name]["configs"]["s1d5g0"]["n_packs_w8"]


@pytest.mark.parametrize("kd", [(8, 64), (24, 64), (12, 4), (0, 64), (14, 2)])
def test_builder_equals_oracle_synthetic(rt, oracle_mod, kd):
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=3000, seed=7))
    s = rt.Scene.from_rtsc(data, kd_max_depth=kd[0], kd_max_leaf_size=kd[1], device=rt.DEVICE_HOST_ONLY)
    o = oracle_mod.Oracle(data, kd[0], kd[1])
    n5, bx, rf = s.tree()
    on5, obx, orf = o.tree()
    assert np.array_equal(n5, on5) and np.array_equal(bx.view(np.uint32), obx.view(np.uint32)) and np.array_equal(rf, orf)


@pytest.mark.parametrize("name", ["hw09_scene5", "hw12_scene4"])
def test_device_layout_encodes_the_tree(rt, name):
    """8-byte nodes and 4-wide SoA packets decode back to the host tree; padding lanes repeat the leaf's last triangle."""
    s = rt.Scene.from_rtsc(scene_bytes(name), device=rt.DEVICE_HOST_ONLY)
    n5, bx, rf = s.tree()
    nodes8, packets = s.device_layout()
    t9, _, _ = s.geometry()
    NONE = np.uint64(2**64 - 1)
    next_packet = 0
    for i in range(len(n5)):
        first, word = int(nodes8[i, 0]), int(nodes8[i, 1])
        if n5[i, 3] == NONE:                                    # inner
            axis = word & 3
            assert axis < 3
            assert bool(word & 4) == (n5[i, 1] != NONE) and bool(word & 8) == (n5[i, 2] != NONE)
            if word & 4:
                assert int(n5[i, 1]) == i + 1
                split = np.uint32(first).view(np.float32)
                assert split == bx[i + 1, 3 + axis]             # child0.max[axis] == split plane
            if word & 8:
                assert (word >> 4) == int(n5[i, 2])
        else:
            assert (word & 3) == 3 and first == next_packet
            count, n = word >> 2, int(n5[i, 4])
            assert count == (n + 3) // 4
            ids = packets[first:first + count, 9, :].reshape(-1)
            want = rf[int(n5[i, 3]):int(n5[i, 3]) + n]
            assert np.array_equal(ids[:n], want) and np.all(ids[n:] == want[-1])
            soa = packets[first:first + count, :9, :].view(np.float32).transpose(0, 2, 1).reshape(-1, 9)
            assert np.array_equal(soa.view(np.uint32), t9[ids].view(np.uint32))
            next_packet += count
    assert next_packet == s.info.n_packets


def test_scene_from_arrays_equals_rtsc(rt):
    sc = crtscene.synthetic_scene(n_tris=500, seed=3)
    a = rt.Scene.from_rtsc(crtscene.to_rtsc_bytes(sc), device=rt.DEVICE_HOST_ONLY)
    b = rt.Scene.from_arrays(
        width=sc.width, height=sc.height, background=tuple(sc.bg), bucket_size=sc.bucket, camera_position=tuple(sc.cam_pos),
        camera_matrix=tuple(sc.cam_matrix), lights=[(*l["pos"], l["intensity"]) for l in sc.lights],
        materials=[dict(kind=int(m["kind"]), albedo=tuple(m["albedo"]), ior=float(m["ior"]), smooth_shading=int(m["smooth"]),
                        texture=int(m["texture"])) for m in sc.materials],
        meshes=[(m.material, m.vertices, m.uvs if len(m.uvs) else None, m.tris) for m in sc.meshes], device=rt.DEVICE_HOST_ONLY)
    for x, y in zip(a.tree() + a.geometry(), b.tree() + b.geometry()):
        assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x, y.view(np.uint32) if y.dtype == np.float32 else y)


CRT = {
    "settings": {"background_color": [0.1, 0.2, 0.3], "image_settings": {"width": 64, "height": 48}, "gi_on": True},
    "camera": {"matrix": [1, 0, 0, 0, 1, 0, 0, 0, 1], "position": [0, 0.5, 3]},
    "lights": [{"intensity": 170.5, "position": [1.1, 2.2, 0.3]}],
    "textures": [{"name": "chk", "type": "checker", "color_A": [1, 0, 0], "color_B": [0, 0, 1], "square_size": 0.125},
                 {"name": "edg", "type": "edges", "edge_color": [0, 0, 0], "inner_color": [1, 1, 1], "edge_width": 0.04},
                 {"name": "alb", "type": "albedo", "albedo": [0.3, 0.6, 0.9]}],
    "materials": [{"type": "diffuse", "albedo": "chk", "smooth_shading": False},
                  {"type": "refractive", "ior": 1.5, "smooth_shading": True, "albedo": [1, 1, 1], "back_face_culling": False},
                  {"type": "reflective", "albedo": [0.9, 0.9, 0.9], "smooth_shading": False},
                  {"type": "constant", "albedo": [0.5, 0.25, 0.125], "smooth_shading": False},
                  {"type": "diffuse", "albedo": [0.1, 0.7, 0.1], "smooth_shading": True}],
    "objects": [{"material_index": 0, "vertices": [-1, 0, -1, 1, 0, -1, 1, 0, 1, -1, 0, 1], "uvs": [0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0],
                 "triangles": [0, 1, 2, 0, 2, 3]},
                {"material_index": 4, "vertices": [0.1, 0.1, 0.1, 0.7, 0.1, 0.2, 0.3, 0.9, 0.123456789012345], "triangles": [0, 1, 2]}],
}


def test_crtscene_loader_follows_the_reference_schema(rt, oracle_mod, tmp_path):
    """C++ .crtscene loader == the test-side loader that produced the fixtures (io/json/loader.hpp semantics: float(double),
    3-component uvs with the third dropped, optional bucket_size, string albedo -> texture material, ignored keys)."""
    path = tmp_path / "t.crtscene"
    path.write_text(json.dumps(CRT))
    want = crtscene.to_rtsc_bytes(crtscene.load_crtscene(str(path)))
    a = rt.Scene.from_crtscene(str(path), device=rt.DEVICE_HOST_ONLY)
    b = rt.Scene.from_rtsc(want, device=rt.DEVICE_HOST_ONLY)
    assert (a.width, a.height) == (64, 48)
    for x, y in zip(a.tree() + a.geometry(), b.tree() + b.geometry()):
        assert x.tobytes() == y.tobytes()


REAL_SCENES = {"hw15_scene2": "scenes/hw15/scene2.crtscene", "hw09_scene5": "scenes/hw09/scene5.crtscene",
               "hw11_scene8": "scenes/hw11/scene8.crtscene", "hw12_scene4": "scenes/hw12/scene4.crtscene"}


def _real_scene_tree(tmp_path, name: str) -> str:
    """the reference's own scene file (and the one JPEG it ships), unpacked from tests/golden/crtscene/ into the directory layout
    the file's relative bitmap path expects (README.md:32-35: paths are relative to the project root)"""
    import gzip
    import shutil
    rel = REAL_SCENES[name]
    dst = tmp_path / rel
    dst.parent.mkdir(parents=True, exist_ok=True)
    with gzip.open(os.path.join(REPO, "tests", "golden", "crtscene", name + ".crtscene.gz"), "rb") as fh:
        dst.write_bytes(fh.read())
    tex = tmp_path / "scenes" / "hw12" / "textures"
    tex.mkdir(parents=True, exist_ok=True)
    shutil.copy(os.path.join(REPO, "tests", "golden", "crtscene", "textures", "dragon.jpg"), tex / "dragon.jpg")
    return str(dst)


@pytest.mark.parametrize("name", sorted(REAL_SCENES))
def test_crtscene_loader_reads_the_reference_scene_files(rt, tmp_path, name):
    """rt_scene_create_from_crtscene on the reference's REAL scene files (io/json/loader.hpp:235-265; configs 1-4 of
    BASELINE.json, committed as gz copies) gives, byte for byte, the RTSC fixture every parity test is fed - the fixture the
    compiled unmodified reference rendered the goldens from.  hw12/scene4 includes the JPEG texture (scene/texture/bitmap.hpp:
    11-37): the product's decoder (host/jpeg_decode.cpp) must produce the very texel bytes of the fixture, which an independent
    restatement (tests/helpers/jpeg_stb.py) decoded and which reproduce the published outputs/textures.png exactly."""
    path = _real_scene_tree(tmp_path, name)
    s = rt.Scene.from_crtscene(path, asset_root=str(tmp_path), device=rt.DEVICE_HOST_ONLY)
    assert s.export_rtsc() == scene_bytes(name)
    # and the test-side loader that writes the fixtures agrees with both
    assert crtscene.to_rtsc_bytes(crtscene.load_crtscene(path, root=str(tmp_path))) == scene_bytes(name)


def test_jpeg_decoder_edge_cases(rt, tmp_path):
    """the bitmap path beyond the one file the reference ships: grey-scale, 4:2:0 / 4:2:2 chroma, restart intervals and odd sizes
    decode (to within the usual +-3/255 of libjpeg, whose IDCT and upsampling differ in the last bits), progressive files and
    garbage are refused with a status instead of a crash"""
    from PIL import Image
    rng = np.random.default_rng(3)
    base = np.clip(np.add.outer(np.linspace(0, 200, 37), np.linspace(0, 55, 53))[..., None] + rng.normal(0, 6, (37, 53, 3)), 0, 255).astype(np.uint8)
    doc = json.loads(json.dumps(CRT))
    doc["textures"].append({"name": "bmp", "type": "bitmap", "file_path": "t.jpg"})

    def load(**save_kw):
        img = Image.fromarray(base if save_kw.pop("colour", True) else base[..., 0])
        img.save(tmp_path / "t.jpg", format="JPEG", quality=92, **save_kw)
        (tmp_path / "t.crtscene").write_text(json.dumps(doc))
        s = rt.Scene.from_crtscene(str(tmp_path / "t.crtscene"), asset_root=str(tmp_path), device=rt.DEVICE_HOST_ONLY)
        got = np.frombuffer(s.export_rtsc()[-37 * 53 * 3:], np.uint8).reshape(37, 53, 3)
        want = np.asarray(Image.open(tmp_path / "t.jpg").convert("RGB"))
        return np.abs(got.astype(int) - want.astype(int)).max()

    assert load(subsampling=0) <= 3                      # 4:4:4
    assert load(subsampling=1) <= 8                      # 4:2:2: libjpeg's default upsampling is another filter than the triangle one
    assert load(subsampling=2) <= 8                      # 4:2:0
    assert load(colour=False) <= 3                       # one component
    assert load(subsampling=2, restart_marker_blocks=3) <= 8
    with pytest.raises(rt.RtError) as e:
        load(progressive=True)
    assert e.value.status == rt.RT_ERR_UNSUPPORTED
    (tmp_path / "t.jpg").write_bytes(b"\xff\xd8\xff\xdb\x00\x03\x00" + bytes(40))
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_crtscene(str(tmp_path / "t.crtscene"), asset_root=str(tmp_path), device=rt.DEVICE_HOST_ONLY)
    assert e.value.status in (rt.RT_ERR_PARSE, rt.RT_ERR_UNSUPPORTED)


@pytest.mark.parametrize("mutate,status", [
    (lambda d: d["materials"].__setitem__(0, {"type": "glossy", "smooth_shading": False}), 5),      # loader.hpp:145
    (lambda d: d["textures"].__setitem__(0, {"name": "chk", "type": "noise"}), 5),                   # loader.hpp:104
    (lambda d: d["objects"][0].__setitem__("material_index", 9), 1),
    (lambda d: d["objects"][1].__setitem__("triangles", [0, 1, 7]), 1),
    (lambda d: d.pop("materials"), 5),                                                               # loader.hpp:256 throws
])
def test_crtscene_loader_errors(rt, tmp_path, mutate, status):
    doc = json.loads(json.dumps(CRT))
    mutate(doc)
    path = tmp_path / "bad.crtscene"
    path.write_text(json.dumps(doc))
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_crtscene(str(path), device=rt.DEVICE_HOST_ONLY)
    assert e.value.status == status


def test_error_statuses(rt, tmp_path):
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_rtsc(b"NOPE" + bytes(64), device=rt.DEVICE_HOST_ONLY)
    assert e.value.status == rt.RT_ERR_PARSE
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_rtsc(scene_bytes("hw12_scene4")[:-5], device=rt.DEVICE_HOST_ONLY)
    assert e.value.status == rt.RT_ERR_PARSE
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_crtscene(str(tmp_path / "missing.crtscene"), device=rt.DEVICE_HOST_ONLY)
    assert e.value.status == rt.RT_ERR_IO
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_rtsc(scene_bytes("hw12_scene4"), kd_max_depth=31, device=rt.DEVICE_HOST_ONLY)
    assert e.value.status == rt.RT_ERR_BAD_ARG
    assert rt.lib.rt_scene_create(None, None, None) == rt.RT_ERR_BAD_ARG


def test_compute_refuses_without_a_device(rt):
    """no CPU fallback: a host-only scene cannot trace or render"""
    s = rt.Scene.from_rtsc(scene_bytes("hw12_scene4"), device=rt.DEVICE_HOST_ONLY)
    rays = np.zeros((4, 6), np.float32)
    for call in (lambda: s.trace_closest(rays, True), lambda: s.trace_occluded(rays, np.ones(4, np.float32)),
                 lambda: s.render_frame(), lambda: s.render_frame_rgb8(), lambda: s.trace_primary(), lambda: s.counters()):
        with pytest.raises(rt.RtError) as e:
            call()
        assert e.value.status == rt.RT_ERR_NO_DEVICE


def test_partitions_cover_exactly(rt):
    for total in (1, 7, 64, 512):
        for world in (1, 2, 3, 8):
            got = [rt.spp_slice(total, r, world) for r in range(world)]
            assert sum(n for _, n in got) == total
            assert all(got[i][0] + got[i][1] == got[i + 1][0] for i in range(world - 1))
            assert max(n for _, n in got) - min(n for _, n in got) <= 1
    for h, b in ((1080, 64), (2160, 24), (5, 64)):
        for world in (1, 2, 8):
            rows = sorted(y for r in range(world) for y0, y1 in rt.row_bands(h, b, r, world) for y in range(y0, y1))
            assert rows == list(range(h))


@pytest.mark.parametrize("kd", [(24, 64), (8, 64)])
def test_parallel_builder_is_thread_invariant(rt, oracle_mod, kd, monkeypatch):
    """the task-parallel host builder (host/kd_parallel.hpp) must emit the reference's tree - same
    median splits, same overlap test, same DFS order - whatever the thread count; the backend's own SAH tree likewise."""
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=40_000, seed=11, width=64, height=48))
    o = oracle_mod.Oracle(data, kd[0], kd[1])
    on5, obx, orf = o.tree()
    layouts = []
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("RT_B200_BUILD_THREADS", threads)
        s = rt.Scene.from_rtsc(data, kd_max_depth=kd[0], kd_max_leaf_size=kd[1], device=rt.DEVICE_HOST_ONLY)
        n5, bx, rf = s.tree()
        assert np.array_equal(n5, on5) and np.array_equal(bx.view(np.uint32), obx.view(np.uint32)) and np.array_equal(rf, orf)
        assert (s.info.n_leaves, s.info.max_leaf_refs, s.info.tree_depth) == (o.n_leaves, o.max_leaf_refs, o.tree_depth)
        layouts.append([np.ascontiguousarray(a).tobytes() for a in s.device_layout() + s.bvh_layout()] +
                       [(s.info.bvh_n_nodes, s.info.bvh_n_leaves, s.info.bvh_n_refs, s.info.bvh_depth, s.info.bvh4_n_nodes, s.info.bvh4_stack_need)])
        s.close()
    assert layouts[0] == layouts[1] == layouts[2]


def test_bvh_builder_big_nodes_are_thread_invariant(rt, monkeypatch):
    """the top nodes of a big scene (>= 2^18 triangles) are binned and partitioned by all build threads (host/bvh_build.cpp):
    counts and boxes merge by + / min / max and the partition keeps the list order, so the tree must not depend on the
    thread count - one thread (everything sequential) against three and eight"""
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=600_000, seed=5, width=64, height=48))
    layouts = []
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("RT_B200_BUILD_THREADS", threads)
        s = rt.Scene.from_rtsc(data, kd_max_depth=4, kd_max_leaf_size=64, device=rt.DEVICE_HOST_ONLY)
        layouts.append([np.ascontiguousarray(a).tobytes() for a in s.bvh_layout()] +
                       [(s.info.bvh_n_nodes, s.info.bvh_n_leaves, s.info.bvh_n_refs, s.info.bvh_depth, s.info.bvh4_n_nodes, s.info.bvh4_stack_need)])
        ids = s.bvh_layout()[1][:, 3]
        assert np.array_equal(np.sort(ids), np.arange(s.info.n_triangles, dtype=np.uint32))
        s.close()
    assert layouts[0] == layouts[1] == layouts[2]
