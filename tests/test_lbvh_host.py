"""The device-side hierarchy builder (simd-raytracer_b200/csrc/rt_lbvh.cuh, rt_build_opts.accel_build = device) checked on the
CPU: tests/helpers/lbvh_host.cpp runs the builder's kernels thread by thread in the order the host glue launches them; the tree
that comes out is checked structurally and walked by the same traversal sources the CUDA kernels compile (tests/helpers/
kd8_host.cpp) against the oracle's reference-order traversal.  No product compute runs here."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from .conftest import REPO, resized, scene_bytes
from .helpers import crtscene
from .test_kd8_host import scene_rays


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    d = tmp_path_factory.mktemp("lbvh")
    out = {}
    for name in ("lbvh_host", "kd8_host"):
        so = d / f"lib{name}.so"
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               os.path.join(REPO, "tests", "helpers", name + ".cpp"), "-o", str(so)])
        out[name] = C.CDLL(str(so))
    out["lbvh_host"].lbvh_build_host.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    out["lbvh_host"].lbvh_fetch_host.argtypes = [C.c_void_p] * 4
    k = out["kd8_host"]
    k.bvh_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_float,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    k.bvh4_trace_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_float,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return out


def build(libs, scene, leaf=4):
    tri9, _, _ = scene.geometry()
    _, bx, _ = scene.tree()
    root6 = np.ascontiguousarray(bx[0], np.float32)
    out6 = np.zeros(6, np.uint32)
    rc = libs["lbvh_host"].lbvh_build_host(tri9.ctypes.data, len(tri9), root6.ctypes.data, leaf, out6.ctypes.data)
    assert rc == 0, rc
    n2, n4, need, depth2, same, host_need = (int(v) for v in out6)
    nodes16 = np.zeros((n2, 16), np.uint32)
    tris12 = np.zeros((len(tri9), 12), np.uint32)
    nodes32 = np.zeros((n4, 32), np.uint32)
    root = np.zeros(6, np.float32)
    libs["lbvh_host"].lbvh_fetch_host(nodes16.ctypes.data, tris12.ctypes.data, nodes32.ctypes.data, root.ctypes.data)
    return dict(leaf=leaf, tri9=tri9, nodes16=nodes16, tris12=tris12, nodes32=nodes32, root=root, need=need, depth2=depth2, same=same, host_need=host_need)


def check_structure(b):
    tri9, nodes16, tris12 = b["tri9"], b["nodes16"], b["tris12"]
    n = len(tri9)
    # every triangle once, its record the geometry's numbers bit for bit
    ids = tris12[:, 3]
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32))
    assert np.array_equal(tris12[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]], tri9[ids].view(np.uint32))
    # the two-wide tree: leaves tile [0, n) once, every child box encloses the vertices of the triangles below it
    f = nodes16[:, :12].view(np.float32)
    corners = np.stack([tri9[:, 0:3], tri9[:, 0:3] + tri9[:, 3:6], tri9[:, 0:3] + tri9[:, 6:9]], axis=1)[ids]       # sorted order
    tmin, tmax = corners.min(axis=1), corners.max(axis=1)
    covered = np.zeros(n, np.int32)
    seen = np.zeros(len(nodes16), np.int32)
    # bottom-up boxes of the records below every node, by an explicit post-order walk
    lo = np.zeros((len(nodes16), 2, 3), np.float32)
    hi = np.zeros((len(nodes16), 2, 3), np.float32)
    order, todo = [], [0]
    while todo:
        i = todo.pop()
        seen[i] += 1
        order.append(i)
        for s in range(2):
            ref, cnt = int(nodes16[i, 12 + s]), int(nodes16[i, 14 + s])
            assert cnt != 0xFFFFFFFF and cnt <= b["leaf"]
            if cnt == 0:
                todo.append(ref)
    assert np.all(seen == 1)
    depth = np.zeros(len(nodes16), np.int32)
    for i in reversed(order):
        for s in range(2):
            ref, cnt = int(nodes16[i, 12 + s]), int(nodes16[i, 14 + s])
            if cnt:
                covered[ref:ref + cnt] += 1
                lo[i, s], hi[i, s] = tmin[ref:ref + cnt].min(axis=0), tmax[ref:ref + cnt].max(axis=0)
            else:
                lo[i, s], hi[i, s] = lo[ref].min(axis=0), hi[ref].max(axis=0)
                depth[i] = max(depth[i], depth[ref] + 1)
    assert np.all(covered == 1)
    box_lo = np.stack([f[:, 0:3], f[:, 6:9]], axis=1)
    box_hi = np.stack([f[:, 3:6], f[:, 9:12]], axis=1)
    assert np.all(box_lo <= lo) and np.all(box_hi >= hi)
    assert depth[0] + 1 == b["depth2"]
    # the level-by-level collapse is the host collapse of the same tree; the stack need it reports is the host's
    assert b["same"] == 1 and b["need"] == b["host_need"]
    assert b["depth2"] <= 44 and b["need"] <= 136


def trace(libs, b, rays, cull, wide):
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    tuv = np.zeros((len(rays), 3), np.float32)
    tri = np.zeros(len(rays), np.int32)
    tie = np.zeros(len(rays), np.uint8)
    visits = np.zeros(len(rays), np.uint32)
    k = libs["kd8_host"]
    eps = C.c_float(1e-6)
    if wide:
        k.bvh4_trace_batch(b["nodes16"].ctypes.data, len(b["nodes16"]), b["tris12"].ctypes.data, b["root"].ctypes.data, rays.ctypes.data, len(rays),
                           int(cull), 0, eps, None, 0, tuv.ctypes.data, tri.ctypes.data, tie.ctypes.data, visits.ctypes.data)
    else:
        k.bvh_trace_batch(b["nodes16"].ctypes.data, b["tris12"].ctypes.data, b["root"].ctypes.data, rays.ctypes.data, len(rays), int(cull), 0, eps,
                          None, 0, tuv.ctypes.data, tri.ctypes.data, tie.ctypes.data)
    return tuv, tri, tie.astype(bool), visits


def check_hits(libs, b, o, rays, cull):
    want_tuv, want_tri = o.trace(rays, cull)
    for wide in (False, True):
        tuv, tri, tie, visits = trace(libs, b, rays, cull, wide)
        rr = tri == -3
        assert np.array_equal((tri >= 0)[~rr], (want_tri >= 0)[~rr])
        h = (want_tri >= 0) & ~rr
        assert np.array_equal(tuv[h, 0].view(np.uint32), want_tuv[h, 0].view(np.uint32))
        assert np.array_equal(tri[~tie], want_tri[~tie])
        assert np.array_equal(tuv[h & ~tie].view(np.uint32), want_tuv[h & ~tie].view(np.uint32))
        if wide:
            assert int((visits >> 16).max()) <= b["need"]               # the deepest stack any query reached is within the reported need


@pytest.mark.parametrize("name", ["hw09_scene5", "hw11_scene8"])
def test_device_builder_on_fixture_scenes(rt, oracle_mod, libs, name):
    data = resized(scene_bytes(name), 320, 180)
    s = rt.Scene.from_rtsc(data, device=rt.DEVICE_HOST_ONLY)
    o = oracle_mod.Oracle(data)
    b = build(libs, s)
    check_structure(b)
    check_hits(libs, b, o, o.primary_rays(), True)
    check_hits(libs, b, o, scene_rays(o, s, 40_000), False)
    # the host-built tree of the same scene: the stack need scene creation computed inside the collapse = the separate pass over the result
    n16, _, _ = s.bvh_layout()
    libs["kd8_host"].bvh4_stack_need_host.restype = C.c_uint64
    libs["kd8_host"].bvh4_stack_need_host.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    assert libs["kd8_host"].bvh4_stack_need_host(n16.ctypes.data, len(n16), None) == s.info.bvh4_stack_need


@pytest.mark.parametrize("leaf", [4, 1])
def test_device_builder_on_a_synthetic_mesh(rt, oracle_mod, libs, leaf):
    """leaf = 1 is what a scene beyond L2 is built with (one triangle per leaf, rt_api.cu finish_create)"""
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=20_000, seed=9, width=160, height=120))
    s = rt.Scene.from_rtsc(data, device=rt.DEVICE_HOST_ONLY)
    o = oracle_mod.Oracle(data)
    b = build(libs, s, leaf)
    check_structure(b)
    check_hits(libs, b, o, o.primary_rays(), True)
    check_hits(libs, b, o, scene_rays(o, s, 40_000), False)


def test_device_builder_with_coincident_triangles(rt, libs):
    """equal Morton codes are told apart by their sorted position: 200 copies of one triangle among others still give a valid tree"""
    base = crtscene.synthetic_scene(n_tris=2_000, seed=4, width=64, height=48)
    m = base.meshes[0]
    m.tris = np.concatenate([m.tris, np.repeat(m.tris[:1], 200, axis=0)]).astype(np.uint32)
    s = rt.Scene.from_rtsc(crtscene.to_rtsc_bytes(base), device=rt.DEVICE_HOST_ONLY)
    assert s.info.n_triangles >= 2_200 - 8
    b = build(libs, s)
    check_structure(b)


def test_device_build_needs_a_device_and_valid_option(rt):
    data = resized(scene_bytes("hw09_scene5"), 64, 36)
    with pytest.raises(rt.RtError):
        rt.Scene.from_rtsc(data, device=rt.DEVICE_HOST_ONLY, accel_build=rt.ACCEL_BUILD_DEVICE)
    with pytest.raises(rt.RtError):
        rt.Scene.from_rtsc(data, device=rt.DEVICE_HOST_ONLY, accel_build=7)


@pytest.mark.parametrize("shape", ["strip", "clusters", "coincident", "plane"])
def test_device_builder_on_degenerate_distributions(rt, oracle_mod, libs, shape):
    """distributions a Morton-code builder could trip over: every triangle on one line (two of three code axes constant), two tight
    clusters far apart (long common prefixes), 3,000 copies of ONE triangle (equal codes, told apart by position only), a flat
    plane - the tree must stay valid, within the traversal's depth / stack limits, and give the oracle's hits"""
    rng = np.random.default_rng(17)
    n = 6000
    base = crtscene.synthetic_scene(n_tris=n, seed=21, width=96, height=64)
    m = base.meshes[0]
    v = np.asarray(m.vertices, np.float32).reshape(n, 3, 3).copy()
    size = 0.02
    if shape == "strip":
        c = np.zeros((n, 1, 3), np.float32); c[:, 0, 0] = np.linspace(-1, 1, n)
    elif shape == "clusters":
        c = (rng.uniform(-1, 1, (n, 1, 3)) * 1e-3).astype(np.float32); c[: n // 2, 0, :] += 1.0; c[n // 2:, 0, :] -= 1.0
        size = 1e-3
    elif shape == "coincident":
        c = rng.uniform(-1, 1, (n, 1, 3)).astype(np.float32)
    else:
        c = rng.uniform(-1, 1, (n, 1, 3)).astype(np.float32); c[:, 0, 1] = 0.25
    v = (c + size * rng.uniform(-1, 1, (n, 3, 3))).astype(np.float32)
    if shape == "coincident":
        v[: n // 2] = v[0]
    if shape == "plane":
        v[:, :, 1] = 0.25
    m.vertices = v.reshape(-1, 3)
    data = crtscene.to_rtsc_bytes(base)
    s = rt.Scene.from_rtsc(data, device=rt.DEVICE_HOST_ONLY)
    o = oracle_mod.Oracle(data)
    for leaf in (4, 1):
        b = build(libs, s, leaf)
        check_structure(b)
        check_hits(libs, b, o, o.primary_rays(), True)
        check_hits(libs, b, o, scene_rays(o, s, 20_000), False)
