"""The accelerated traversals (RT_FLAG_ORDERED) checked on the CPU: the two-wide bounding-volume hierarchy
(simd-raytracer_b200/csrc/rt_bvh.cuh) and its four-wide collapse (rt_bvh4.cuh, host/bvh4_collapse.hpp).  The same source the CUDA
kernels compile is built as plain C++ (tests/helpers/kd8_host.cpp) and run over the product's flattened structures
(rt_scene_get_bvh_layout, host-only scene) against the oracle's reference-order traversal.  No product compute runs here."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from .conftest import REPO, SCENES, resized, scene_bytes
from .helpers import crtscene


@pytest.fixture(scope="module", params=["bvh", "bvh4"])
def kd8(tmp_path_factory, request):
    structure = request.param
    out = tmp_path_factory.mktemp("kd8") / "libkd8_host.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                           os.path.join(REPO, "tests", "helpers", "kd8_host.cpp"), "-o", str(out)])
    lib = C.CDLL(str(out))
    for fn in (lib.bvh_trace_batch,):
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_float,
                       C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bvh4_trace_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_float,
                                     C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    batch = lib.bvh_trace_batch

    def trace(scene, rays, cull, fast=False, t_far=None, any_hit=False, eps=np.float32(1e-6)):
        nodes8, packets, root = scene.bvh_layout()
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        tuv = np.zeros((len(rays), 3), np.float32)
        tri = np.zeros(len(rays), np.int32)
        tie = np.zeros(len(rays), np.uint8)
        far = None if t_far is None else np.ascontiguousarray(t_far, np.float32)
        if structure == "bvh4":
            # the four-wide hierarchy (csrc/rt_bvh4.cuh), collapsed from the scene's two-wide nodes (host/bvh4_collapse.hpp)
            lib.bvh4_trace_batch(nodes8.ctypes.data, int(scene.info.bvh_n_nodes), packets.ctypes.data, root.ctypes.data, rays.ctypes.data, len(rays),
                                 int(cull), int(fast), C.c_float(eps), None if far is None else far.ctypes.data, int(any_hit),
                                 tuv.ctypes.data, tri.ctypes.data, tie.ctypes.data, None)
            return tuv, tri, tie.astype(bool)
        batch(nodes8.ctypes.data, packets.ctypes.data, root.ctypes.data, rays.ctypes.data, len(rays), int(cull), int(fast),
                            C.c_float(eps), None if far is None else far.ctypes.data, int(any_hit), tuv.ctypes.data, tri.ctypes.data, tie.ctypes.data)
        return tuv, tri, tie.astype(bool)

    def counters(reset=True):
        n, t = C.c_uint64(0), C.c_uint64(0)
        lib.kd8_counters(C.byref(n), C.byref(t), int(reset))
        return n.value, t.value
    trace.counters = counters
    return trace


def scene_rays(o, s, n=150_000, seed=3):
    rng = np.random.default_rng(seed)
    _, bx, _ = s.tree()
    c, ext = (bx[0, :3] + bx[0, 3:]) / 2, (bx[0, 3:] - bx[0, :3]).max()
    org = rng.uniform(-1, 1, (n, 3)) * ext * 0.9 + c
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, d], axis=1).astype(np.float32)
    rays[0, 3:] = (0, 0, -1); rays[1, 3:] = (1, 0, 0); rays[2, 3:] = (0, -1, 0); rays[3, 3:] = (0, 0, 0); rays[4, :3] = np.nan
    rays[5, 3:] = (-0.0, 1, 0); rays[6, :3] = c; rays[7, :3] = bx[0, :3]
    rays[8, 3:] = (1e-39, 0.6, 0.8); rays[9, 3:] = (0.6, -1e-35, 0.8); rays[10, 3:] = (0.6, 0.8, 1e-20); rays[11, 3:] = (-1e-25, -1e-25, 1)
    k = min(n // 4, 3000)                                  # axis-parallel rays from many origins
    rays[12:12 + k, 3:] = np.eye(3, dtype=np.float32)[np.arange(k) % 3] * np.where(np.arange(k) % 2, -1.0, 1.0).astype(np.float32)[:, None]
    return rays


@pytest.mark.parametrize("name", SCENES)
def test_kd8_equals_reference_traversal(rt, oracle_mod, kd8, name):
    """primary rays (culling) at reduced resolution + incoherent rays (no culling): same hit/miss, t bit-exact, and the same
    triangle with the same u/v on every ray the traversal does not flag as an exact-t tie between different triangles
    (flagged rays - a shared edge, or config 1's cube standing on the floor - are re-run in reference order on the device)"""
    data = resized(scene_bytes(name), 640, 360 if name != "hw15_scene2" else 640)
    s = rt.Scene.from_rtsc(data, device=rt.DEVICE_HOST_ONLY)
    o = oracle_mod.Oracle(data)
    assert s.info.bvh_n_nodes >= 1 and s.info.bvh_n_refs == s.info.n_triangles
    for rays, cull in ((o.primary_rays(), True), (scene_rays(o, s), False)):
        want_tuv, want_tri = o.trace(rays, cull)
        tuv, tri, tie = kd8(s, rays, cull)
        rr = tri == -3                                      # KD_RERUN: handed to the reference-order traversal on the device
        assert np.array_equal((tri >= 0)[~rr], (want_tri >= 0)[~rr])
        h = (want_tri >= 0) & ~rr
        assert np.array_equal(tuv[h, 0].view(np.uint32), want_tuv[h, 0].view(np.uint32))
        assert tie.mean() < (2e-3 if cull else 2.5e-2), (name, cull, tie.mean())     # incoherent set: 2 % axis-parallel rays are re-run
        assert np.array_equal(tri[~tie], want_tri[~tie])
        assert np.array_equal(tuv[h & ~tie].view(np.uint32), want_tuv[h & ~tie].view(np.uint32))


@pytest.mark.parametrize("kd", [(8, 64), (24, 64)])
def test_kd8_synthetic_mesh(rt, oracle_mod, kd8, kd):
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=20_000, seed=9, width=160, height=120))
    s = rt.Scene.from_rtsc(data, kd_max_depth=kd[0], kd_max_leaf_size=kd[1], device=rt.DEVICE_HOST_ONLY)
    o = oracle_mod.Oracle(data, *kd)
    for rays, cull in ((o.primary_rays(), True), (scene_rays(o, s, 60_000), False)):
        want_tuv, want_tri = o.trace(rays, cull)
        tuv, tri, tie = kd8(s, rays, cull)
        rr = tri == -3
        assert np.array_equal(tri[~tie], want_tri[~tie]) and np.array_equal((tri >= 0)[~rr], (want_tri >= 0)[~rr])
        h = (want_tri >= 0) & ~tie
        assert np.array_equal(tuv[h].view(np.uint32), want_tuv[h].view(np.uint32))


def test_kd8_any_hit_and_far_limit(rt, oracle_mod, kd8):
    """t_far: a hit beyond it may be reported as a miss, a hit inside it never; any_hit returns SOME hit inside it"""
    data = resized(scene_bytes("hw09_scene5"), 320, 180)
    s = rt.Scene.from_rtsc(data, device=rt.DEVICE_HOST_ONLY)
    o = oracle_mod.Oracle(data)
    rays = scene_rays(o, s, 80_000, seed=8)
    want_tuv, want_tri = o.trace(rays, False)
    far = np.random.default_rng(2).uniform(0.05, 3.0, len(rays)).astype(np.float32)
    inside = (want_tri >= 0) & (want_tuv[:, 0] <= far)
    tuv, tri, tie = kd8(s, rays, False, t_far=far)
    inside &= ~tie
    assert np.array_equal(tri[inside], want_tri[inside]) and np.array_equal(tuv[inside, 0], want_tuv[inside, 0])
    assert not np.any((tri >= 0) & (tuv[:, 0] <= far) & ~((want_tri >= 0) & (want_tuv[:, 0] <= far)))
    tuv, tri, _ = kd8(s, rays, False, t_far=far, any_hit=True)
    rr = tri == -3                                          # KD_RERUN (axis-parallel rays): re-run in reference order on the device
    assert rr.mean() < 0.05
    assert np.array_equal(((tri >= 0) & (tuv[:, 0] <= far))[~rr], ((want_tri >= 0) & (want_tuv[:, 0] <= far))[~rr])


def test_origin_on_split_planes_stays_cheap(rt, oracle_mod, kd8):
    """The synthetic scene's camera sits at x = y = 0, exactly on the binned SAH planes of the top of the tree.  A query from
    such an origin must stay on ONE side of every such plane (the side its direction points to); visiting both halves
    multiplied the work per ray by 12 on config 5 without changing a single hit, so only a work bound can catch it."""
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=100_000, seed=1234, width=160, height=90))
    s = rt.Scene.from_rtsc(data, kd_max_depth=24, kd_max_leaf_size=64, device=rt.DEVICE_HOST_ONLY)
    o = oracle_mod.Oracle(data, 24, 64)
    rays = o.primary_rays()
    kd8.counters()
    tuv, tri, tie = kd8(s, rays, True)
    nodes, tris = kd8.counters()
    want_tuv, want_tri = o.trace(rays, True)
    assert np.array_equal(tri[~tie], want_tri[~tie]) and np.array_equal(tri >= 0, want_tri >= 0)
    assert nodes / len(rays) < 40 and tris / len(rays) < 12, (nodes / len(rays), tris / len(rays))


def test_rays_from_walls_edges_and_triangles(rt, oracle_mod, kd8):
    """Adversarial origins: exactly on the scene's boundary walls, their edges and corners, and on triangles (edges and
    vertices included), with generic and nearly axis-parallel directions.  The reference accepts a hit when its ROUNDED u, v, t
    pass, so accepted hit points may sit a few ulps outside a triangle's own box (-> the BVH boxes are padded), and whether a
    ray that touches the scene box at a single parameter value still meets a wall depends on the reference's own leaf boxes
    (-> such rays carry the re-run mark and are answered in reference order on the device).  Everything else must agree bit
    for bit."""
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=20_000, seed=77, width=64, height=48))
    s = rt.Scene.from_rtsc(data, kd_max_depth=24, kd_max_leaf_size=64, device=rt.DEVICE_HOST_ONLY)
    o = oracle_mod.Oracle(data, 24, 64)
    rng = np.random.default_rng(5)
    n = 60_000
    t9, _, _ = s.geometry()

    def on_walls():
        p = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
        ax = rng.integers(0, 3, n)
        p[np.arange(n), ax] = rng.choice([-1.5, 1.5], n).astype(np.float32)
        k = n // 4
        p[np.arange(k), (ax[:k] + 1) % 3] = rng.choice([-1.5, 1.5], k).astype(np.float32)      # edges
        return p

    def on_tris():
        k = rng.integers(0, len(t9), n)
        u, v = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
        f = u + v > 1
        u[f], v[f] = 1 - u[f], 1 - v[f]
        v[rng.uniform(size=n) < 0.3] = 0                                                        # edges
        w = rng.uniform(size=n) < 0.1
        u[w], v[w] = 0, 0                                                                       # vertices
        return (t9[k, :3] + u[:, None] * t9[k, 3:6] + v[:, None] * t9[k, 6:9]).astype(np.float32)

    def dirs(flat):
        d = rng.normal(size=(n, 3))
        if flat:
            d[np.arange(n), rng.integers(0, 3, n)] *= 1e-4
        return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)

    for org in (on_walls, on_tris):
        for flat in (False, True):
            rays = np.concatenate([org(), dirs(flat)], axis=1)
            for cull in (False, True):
                want_tuv, want_tri = o.trace(rays, cull)
                tuv, tri, rerun = kd8(s, rays, cull)
                ok = ~rerun
                assert np.array_equal(tri[ok], want_tri[ok])
                h = ok & (want_tri >= 0)
                assert np.array_equal(tuv[h].view(np.uint32), want_tuv[h].view(np.uint32))
                if org is on_tris:
                    assert rerun.mean() < 2e-2


# ---- tile culling (csrc/rt_tilecull.cuh) --------------------------------------------------------------------------------------
def _camera_dirs(m, W, H, tan_half, xs, ys):
    """un-normalised camera-ray directions of render/render.hpp:47-60 in float64 for raster positions xs x ys"""
    X, Y = np.meshgrid(xs, ys)
    sx = (2.0 * (X / W) - 1.0) * (W / H) * tan_half
    sy = (1.0 - 2.0 * (Y / H)) * tan_half
    m = m.reshape(3, 3)
    v = np.stack([sx, sy, -np.ones_like(sx)], axis=-1)
    return v @ m                                           # transpose(M) * v  ==  v (row) * M


def _hits_box(o, d, lo, hi):
    """exact (float64) slab test, t >= 0, of rays o + t d against the closed box"""
    with np.errstate(divide="ignore", invalid="ignore"):
        t1, t2 = (lo - o) / d, (hi - o) / d
    tn, tf = np.minimum(t1, t2), np.maximum(t1, t2)
    par = d == 0                                            # parallel to a slab: inside it or never
    inside = (o >= lo) & (o <= hi)
    tn = np.where(par, np.where(inside, -np.inf, np.inf), tn)
    tf = np.where(par, np.where(inside, np.inf, -np.inf), tf)
    return np.maximum(tn.max(-1), 0.0) <= tf.min(-1)


def test_tile_culling_predicate_is_conservative_and_effective(tmp_path):
    """k_tile_cull's predicate on the CPU: whenever it says "no ray of this 8x4 tile reaches the box", an exact float64 slab
    test of a 9x9 grid of raster positions in the tile (corners and edges included) agrees for every one of them; and it is
    not vacuous - over cameras that see the box in a part of the frame most of the empty tiles are culled."""
    out = tmp_path / "libkd8_host.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                           os.path.join(REPO, "tests", "helpers", "kd8_host.cpp"), "-o", str(out)])
    lib = C.CDLL(str(out))
    lib.tile_misses_box_host.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p] + [C.c_float] * 4
    rng = np.random.default_rng(11)
    W, H = 192, 120
    culled = empty = wrong = 0
    for k in range(40):
        ang = rng.uniform(-np.pi, np.pi, 3)
        cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
        R = (np.asarray([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]) @ np.asarray([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @
             np.asarray([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]))
        if k % 4 == 3:
            R = R @ np.asarray([[1.0, 0.4, 0.0], [0.0, 0.7, 0.1], [0.2, 0.0, 1.3]])
        m = np.ascontiguousarray(R.reshape(-1), np.float32)
        lo = rng.uniform(-1, 0.5, 3).astype(np.float32)
        hi = (lo + rng.uniform(0.05, 1.5, 3)).astype(np.float32)
        if k % 5 == 0:
            hi[k % 3] = lo[k % 3]                            # a flat box (the floor of a scene)
        pos = rng.uniform(-3, 3, 3).astype(np.float32)
        if np.all((lo <= pos) & (pos <= hi)):
            pos[0] = hi[0] + 1.0                             # the host only culls with the camera outside the box
        tan_half = np.float32(np.tan(np.radians(rng.choice([30.0, 90.0, 150.0])) / 2))
        m64, pos64, lo64, hi64 = m.astype(np.float64), pos.astype(np.float64), lo.astype(np.float64), hi.astype(np.float64)
        for ty in range(0, H, 4):
            for tx in range(0, W, 8):
                x1, y1 = min(tx + 8, W), min(ty + 4, H)
                d = _camera_dirs(m64, W, H, float(tan_half), np.linspace(tx, x1, 9), np.linspace(ty, y1, 9)).reshape(-1, 3)
                hit = _hits_box(pos64, d, lo64, hi64)
                says_miss = lib.tile_misses_box_host(m.ctypes.data, pos.ctypes.data, W, H, tan_half, lo.ctypes.data, hi.ctypes.data, tx, ty, x1, y1)
                empty += int(not hit.any())
                culled += int(says_miss)
                wrong += int(says_miss and hit.any())
    assert wrong == 0
    assert culled >= 0.8 * empty, (culled, empty)            # sub-pixel margin: nearly every empty tile is culled


def test_bvh4_stack_bound_on_a_degenerate_chain(tmp_path):
    """the four-wide traversal's stack must be PROVEN deep enough.  A two-wide hierarchy that is one long chain
    (every node = one leaf + the rest of the chain, 44 levels - the builder's depth cap) is the worst shape for the collapse:
    host/bvh4_collapse.hpp computes the stack entries any ray can need, that number must stay within BVH4_STACK (3 * 44 + 4), and
    the four-wide traversal must answer rays through all 45 nested boxes exactly as the two-wide one does."""
    out = tmp_path / "libkd8_host.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                           os.path.join(REPO, "tests", "helpers", "kd8_host.cpp"), "-o", str(out)])
    lib = C.CDLL(str(out))
    lib.bvh4_stack_need_host.restype = C.c_uint64
    lib.bvh4_stack_need_host.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    depth = 44
    nodes = np.zeros((depth, 16), np.uint32)
    tris = np.zeros((depth + 1, 12), np.float32)
    f = nodes.view(np.float32)
    for i in range(depth + 1):                        # triangle i: a small quad half in the plane z = i, all inside the nested boxes
        tris[i, 0:3] = (-0.4, -0.4, float(i)); tris[i, 4:7] = (0.8, 0.0, 0.0); tris[i, 8:11] = (0.0, 0.8, 0.0)
        tris[i].view(np.uint32)[3] = i
    for i in range(depth):
        f[i, 0:6] = (-0.5, -0.5, i - 0.01, 0.5, 0.5, i + 0.01)                       # child 0: the leaf with triangle i
        f[i, 6:12] = (-0.5, -0.5, i + 0.99, 0.5, 0.5, depth + 0.01)                  # child 1: the rest of the chain
        nodes[i, 12:16] = (i, i + 1, 1, 0) if i + 1 < depth else (i, depth, 1, 1)    # ref0, ref1, cnt0, cnt1
    cap = C.c_uint64(0)
    need = lib.bvh4_stack_need_host(nodes.ctypes.data, depth, C.byref(cap))
    assert 3 <= need <= cap.value == 3 * 44 + 4
    lib.bvh_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_float,
                                    C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bvh4_trace_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_float,
                                     C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rays = np.zeros((64, 6), np.float32)
    rng = np.random.default_rng(1)
    rays[:, 0:2] = rng.uniform(-0.45, 0.45, (64, 2)); rays[:, 2] = depth + 1.0
    rays[:, 3:5] = rng.uniform(-0.01, 0.01, (64, 2)); rays[:, 5] = -1.0                # down through all 45 planes: the LAST one is the farthest
    rays[:32, 2] = -1.0; rays[:32, 5] = 1.0                                          # and up: triangle 0 first
    root = np.array([-0.5, -0.5, -0.01, 0.5, 0.5, depth + 0.01], np.float32)
    res = []
    for wide in (False, True):
        tuv, tri, tie = np.zeros((64, 3), np.float32), np.zeros(64, np.int32), np.zeros(64, np.uint8)
        args = [nodes.ctypes.data] + ([depth] if wide else []) + [tris.ctypes.data, root.ctypes.data, rays.ctypes.data, 64, 0, 0,
                C.c_float(1e-6), None, 0, tuv.ctypes.data, tri.ctypes.data, tie.ctypes.data] + ([None] if wide else [])
        (lib.bvh4_trace_batch if wide else lib.bvh_trace_batch)(*args)
        res.append((tuv.copy(), tri.copy()))
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][0].view(np.uint32), res[1][0].view(np.uint32))
    assert (res[0][1] >= 0).sum() >= 16 and len(set(res[0][1].tolist())) >= 8          # rays end on many different links of the chain
