"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's goldens.  Needs a B200: -m gpu.

Bars (BASELINE.json north_star):
  exact mode (default): bit-exact hits (t, u, v, triangle index), bit-exact float frames and ray counts on the
                        deterministic configs - i.e. 100 %, stricter than the 99.99 % / 1e-5 the north star asks for
  fast mode (FMA)     : >= 99.99 % hit/miss + triangle agreement on primary rays, frame PSNR >= 50 dB
  GI / multi-sample   : same Philox stream as the oracle; only cosf/sinf may differ in the last bit, so frames agree on
                        >= 99 % of pixels and to PSNR >= 45 dB (the reference itself is not run-to-run reproducible
                        there: 32.95 dB)
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import pytest

import os

from .conftest import HERE, SCENES, check_textures_png, psnr8, quantise, resized, scene_bytes, sha
from .helpers import crtscene

pytestmark = pytest.mark.gpu

CONFIGS = [(n, "s1d5g0", 5) for n in SCENES] + [("hw11_scene8", "s1d10g0", 10)]

_open: dict = {}


def gpu_scene(rt, name: str, size=None, kd=(8, 64), width=0):
    """width: rt_build_opts.accel_width - the hierarchy the accelerated mode walks (0 = the default, four-wide; 2 = two-wide)"""
    key = (name, size, kd, width)
    if key not in _open:
        data = scene_bytes(name)
        if size:
            data = resized(data, *size)
        _open[key] = (rt.Scene.from_rtsc(data, kd_max_depth=kd[0], kd_max_leaf_size=kd[1], accel_width=width), data)
    return _open[key]


WIDTHS = [4, 2]          # both hierarchies of the accelerated mode (csrc/rt_bvh4.cuh, csrc/rt_bvh.cuh) are held to the same bar


def assert_hits_equal(hits, tuv, tri):
    assert np.array_equal(hits["tri"], tri)
    h = tri >= 0
    for k, f in enumerate(("t", "u", "v")):
        assert np.array_equal(hits[f][h].view(np.uint32), tuv[h, k].view(np.uint32)), f


# ---- closest hit ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SCENES)
def test_primary_hits_bit_exact_full_resolution(rt, oracle_mod, golden, name):
    s, data = gpu_scene(rt, name)
    o = oracle_mod.Oracle(data)
    hits = s.trace_primary().reshape(-1)
    tuv, tri = o.trace(o.primary_rays(), True)
    assert_hits_equal(hits, tuv, tri)
    assert int((hits["tri"] >= 0).sum()) == golden["scenes"][name]["configs"]["s1d5g0"]["counts"]["cull_hit"]


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("tag", ["default", "gi"])
def test_reference_query_stream_bit_exact(rt, records, name, tag):
    """the reference's own queries (ray in, t/u/v/triangle out; culling on for primaries, off for the rest)"""
    s, _ = gpu_scene(rt, name)
    rec = records(name)[f"{tag}_records"]
    rays = np.concatenate([rec["o"], rec["d"]], axis=1)
    for cull in (0, 1):
        sel = rec["cull"] == cull
        if sel.any():
            hits = s.trace_closest(rays[sel], bool(cull))
            assert_hits_equal(hits, np.stack([rec["t"][sel], rec["u"][sel], rec["v"][sel]], axis=1), rec["tri"][sel])


def random_rays(n, seed, lo=-2.0, hi=2.0):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    # edge cases: axis-parallel directions (inv_direction = +-inf, 0*inf = NaN in the slab test), a zero direction, NaNs
    rays[0, 3:] = (0, 0, -1)
    rays[1, 3:] = (1, 0, 0)
    rays[2, 3:] = (0, -1, 0)
    rays[3, 3:] = (0, 0, 0)
    rays[4, :3] = np.nan
    rays[5, 3:] = (-0.0, 1, 0)
    # nearly axis-parallel: a component that flushes to zero in rcp.approx.ftz, a denormal one, huge finite reciprocals
    rays[6, 3:] = (1e-39, 0.6, 0.8)
    rays[7, 3:] = (0.6, -1e-35, 0.8)
    rays[8, 3:] = (0.6, 0.8, 1e-20)
    rays[9, 3:] = (-1e-25, -1e-25, 1)
    # axis-parallel rays from many origins (the special rows above start wherever the generator put them)
    k = min(len(rays) // 4, 3000)
    axes = np.eye(3, dtype=np.float32)[np.arange(k) % 3] * np.where(np.arange(k) % 2, -1.0, 1.0).astype(np.float32)[:, None]
    rays[10:10 + k, 3:] = axes
    return rays


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("cull", [False, True])
def test_incoherent_rays_bit_exact(rt, oracle_mod, name, cull):
    s, data = gpu_scene(rt, name)
    o = oracle_mod.Oracle(data)
    rays = random_rays(200_000, 11)
    n5, bx, _ = s.tree()
    c = (bx[0, :3] + bx[0, 3:]) / 2
    rays[:, :3] = rays[:, :3] * (bx[0, 3:] - bx[0, :3]).max() / 2 + c
    assert_hits_equal(s.trace_closest(rays, cull), *o.trace(rays, cull))


@pytest.mark.parametrize("kd", [(8, 64), (24, 64), (16, 8)])
def test_synthetic_mesh_bit_exact(rt, oracle_mod, kd):
    """config-5 style random mesh, deeper trees than the reference defaults"""
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=60_000, seed=5, width=320, height=200))
    s = rt.Scene.from_rtsc(data, kd_max_depth=kd[0], kd_max_leaf_size=kd[1])
    o = oracle_mod.Oracle(data, *kd)
    assert_hits_equal(s.trace_primary().reshape(-1), *o.trace(o.primary_rays(), True))
    rays = random_rays(100_000, 3, -1.4, 1.4)
    assert_hits_equal(s.trace_closest(rays, False), *o.trace(rays, False))
    img = s.render_frame()
    oi, oc = o.render(oracle_mod.default_params())
    assert np.array_equal(img.view(np.uint32), oi.view(np.uint32))
    c = s.counters()
    assert (c.primary, c.primary_hits, c.shadow, c.shadow_hits) == tuple(int(x) for x in oc[:4])
    s.close()


def test_empty_and_tiny_batches(rt):
    s, _ = gpu_scene(rt, "hw12_scene4")
    assert len(s.trace_closest(np.zeros((0, 6), np.float32), True)) == 0
    assert len(s.trace_occluded(np.zeros((0, 6), np.float32), np.zeros(0, np.float32))) == 0
    one = s.trace_closest(np.asarray([[0, 0, 0, 0, 0, -1]], np.float32), False)
    assert one.shape == (1,)


# ---- shadow queries -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SCENES)
def test_occluded_equals_oracle(rt, oracle_mod, name):
    """is_occluded incl. the refractive pass-through loop (hw15_scene2, hw11_scene8 have refractive meshes)"""
    s, data = gpu_scene(rt, name)
    o = oracle_mod.Oracle(data)
    # shadow rays as the renderer builds them: from primary hit points towards random "lights"
    prim = o.primary_rays()
    tuv, tri = o.trace(prim, True)
    idx = np.flatnonzero(tri >= 0)[:: max(1, int((tri >= 0).sum()) // 150_000)]
    rng = np.random.default_rng(5)
    p = prim[idx, :3] + tuv[idx, :1] * prim[idx, 3:]
    n5, bx, _ = s.tree()
    lights = rng.uniform(bx[0, :3] - 1, bx[0, 3:] + 1, (len(idx), 3)).astype(np.float32)
    d = lights - p
    r = np.linalg.norm(d, axis=1).astype(np.float32)
    d = (d / r[:, None]).astype(np.float32)
    rays = np.concatenate([p + np.float32(1e-4) * d, d], axis=1).astype(np.float32)
    max_t = r.copy()
    max_t[:7] = (0.0, -1.0, np.nan, np.inf, 3.4e38, 1e-7, 1e-3)
    got = s.trace_occluded(rays, max_t)
    want, _ = o.occluded(rays, max_t)
    assert np.array_equal(got, want)
    if name != "hw12_scene4":                                      # four coplanar quads: nothing can be shadowed
        assert 0 < int(got.sum()) < len(got)


# ---- frames ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,key,depth", CONFIGS)
def test_frame_bit_exact_and_ray_counts(rt, golden, name, key, depth):
    """float frame == the compiled reference's frame bit for bit (sha256), and so are the 8-bit frame and the number of
    closest-hit queries by kind - BASELINE configs 1-4 at spp 1 (+ config 3 at max_ray_depth 10)"""
    g = golden["scenes"][name]["configs"][key]
    s, _ = gpu_scene(rt, name)
    p = rt.default_params(max_ray_depth=depth)
    img = s.render_frame(p)
    assert sha(img) == g["sha256_f32"]
    assert sha(quantise(img)) == g["sha256_rgb8"]
    c = s.counters()
    assert (c.primary, c.primary_hits) == (g["counts"]["cull"], g["counts"]["cull_hit"])
    assert c.shadow + c.secondary == g["counts"]["nocull"]
    assert c.shadow_hits + c.secondary_hits == g["counts"]["nocull_hit"]
    assert c.ms_total > 0 and c.kernel_launches > 0
    assert sha(s.render_frame_rgb8(p)) == g["sha256_rgb8"]          # fused quantise (io/image/ppm.hpp:17-19)


def test_published_golden_image(rt, golden):
    """outputs/refractive_dragon.png of the reference == our 8-bit frame of scenes/hw11/scene8.crtscene, every pixel"""
    s, _ = gpu_scene(rt, "hw11_scene8")
    assert sha(s.render_frame_rgb8()) == golden["published"]["refractive_dragon.png"]["sha256_rgb8"]


@pytest.mark.parametrize("name", ["hw09_scene5", "hw11_scene8"])
def test_ray_counts_by_kind_equal_oracle(rt, oracle_mod, name):
    s, data = gpu_scene(rt, name, size=(480, 270))
    o = oracle_mod.Oracle(data)
    img = s.render_frame()
    oi, oc = o.render(oracle_mod.default_params())
    c = s.counters()
    assert np.array_equal(img.view(np.uint32), oi.view(np.uint32))
    assert (c.primary, c.primary_hits, c.shadow, c.shadow_hits, c.secondary, c.secondary_hits) == tuple(int(x) for x in oc[:6])


def test_tiles_and_sample_slices_compose(rt):
    """multi-GPU sharding contract: a tile render writes exactly its rectangle of the full frame; raw sample-slice sums add
    up to the full multi-sample frame"""
    s, _ = gpu_scene(rt, "hw15_scene2", size=(200, 120))
    kw = dict(samples_per_pixel=4, diffuse_reflection_ray_count=1, max_ray_depth=3)
    full = s.render_frame(rt.default_params(**kw))
    tiled = np.full_like(full, -7.0)
    for x0, y0, x1, y1 in ((0, 0, 77, 120), (77, 0, 200, 33), (77, 33, 200, 120)):
        s.render_frame(rt.default_params(x0=x0, y0=y0, x1=x1, y1=y1, **kw), out=tiled)
    assert np.array_equal(full.view(np.uint32), tiled.view(np.uint32))
    acc = np.zeros_like(full)
    for first, n in ((0, 1), (1, 2), (3, 1)):
        acc += s.render_frame(rt.default_params(samples_per_pixel=n, sample_offset=first, spp_total=4, flags=rt.FLAG_RAW_SUM,
                                                diffuse_reflection_ray_count=1, max_ray_depth=3))
    np.testing.assert_allclose(acc / np.float32(4), full, rtol=1e-6, atol=2e-6)


@pytest.mark.parametrize("name,kw", [
    ("hw15_scene2", dict(spp=4, gi_rays=2, max_ray_depth=3)),
    ("hw09_scene5", dict(spp=3, gi_rays=1, max_ray_depth=4)),
    ("hw11_scene8", dict(spp=2, gi_rays=1, max_ray_depth=5)),
    ("hw12_scene4", dict(spp=8, gi_rays=1, max_ray_depth=5)),
])
def test_gi_multisample_matches_oracle_philox(rt, oracle_mod, name, kw):
    s, data = gpu_scene(rt, name, size=(240, 136))
    o = oracle_mod.Oracle(data)
    img = s.render_frame(rt.default_params(samples_per_pixel=kw["spp"], diffuse_reflection_ray_count=kw["gi_rays"],
                                           max_ray_depth=kw["max_ray_depth"]))
    oi, oc = o.render(oracle_mod.default_params(**kw))
    c = s.counters()
    same = (quantise(img) == quantise(oi)).all(axis=2).mean()
    assert same >= 0.99, same
    assert psnr8(img, oi) >= 45.0
    assert c.primary == int(oc[0]) and abs(int(c.primary_hits) - int(oc[1])) <= 2
    total, ototal = c.shadow + c.secondary, int(oc[2] + oc[4])
    assert abs(int(total) - ototal) <= 1e-3 * ototal


@pytest.mark.parametrize("flags", ["exact", "accelerated"])
def test_gi_128spp_frame_within_the_reference_bound(rt, flags):
    """north star: "multi-bounce GI images must be within a stated RMSE/PSNR bound at matched spp".  The frame is the
    reference's own published GI render - scenes/hw15/scene2.crtscene at 1080x1080, 128 spp, max_ray_depth 5, 1 GI ray
    (outputs/gi_128spp_5_1.png, README.md:46-51) - and the CUDA frame is compared with BOTH reference renders the fixture holds
    (run A: the published image; run B: the unmodified reference compiled into oracle/_ref, tests/golden/make_gi_fixtures.py).
    The reference is not reproducible run to run there (per-thread minstd streams, dynamic tiles), so the bound is the one
    SURVEY.md section 8d states from the measured floor: PSNR >= PSNR(run A, run B) - 1 dB (floor 32.955 dB => >= 31.955),
    8x8 box-filtered PSNR >= 48 dB, every channel mean within 0.5 %."""
    from .helpers import stats
    z = np.load(os.path.join(HERE, "golden", "gi_hw15_scene2_1080_s128d5g1.npz"))
    a, b = z["published_rgb8"], z["ref_rgb8"]
    floor = stats.psnr_u8(a, b)
    assert abs(floor - float(z["floor_psnr"])) < 1e-9 and 32.5 < floor < 33.5
    s, _ = gpu_scene(rt, "hw15_scene2", size=(1080, 1080))
    img = s.render_frame(rt.default_params(samples_per_pixel=128, diffuse_reflection_ray_count=1, max_ray_depth=5,
                                           flags=rt.FLAG_ORDERED if flags == "accelerated" else 0))
    q = quantise(img)
    for ref in (a, b):
        assert stats.psnr_u8(q, ref) >= floor - 1.0
        assert stats.psnr_box(q, ref, 8) >= 48.0
        assert np.all(np.abs(stats.channel_means(q) / stats.channel_means(ref) - 1.0) <= 0.005)
    assert np.all(np.abs(stats.channel_means(img) / z["ref_mean_f32"] - 1.0) <= 0.005)      # float frames, before quantisation


@pytest.mark.parametrize("flags", ["exact", "accelerated"])
def test_config4_frame_equals_the_published_textures_png(rt, flags):
    """outputs/textures.png (README.md:64-65) == scenes/hw12/scene4.crtscene at spp 1: every pixel, the bitmap
    quadrant included (tests/conftest.py check_textures_png)"""
    tex = np.load(os.path.join(HERE, "golden", "textures_png.npz"))["rgb8"]
    s, _ = gpu_scene(rt, "hw12_scene4")
    check_textures_png(s.render_frame_rgb8(rt.default_params(flags=rt.FLAG_ORDERED if flags == "accelerated" else 0)), tex)


def test_textures_exact(rt, oracle_mod):
    """all four texture kinds (albedo / edges / checker / bitmap), multi-sample: jitter is Philox on both sides, no
    transcendental is involved, so the frame is bit-exact"""
    s, data = gpu_scene(rt, "hw12_scene4", size=(480, 270))
    o = oracle_mod.Oracle(data)
    img = s.render_frame(rt.default_params(samples_per_pixel=8))
    oi, _ = o.render(oracle_mod.default_params(spp=8))
    assert np.array_equal(img.view(np.uint32), oi.view(np.uint32))


def test_sparse_level0_equals_dense_frames(rt):
    """accelerated mode: only the hits of the camera rays become level-0 entries (k_stream_primary_sparse); misses write
    their pixel at once (one-sample passes) or are added by k_accumulate (multi-sample passes).  Same float frame and counts
    as the reference-order mode (dense level 0) for odd frame sizes, tile rectangles that cut the 8x4 pixel tiles, raw sums
    (no divide), GI keys and multi-sample passes."""
    for name, size in (("hw09_scene5", (203, 117)), ("hw15_scene2", (97, 61))):
        s, _ = gpu_scene(rt, name, size=size)
        for kw in (dict(), dict(diffuse_reflection_ray_count=2, max_ray_depth=3), dict(flags=rt.FLAG_RAW_SUM, spp_total=3, sample_offset=1),
                   dict(samples_per_pixel=2), dict(samples_per_pixel=3, diffuse_reflection_ray_count=1, max_ray_depth=2),
                   dict(samples_per_pixel=2, sample_offset=2, spp_total=5, flags=rt.FLAG_RAW_SUM)):
            flags = kw.pop("flags", 0)
            want = s.render_frame(rt.default_params(flags=flags, **kw))
            cw = s.counters()
            got = s.render_frame(rt.default_params(flags=flags | rt.FLAG_ORDERED, **kw))
            cg = s.counters()
            assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), (name, kw)
            assert (cw.primary, cw.primary_hits, cw.shadow, cw.secondary) == (cg.primary, cg.primary_hits, cg.shadow, cg.secondary)
            tiled = np.full_like(want, -7.0)
            W, H = size
            for x0, y0, x1, y1 in ((0, 0, 77, H), (77, 0, W, 33), (77, 33, W, H)):
                s.render_frame(rt.default_params(x0=x0, y0=y0, x1=x1, y1=y1, flags=flags | rt.FLAG_ORDERED, **kw), out=tiled)
            assert np.array_equal(want.view(np.uint32), tiled.view(np.uint32)), (name, kw, "tiles")


def test_sparse_level0_in_later_one_sample_passes(rt):
    """a frame whose samples do not fit one pass: 1920x1080 at 17 spp renders a 16-sample pass and then a ONE-sample pass that
    adds to the framebuffer (sparse level 0 through k_accumulate, not the fused first-pass path)"""
    s, _ = gpu_scene(rt, "hw09_scene5")
    want = s.render_frame(rt.default_params(samples_per_pixel=17))
    cw = s.counters()
    got = s.render_frame(rt.default_params(samples_per_pixel=17, flags=rt.FLAG_ORDERED))
    cg = s.counters()
    assert cg.passes == 2
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    assert (cw.primary, cw.primary_hits, cw.shadow, cw.secondary) == (cg.primary, cg.primary_hits, cg.shadow, cg.secondary)


_MULTIPASS_CHILD = r"""
import importlib, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
rt = importlib.import_module("simd-raytracer_b200")
from tests.conftest import resized, scene_bytes
out = {}
for name, size, kw in (("hw15_scene2", (200, 120), dict(samples_per_pixel=7, diffuse_reflection_ray_count=2, max_ray_depth=4)),
                       ("hw11_scene8", (160, 96), dict(samples_per_pixel=5, max_ray_depth=6))):
    for flags in (0, rt.FLAG_ORDERED):
        s = rt.Scene.from_rtsc(resized(scene_bytes(name), *size))           # FRESH scene: its pools are grown by overflowing passes
        a = s.render_frame(rt.default_params(flags=flags, **kw))
        ca = s.counters()
        b = s.render_frame(rt.default_params(flags=flags, **kw))            # pools are large enough now: no pass is discarded
        host = np.zeros_like(a)
        t = s.render_frame_begin(rt.default_params(flags=flags, **kw), host)  # the same frame queued as part of a sequence
        s.frame_wait(t)
        cq = s.counters()
        out[f"{name}_{flags}"] = a
        out[f"{name}_{flags}_again"] = b
        out[f"{name}_{flags}_queued"] = host
        out[f"{name}_{flags}_passes"] = np.array([ca.passes, cq.passes, ca.primary, ca.shadow, ca.secondary, cq.primary, cq.shadow, cq.secondary])
        s.close()
np.savez(sys.argv[2], **out)
"""


def test_multi_pass_frames_queue_without_a_host_round_trip_per_pass(rt, tmp_path):
    """render.hpp:35-74's sample loop over several passes: a child process with a tiny pass budget (RT_B200_PASS_ENTRIES, read
    once per process) renders multi-sample frames as 5-7 one-sample passes queued back to back - on a FRESH scene, so passes
    outgrow the pools, are discarded on the device, the passes behind them skip themselves and the host resumes from the failed
    pass.  Every such frame (synchronous, repeated, queued as a sequence frame) must equal the one-pass frame of this process
    bit for bit, with the same ray counts: samples are added in the same order whatever the pass structure."""
    import subprocess, sys
    from .conftest import REPO
    out = tmp_path / "multipass.npz"
    env = dict(os.environ, RT_B200_PASS_ENTRIES="30000")
    subprocess.run([sys.executable, "-c", _MULTIPASS_CHILD, REPO, str(out)], check=True, env=env, timeout=600)
    z = np.load(out)
    for name, size, kw, n_pass in (("hw15_scene2", (200, 120), dict(samples_per_pixel=7, diffuse_reflection_ray_count=2, max_ray_depth=4), 7),
                                   ("hw11_scene8", (160, 96), dict(samples_per_pixel=5, max_ray_depth=6), 5)):
        s, _ = gpu_scene(rt, name, size=size)
        for flags in (0, rt.FLAG_ORDERED):
            want = s.render_frame(rt.default_params(flags=flags, **kw))
            c = s.counters()
            assert c.passes == 1 or os.environ.get("RT_B200_PASS_ENTRIES")          # (the suite itself may run with a small budget)
            for tag in ("", "_again", "_queued"):
                assert np.array_equal(want.view(np.uint32), z[f"{name}_{flags}{tag}"].view(np.uint32)), (name, flags, tag)
            p = z[f"{name}_{flags}_passes"]
            assert p[0] == n_pass and p[1] == n_pass
            assert tuple(p[2:5]) == (c.primary, c.shadow, c.secondary) and tuple(p[5:8]) == (c.primary, c.shadow, c.secondary)


_OVERFLOW_CHILD = r"""
import importlib, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
rt = importlib.import_module("simd-raytracer_b200")
from tests.conftest import resized, scene_bytes
out = {}
for name, size, kw in (("hw15_scene2", (200, 120), dict(samples_per_pixel=3, diffuse_reflection_ray_count=2, max_ray_depth=4)),
                       ("hw11_scene8", (240, 136), dict(max_ray_depth=8)), ("hw09_scene5", (240, 136), dict())):
    s = rt.Scene.from_rtsc(resized(scene_bytes(name), *size))
    for k in range(4):                                  # the host doubles the rows between frames while queries overflow
        out[f"{name}_{k}"] = s.render_frame(rt.default_params(flags=rt.FLAG_ORDERED, **kw))
        c = s.counters()
        out[f"{name}_{k}_counts"] = np.array([c.primary, c.primary_hits, c.shadow, c.secondary])
    s.close()
np.savez(sys.argv[2], **out)
"""


def test_queries_that_outgrow_the_shared_stack_are_answered_exactly(rt, tmp_path):
    """The four-wide stream kernels keep a few traversal-stack entries per lane in shared memory; a query that needs more ends
    with the KD_OVERFLOW mark and is answered by the reference-order traversal, and the host doubles the rows for the next frame
    while more than one query in 2^14 overflows (csrc/rt_stream.cuh).  A child process that starts every scene with TWO rows
    (RT_B200_STACK_ROWS) overflows on most queries of its first frames: every frame, from the first (2 rows) to the fourth (16),
    must be the frame of this process bit for bit, with the same ray counts."""
    import subprocess, sys
    from .conftest import REPO
    out = tmp_path / "overflow.npz"
    subprocess.run([sys.executable, "-c", _OVERFLOW_CHILD, REPO, str(out)], check=True, env=dict(os.environ, RT_B200_STACK_ROWS="2"), timeout=600)
    z = np.load(out)
    for name, size, kw in (("hw15_scene2", (200, 120), dict(samples_per_pixel=3, diffuse_reflection_ray_count=2, max_ray_depth=4)),
                           ("hw11_scene8", (240, 136), dict(max_ray_depth=8)), ("hw09_scene5", (240, 136), dict())):
        s, _ = gpu_scene(rt, name, size=size)
        want = s.render_frame(rt.default_params(flags=rt.FLAG_ORDERED, **kw))
        c = s.counters()
        for k in range(4):
            assert np.array_equal(want.view(np.uint32), z[f"{name}_{k}"].view(np.uint32)), (name, k)
            assert tuple(z[f"{name}_{k}_counts"]) == (c.primary, c.primary_hits, c.shadow, c.secondary)


def test_tile_culling_is_conservative(rt):
    """k_tile_cull finishes 8x4 pixel tiles whose camera rays cannot reach the scene's root box.  A small constant-colour
    quad seen by rotated / sheared cameras from many positions (box in a corner of the frame, partly off-screen, behind the
    camera, camera on the box's faces): every pixel the exact primary query reports as a hit must carry the quad's colour, every
    other pixel the background - culling a tile one of whose rays hits would leave a background pixel there."""
    rng = np.random.default_rng(7)
    v = np.asarray([[-0.3, -0.2, 0.0], [0.3, -0.2, 0.0], [0.3, 0.2, 0.05], [-0.3, 0.2, 0.05]], np.float32)
    tri = np.asarray([[0, 1, 2], [0, 2, 3], [2, 1, 0], [3, 2, 0]], np.uint32)          # both windings: culling ON still sees a face
    bg, col = (0.125, 0.25, 0.5), (0.9, 0.8, 0.7)
    checked = hits_seen = 0
    for k in range(24):
        ang = rng.uniform(-1.2, 1.2, 3)
        cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
        R = (np.asarray([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]) @ np.asarray([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @
             np.asarray([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]))
        if k % 5 == 4:
            R = R @ np.asarray([[1.0, 0.3, 0.0], [0.0, 0.8, 0.0], [0.1, 0.0, 1.2]])     # not a rotation: the argument is affine only
        pos = rng.uniform(-1.5, 1.5, 3) if k % 6 else np.asarray([0.3, rng.uniform(-0.2, 0.2), 0.02])   # k % 6 == 0: on the box's face
        size = [(64, 40), (203, 117), (97, 61)][k % 3]
        s = rt.Scene.from_arrays(width=size[0], height=size[1], background=bg, camera_position=tuple(pos), camera_matrix=tuple(R.reshape(-1)),
                                 materials=[dict(kind=3, albedo=col)], meshes=[(0, v, None, tri)])
        for fov in (40.0, 90.0, 140.0):
            p = rt.default_params(flags=rt.FLAG_ORDERED, fov_degrees=fov)
            img = s.render_frame(p)
            hit = s.trace_primary(rt.default_params(fov_degrees=fov))["tri"].reshape(size[1], size[0]) >= 0
            want = np.where(hit[..., None], np.asarray(col, np.float32), np.asarray(bg, np.float32))
            assert np.array_equal(img, want), (k, fov)
            c = s.counters()
            assert (c.primary, c.primary_hits) == (size[0] * size[1], int(hit.sum()))
            checked += 1; hits_seen += int(hit.any())
        s.close()
    assert hits_seen >= checked // 4          # the quad is on screen often enough for the test to mean something


def test_depth_zero_and_empty_scene(rt):
    s, _ = gpu_scene(rt, "hw12_scene4", size=(64, 40))
    img = s.render_frame(rt.default_params(max_ray_depth=0))          # every hit returns the background (render.hpp:138)
    assert np.all(img == img[0, 0])
    e = rt.Scene.from_arrays(width=33, height=17, background=(0.25, 0.5, 0.75), materials=[dict(kind=0, albedo=(1, 1, 1))])
    img = e.render_frame()
    assert np.all(img == np.asarray([0.25, 0.5, 0.75], np.float32))
    c = e.counters()
    assert (c.primary, c.primary_hits) == (33 * 17, 0)
    assert np.all(e.trace_primary()["tri"] == -1)
    e.close()


# ---- non-exact modes -----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["hw09_scene5", "hw11_scene8", "hw15_scene2"])
def test_fast_mode_within_north_star_tolerance(rt, oracle_mod, name):
    s, data = gpu_scene(rt, name)
    o = oracle_mod.Oracle(data)
    hits = s.trace_primary(rt.default_params(flags=rt.FLAG_FAST_MATH)).reshape(-1)
    tuv, tri = o.trace(o.primary_rays(), True)
    agree = (hits["tri"] == tri).mean()
    assert agree >= 0.9999, agree
    both = (tri >= 0) & (hits["tri"] == tri)
    rel = np.abs(hits["t"][both] - tuv[both, 0]) / np.abs(tuv[both, 0])
    assert np.quantile(rel, 0.9999) <= 1e-5
    img = s.render_frame(rt.default_params(flags=rt.FLAG_FAST_MATH))
    oi, _ = o.render(oracle_mod.default_params())
    assert psnr8(img, oi) >= 50.0


@pytest.mark.parametrize("width", WIDTHS)
@pytest.mark.parametrize("name", SCENES)
def test_accelerated_mode_same_hits(rt, oracle_mod, name, width):
    """RT_FLAG_ORDERED: the backend's own deeper tree, front-to-back (rt_kd8.cuh).  Same arithmetic per triangle, so t/u/v are
    the reference's bits; exact-t ties between different triangles are re-run in reference order -> identical output"""
    s, data = gpu_scene(rt, name, width=width)
    o = oracle_mod.Oracle(data)
    assert s.info.bvh_n_refs == s.info.n_triangles and s.info.bvh_n_nodes >= 1
    assert s.info.accel_width == width and (s.info.bvh4_n_nodes >= 1) == (width == 4)
    hits = s.trace_primary(rt.default_params(flags=rt.FLAG_ORDERED)).reshape(-1)
    assert_hits_equal(hits, *o.trace(o.primary_rays(), True))
    rays = random_rays(200_000, 23)
    n5, bx, _ = s.tree()
    rays[:, :3] = rays[:, :3] * (bx[0, 3:] - bx[0, :3]).max() / 2 + (bx[0, :3] + bx[0, 3:]) / 2
    for cull in (False, True):
        assert_hits_equal(s.trace_closest(rays, cull, flags=rt.FLAG_ORDERED), *o.trace(rays, cull))


@pytest.mark.parametrize("width", WIDTHS)
@pytest.mark.parametrize("name,key,depth", CONFIGS)
def test_accelerated_mode_frames_bit_exact(rt, golden, name, key, depth, width):
    """the accelerated mode renders the same float frame, bit for bit, and issues the same queries; only the shadow-hit
    STATISTIC differs (it does not look for hits beyond the light)"""
    g = golden["scenes"][name]["configs"][key]
    s, _ = gpu_scene(rt, name, width=width)
    img = s.render_frame(rt.default_params(max_ray_depth=depth, flags=rt.FLAG_ORDERED))
    assert sha(img) == g["sha256_f32"]
    c = s.counters()
    assert (c.primary, c.primary_hits) == (g["counts"]["cull"], g["counts"]["cull_hit"])
    assert c.shadow + c.secondary == g["counts"]["nocull"]
    assert c.shadow_hits + c.secondary_hits <= g["counts"]["nocull_hit"]


@pytest.mark.parametrize("width", WIDTHS)
def test_accelerated_mode_occluded_and_synthetic(rt, oracle_mod, width):
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=60_000, seed=5, width=320, height=200))
    s = rt.Scene.from_rtsc(data, kd_max_depth=24, kd_max_leaf_size=64, accel_width=width)
    o = oracle_mod.Oracle(data, 24, 64)
    assert_hits_equal(s.trace_primary(rt.default_params(flags=rt.FLAG_ORDERED)).reshape(-1), *o.trace(o.primary_rays(), True))
    rays = random_rays(100_000, 3, -1.4, 1.4)
    assert_hits_equal(s.trace_closest(rays, False, flags=rt.FLAG_ORDERED), *o.trace(rays, False))
    max_t = np.random.default_rng(4).uniform(0.01, 2.5, len(rays)).astype(np.float32)
    want, _ = o.occluded(rays, max_t)
    assert np.array_equal(s.trace_occluded(rays, max_t, flags=rt.FLAG_ORDERED), want)
    img = s.render_frame(rt.default_params(flags=rt.FLAG_ORDERED))
    oi, _ = o.render(oracle_mod.default_params())
    assert np.array_equal(img.view(np.uint32), oi.view(np.uint32))
    s.close()
    for name in ("hw15_scene2", "hw11_scene8"):                 # refractive pass-through in the shadow loop
        s, data = gpu_scene(rt, name, width=width)
        o = oracle_mod.Oracle(data)
        n5, bx, _ = s.tree()
        rays = random_rays(100_000, 31)
        rays[:, :3] = rays[:, :3] * (bx[0, 3:] - bx[0, :3]).max() / 2 + (bx[0, :3] + bx[0, 3:]) / 2
        max_t = np.random.default_rng(6).uniform(0.01, 6.0, len(rays)).astype(np.float32)
        want, _ = o.occluded(rays, max_t)
        assert np.array_equal(s.trace_occluded(rays, max_t, flags=rt.FLAG_ORDERED), want)


# ---- the hierarchy built on the device (rt_build_opts.accel_build, csrc/rt_lbvh.cuh) ---------------------------------------
@pytest.mark.parametrize("width", WIDTHS)
@pytest.mark.parametrize("name", ["hw09_scene5", "hw11_scene8", "hw15_scene2"])
def test_device_built_hierarchy_same_hits_and_frames(rt, oracle_mod, golden, name, width):
    """a different tree (linear BVH from CUDA kernels) under the same queries: hits, frames and ray counts are the reference's,
    bit for bit; the exported tree covers every triangle exactly once"""
    data = scene_bytes(name)
    s = rt.Scene.from_rtsc(data, accel_width=width, accel_build=rt.ACCEL_BUILD_DEVICE)
    o = oracle_mod.Oracle(data)
    assert s.info.accel_build == rt.ACCEL_BUILD_DEVICE and s.info.accel_width == width
    assert s.info.bvh_n_refs == s.info.n_triangles and s.info.bvh_n_nodes >= 1 and s.info.bvh_depth <= 44
    assert (s.info.bvh4_n_nodes >= 1) == (width == 4) and s.info.bvh4_stack_need <= 136
    nodes16, tris12, _ = s.bvh_layout()
    assert np.array_equal(np.sort(tris12[:, 3]), np.arange(s.info.n_triangles, dtype=np.uint32))
    leaves = nodes16[:, 14:16].reshape(-1)
    assert int(leaves.sum()) == s.info.n_triangles and int(leaves.max()) <= 4
    hits = s.trace_primary(rt.default_params(flags=rt.FLAG_ORDERED)).reshape(-1)
    assert_hits_equal(hits, *o.trace(o.primary_rays(), True))
    rays = random_rays(200_000, 29)
    n5, bx, _ = s.tree()
    rays[:, :3] = rays[:, :3] * (bx[0, 3:] - bx[0, :3]).max() / 2 + (bx[0, :3] + bx[0, 3:]) / 2
    for cull in (False, True):
        assert_hits_equal(s.trace_closest(rays, cull, flags=rt.FLAG_ORDERED), *o.trace(rays, cull))
    max_t = np.random.default_rng(8).uniform(0.01, 6.0, len(rays)).astype(np.float32)
    want, _ = o.occluded(rays, max_t)
    assert np.array_equal(s.trace_occluded(rays, max_t, flags=rt.FLAG_ORDERED), want)
    for cname, key, depth in CONFIGS:
        if cname != name:
            continue
        g = golden["scenes"][name]["configs"][key]
        img = s.render_frame(rt.default_params(max_ray_depth=depth, flags=rt.FLAG_ORDERED))
        assert sha(img) == g["sha256_f32"]
        c = s.counters()
        assert (c.primary, c.primary_hits) == (g["counts"]["cull"], g["counts"]["cull_hit"])
        assert c.shadow + c.secondary == g["counts"]["nocull"]
    s.close()


@pytest.mark.parametrize("build", ["host", "device"])
def test_one_triangle_per_leaf_same_hits_and_frames(rt, oracle_mod, golden, monkeypatch, build):
    """the leaf size scenes beyond L2 are built with (finish_create; forced here on a small scene with RT_B200_BVH_LEAF=1),
    both builders: every leaf holds one triangle, hits and the frame are the reference's"""
    monkeypatch.setenv("RT_B200_BVH_LEAF", "1")
    name = "hw09_scene5"
    data = scene_bytes(name)
    s = rt.Scene.from_rtsc(data, accel_build=rt.ACCEL_BUILD_DEVICE if build == "device" else rt.ACCEL_BUILD_HOST)
    monkeypatch.delenv("RT_B200_BVH_LEAF")
    o = oracle_mod.Oracle(data)
    assert s.info.bvh_leaf_size == 1
    nodes16, tris12, _ = s.bvh_layout()
    cnt = nodes16[:, 14:16].reshape(-1)
    assert set(np.unique(cnt[cnt != 0xFFFFFFFF]).tolist()) <= {0, 1} and int((cnt == 1).sum()) == s.info.n_triangles
    hits = s.trace_primary(rt.default_params(flags=rt.FLAG_ORDERED)).reshape(-1)
    assert_hits_equal(hits, *o.trace(o.primary_rays(), True))
    rays = random_rays(100_000, 31)
    n5, bx, _ = s.tree()
    rays[:, :3] = rays[:, :3] * (bx[0, 3:] - bx[0, :3]).max() / 2 + (bx[0, :3] + bx[0, 3:]) / 2
    assert_hits_equal(s.trace_closest(rays, False, flags=rt.FLAG_ORDERED), *o.trace(rays, False))
    img = s.render_frame(rt.default_params(flags=rt.FLAG_ORDERED))
    assert sha(img) == golden["scenes"][name]["configs"]["s1d5g0"]["sha256_f32"]
    s.close()


def test_device_built_hierarchy_large_mesh_and_fallback(rt, oracle_mod):
    """300 K triangles: the device-built and the host-built scene render the same GI frame bit for bit (Philox-keyed rays) and
    give the oracle's hits; a scene too small for the device builder (<= 16 triangles) is built on the host and says so"""
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=300_000, seed=1234, width=480, height=270))
    kw = dict(samples_per_pixel=2, diffuse_reflection_ray_count=1, max_ray_depth=5, flags=rt.FLAG_ORDERED)
    a = rt.Scene.from_rtsc(data, kd_max_depth=24, kd_max_leaf_size=64)
    fa = a.render_frame(rt.default_params(**kw))
    ca = a.counters()
    a.close()
    b = rt.Scene.from_rtsc(data, kd_max_depth=24, kd_max_leaf_size=64, accel_build=rt.ACCEL_BUILD_DEVICE)
    assert b.info.accel_build == rt.ACCEL_BUILD_DEVICE and b.info.accel_build_seconds < 2.0
    fb = b.render_frame(rt.default_params(**kw))
    cb = b.counters()
    assert np.array_equal(fa.view(np.uint32), fb.view(np.uint32))
    assert (ca.primary, ca.primary_hits, ca.shadow, ca.secondary, ca.secondary_hits) == (cb.primary, cb.primary_hits, cb.shadow, cb.secondary, cb.secondary_hits)
    o = oracle_mod.Oracle(data, 24, 64)
    rays = random_rays(100_000, 3, -1.4, 1.4)
    assert_hits_equal(b.trace_closest(rays, False, flags=rt.FLAG_ORDERED), *o.trace(rays, False))
    b.close()
    small = rt.Scene.from_rtsc(scene_bytes("hw12_scene4"), accel_build=rt.ACCEL_BUILD_DEVICE)
    assert small.info.n_triangles <= 16 and small.info.accel_build == rt.ACCEL_BUILD_HOST
    small.close()


@pytest.mark.parametrize("width", WIDTHS)
def test_config5_shape_synthetic_gi_frame(rt, oracle_mod, width):
    """BASELINE.json configs[4] in miniature: a 300 K-triangle random mesh in the diffuse box, kd<24,64>, GI 1, depth 5.
    Size-independent properties at a size the full-resolution job shares: the two query modes render the SAME float frame
    (bit for bit - the GI rays are Philox-keyed, so both modes trace identical rays) with the same query counts; a crop is
    checked against the oracle's Philox render (cos/sin last-bit differences only); 1-spp deterministic part is bit-exact."""
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=300_000, seed=1234, width=480, height=270))
    s = rt.Scene.from_rtsc(data, kd_max_depth=24, kd_max_leaf_size=64, accel_width=width)
    o = oracle_mod.Oracle(data, 24, 64)
    kw = dict(samples_per_pixel=1, diffuse_reflection_ray_count=1, max_ray_depth=5)
    a = s.render_frame(rt.default_params(**kw))
    ca = s.counters()
    b = s.render_frame(rt.default_params(flags=rt.FLAG_ORDERED, **kw))
    cb = s.counters()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert (ca.primary, ca.primary_hits, ca.shadow, ca.secondary, ca.secondary_hits) == (cb.primary, cb.primary_hits, cb.shadow, cb.secondary, cb.secondary_hits)
    rect = (120, 60, 360, 200)
    oi, oc = o.render(oracle_mod.default_params(spp=1, gi_rays=1, max_ray_depth=5), rect=rect)
    crop = (slice(rect[1], rect[3]), slice(rect[0], rect[2]))
    same = (quantise(a[crop]) == quantise(oi[crop])).all(axis=2).mean()
    assert same >= 0.99, same
    assert psnr8(a[crop], oi[crop]) >= 45.0
    # no GI: the frame is deterministic and must equal the oracle bit for bit in both modes
    oi0, oc0 = o.render(oracle_mod.default_params(), rect=rect)
    for flags in (0, rt.FLAG_ORDERED):
        img = s.render_frame(rt.default_params(flags=flags))
        assert np.array_equal(img[crop].view(np.uint32), oi0[crop].view(np.uint32))
    s.close()


def test_config5_at_full_scale_crops_equal_the_oracle(rt, oracle_mod):
    """BASELINE.json configs[4] at the size bench.py times it: 10,000,000 triangles, 3840x2160, kd<24,64>.  Crops of the 4K frame
    - the centre of the mesh, a corner of the box - must equal the oracle bit for bit in the deterministic configuration (1 spp,
    no GI) in both query modes, and to the GI tolerance (same Philox rays, cos/sin last bits) with GI 1."""
    data = crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=10_000_000, seed=1234, width=3840, height=2160))
    s = rt.Scene.from_rtsc(data, kd_max_depth=24, kd_max_leaf_size=64)
    o = oracle_mod.Oracle(data, 24, 64)
    assert s.info.n_triangles == o.n_tris == 10_000_010 and s.info.n_nodes == o.n_nodes
    img = np.zeros((2160, 3840, 3), np.float32)
    for rect in ((1888, 1048, 1952, 1112), (40, 30, 104, 62)):
        crop = (slice(rect[1], rect[3]), slice(rect[0], rect[2]))
        want, oc = o.render(oracle_mod.default_params(), rect=rect)
        for flags in (rt.FLAG_ORDERED, 0):
            s.render_frame(rt.default_params(x0=rect[0], y0=rect[1], x1=rect[2], y1=rect[3], flags=flags), out=img)
            c = s.counters()
            assert np.array_equal(img[crop].view(np.uint32), want[crop].view(np.uint32)), (rect, flags)
            assert (c.primary, c.primary_hits, c.shadow + c.secondary) == (int(oc[0]), int(oc[1]), int(oc[2] + oc[4]))
        gi, _ = o.render(oracle_mod.default_params(spp=2, gi_rays=1, max_ray_depth=5), rect=rect)
        s.render_frame(rt.default_params(x0=rect[0], y0=rect[1], x1=rect[2], y1=rect[3], samples_per_pixel=2, diffuse_reflection_ray_count=1,
                                         max_ray_depth=5, flags=rt.FLAG_ORDERED), out=img)
        assert (quantise(img[crop]) == quantise(gi[crop])).all(axis=2).mean() >= 0.99
        assert psnr8(img[crop], gi[crop]) >= 45.0
    s.close()


@pytest.mark.parametrize("size", [(320, 180), (322, 181)])
def test_peer_combine_equals_single_gpu_frame(rt, size):
    """N ranks, one sample slice each, combined by the fused peer-memory kernel (rt_peer_combine's three
    steps) = the single-GPU frame at spp = N, bit for bit, float and 8-bit.  The ranks are emulated on this one GPU inside one
    process (rt_peer_group_connect_local; every rank is signalled before any rank reduces); two frames exercise the epochs;
    322x181 has a tail that is not a multiple of four floats."""
    import torch
    s, _ = gpu_scene(rt, "hw11_scene8", size=size)
    world = 4
    st = torch.cuda.current_stream().cuda_stream
    groups = [rt.PeerGroup(world, r, 0, size[0], size[1]) for r in range(world)]
    rt.PeerGroup.connect_local(groups)
    kw = dict(max_ray_depth=3, diffuse_reflection_ray_count=1)
    want = s.render_frame(rt.default_params(samples_per_pixel=world, **kw))
    for _frame in range(2):
        for r, g in enumerate(groups):
            first, count = rt.spp_slice(world, r, world)
            s.render_frame_device(rt.default_params(samples_per_pixel=count, sample_offset=first, spp_total=world,
                                                    flags=rt.FLAG_RAW_SUM, **kw), g.framebuffer, stream=st)
        for g in groups:
            g.signal_ready(st)
        for g in groups:
            g.reduce_resolve(world, stream=st)
        for g in groups:
            g.wait_done(st)
        rgb, rgb8 = groups[0].read_result(st)
        assert np.array_equal(rgb.view(np.uint32), want.view(np.uint32))
        assert np.array_equal(rgb8, quantise(want))
    for g in groups:
        g.close()


@pytest.mark.parametrize("flags_name", ["exact", "accelerated"])
def test_row_bands_in_one_call_tile_the_frame(rt, flags_name):
    """rt_params.band_rows / band_period / band_phase (the reference's bucket rows, render/tile/bucket.hpp:7-21, dealt
    round-robin to N ranks): ONE call per rank renders all of the rank's bands, touches no other row, and the N device frames
    add up to the whole-frame render bit for bit - 1 spp, multi-sample GI in several passes, a frame height that is no multiple
    of the band, and the band-by-band rectangles give the same rows."""
    import torch
    flags = rt.FLAG_ORDERED if flags_name == "accelerated" else 0
    for size, band, world, kw in (((320, 180), 16, 3, dict(max_ray_depth=5)),
                                  ((322, 182), 8, 4, dict(max_ray_depth=3, samples_per_pixel=4, diffuse_reflection_ray_count=1))):
        s, _ = gpu_scene(rt, "hw11_scene8", size=size)
        H, W = size[1], size[0]
        want = s.render_frame(rt.default_params(flags=flags, **kw))
        total = np.zeros((H, W, 3), np.float32)
        for r in range(world):
            fb = torch.full((H, W, 3), -7.0, dtype=torch.float32, device="cuda")
            s.render_frame_device(rt.default_params(flags=flags, band_rows=band, band_period=world, band_phase=r, **kw), fb.data_ptr(),
                                  stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got = fb.cpu().numpy()
            mine = np.zeros(H, bool)
            for y0, y1 in rt.row_bands(H, band, r, world):
                mine[y0:y1] = True
            assert np.all(got[~mine] == -7.0)                                  # rows of other ranks are not touched
            assert np.array_equal(got[mine].view(np.uint32), want[mine].view(np.uint32))
            total[mine] = got[mine]
        assert np.array_equal(total.view(np.uint32), want.view(np.uint32))
    # queued on the scene's own streams (two frames in flight): the bands of two ranks, rendered as two frames of a sequence
    s, _ = gpu_scene(rt, "hw11_scene8", size=(320, 180))
    want = s.render_frame(rt.default_params(flags=flags))
    fbs = [torch.zeros((180, 320, 3), dtype=torch.float32, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()                                     # the frames render on the scene's streams, not torch's
    tickets = [s.render_frame_device_begin(rt.default_params(flags=flags, band_rows=16, band_period=2, band_phase=r), fbs[r].data_ptr()) for r in range(2)]
    for t in tickets:
        s.frame_wait(t)
    got = np.zeros_like(want)
    for r in range(2):
        for y0, y1 in rt.row_bands(180, 16, r, 2):
            got[y0:y1] = fbs[r][y0:y1].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    s, _ = gpu_scene(rt, "hw11_scene8", size=(320, 180))
    fb = torch.zeros((180, 320, 3), dtype=torch.float32, device="cuda")
    for bad in (dict(band_rows=6, band_period=2, band_phase=0), dict(band_rows=8, band_period=0, band_phase=0), dict(band_rows=8, band_period=2, band_phase=2),
                dict(band_rows=8, band_period=2, band_phase=0, y0=8), dict(band_rows=64, band_period=8, band_phase=5)):
        with pytest.raises(rt.RtError) as e:
            s.render_frame_device(rt.default_params(**bad), fb.data_ptr())
        assert e.value.status == rt.RT_ERR_BAD_ARG
    with pytest.raises(rt.RtError):
        s.render_frame(rt.default_params(band_rows=8, band_period=2, band_phase=0))     # host images take rectangles only


def test_peer_wait_is_bounded(rt, monkeypatch):
    """a peer that never signals does not leave the GPU spinning: the waiting kernels give up after RT_B200_PEER_TIMEOUT_MS,
    the stream drains, and the group reports RT_ERR_TIMEOUT (here and at every later call)"""
    import time
    import torch
    size = (64, 36)
    s, _ = gpu_scene(rt, "hw11_scene8", size=size)
    st = torch.cuda.current_stream().cuda_stream
    for gather in ("ce", "kernel"):
        monkeypatch.setenv("RT_B200_PEER_TIMEOUT_MS", "250")
        groups = [rt.PeerGroup(2, r, 0, size[0], size[1]) for r in range(2)]
        rt.PeerGroup.connect_local(groups)
        s.render_frame_device(rt.default_params(samples_per_pixel=1, spp_total=2, flags=rt.FLAG_RAW_SUM), groups[0].framebuffer, stream=st)
        groups[0].signal_ready(st)                          # rank 1 never renders, never signals
        groups[0].reduce_resolve(2, stream=st)
        t0 = time.perf_counter()
        with pytest.raises(rt.RtError) as e:
            groups[0].read_result(st)
        assert e.value.status == rt.RT_ERR_TIMEOUT
        assert 0.2 <= time.perf_counter() - t0 < 5.0
        with pytest.raises(rt.RtError) as e:
            groups[0].wait_done(st)
        assert e.value.status == rt.RT_ERR_TIMEOUT
        for g in groups:
            g.close()
        if gather == "ce":
            break                                           # the gather mode is fixed per process (RT_B200_PEER_GATHER); one pass covers the default
    img = s.render_frame(rt.default_params())               # the device is fine afterwards
    assert np.isfinite(img).all()


def test_peer_combine_into_the_shared_host_frame(rt):
    """rt_peer_host_result_attach: every rank maps and pins the same POSIX shared-memory frame and stores ITS slice of the
    combined frame straight into it (RT_PEER_OUT_HOST_RGB), so N PCIe links carry the frame and nobody downloads it.  Emulated
    ranks in one process (each with its own mapping of the object); three frames alternate the two host slots."""
    import os
    import torch
    size = (322, 181)
    s, _ = gpu_scene(rt, "hw11_scene8", size=size)
    world = 4
    st = torch.cuda.current_stream().cuda_stream
    groups = [rt.PeerGroup(world, r, 0, size[0], size[1]) for r in range(world)]
    rt.PeerGroup.connect_local(groups)
    name = f"/rt_b200_test_{os.getpid()}"
    views = [g.attach_host_result(name, create=(r == 0)) for r, g in enumerate(groups)]
    os.unlink("/dev/shm" + name)                                  # the mappings keep the object alive
    with pytest.raises(rt.RtError):
        groups[0].attach_host_result(name, create=False)          # already attached
    for frame, depth in enumerate((3, 1, 2)):
        kw = dict(max_ray_depth=depth)
        want = s.render_frame(rt.default_params(samples_per_pixel=world, **kw))
        for r, g in enumerate(groups):
            first, count = rt.spp_slice(world, r, world)
            s.render_frame_device(rt.default_params(samples_per_pixel=count, sample_offset=first, spp_total=world,
                                                    flags=rt.FLAG_RAW_SUM, **kw), g.framebuffer, stream=st)
        for g in groups:
            g.signal_ready(st)
        for g in groups:
            g.reduce_resolve(world, rt.PEER_OUT_HOST_RGB | rt.PEER_OUT_RGB, stream=st)
        for g in groups:
            g.wait_done(st)
        torch.cuda.synchronize()
        for v in views:                                           # every process-side mapping sees the whole frame
            assert np.array_equal(v[frame & 1].view(np.uint32), want.view(np.uint32)), frame
        rgb, _ = groups[0].read_result(st)
        assert np.array_equal(rgb.view(np.uint32), want.view(np.uint32))
    del views, v
    for g in groups:
        g.close()


def test_peer_combine_pipelined_over_two_frame_slots(rt):
    """The pipelined use of the peer combine (bench.py --gpus N): frame i is queued on each rank's render stream
    (rt_render_frame_device_begin, no host sync) while frame i-1 is reduced on a second stream; every rank owns two frame
    slots.  Two ranks are emulated on this GPU, each with its own scene handle and streams; the frames differ (seed), and every
    combined frame must equal the single-GPU frame at spp = 2 with that seed, bit for bit."""
    torch = pytest.importorskip("torch")
    size, world, frames = (200, 120), 2, 5
    data = resized(scene_bytes("hw15_scene2"), *size)
    scenes = [rt.Scene.from_rtsc(data) for _ in range(world)]
    kw = dict(max_ray_depth=4, diffuse_reflection_ray_count=1)
    want = [scenes[0].render_frame(rt.default_params(samples_per_pixel=world, seed=100 + f, **kw)) for f in range(frames)]
    for sc in scenes:                                     # warm the pools: the queued frames below must not be re-rendered
        sc.render_frame(rt.default_params(samples_per_pixel=1, spp_total=world, flags=rt.FLAG_RAW_SUM, **kw))
    groups = [rt.PeerGroup(world, r, 0, *size) for r in range(world)]
    rt.PeerGroup.connect_local(groups)
    main = [torch.cuda.Stream() for _ in range(world)]
    side = [torch.cuda.Stream() for _ in range(world)]
    done = [torch.cuda.Event() for _ in range(world)]
    got = [torch.zeros((size[1], size[0], 3), dtype=torch.float32).pin_memory().numpy() for _ in range(frames)]
    tickets = []
    for i in range(frames + 1):
        starts = []
        for r in range(world):
            e = torch.cuda.Event()
            e.record(main[r])
            starts.append(e)
        if i > 0:
            for r in range(world):
                side[r].wait_event(starts[r])
                groups[r].reduce_resolve(world, rt.PEER_OUT_RGB, stream=side[r].cuda_stream)
                groups[r].wait_done(side[r].cuda_stream)
                done[r].record(side[r])
            groups[0].download_result(got[i - 1], stream=side[0].cuda_stream)
        if i < frames:
            for r in range(world):
                first, count = rt.spp_slice(world, r, world)
                p = rt.default_params(samples_per_pixel=count, sample_offset=first, spp_total=world, seed=100 + i,
                                      flags=rt.FLAG_RAW_SUM, **kw)
                tickets.append((r, scenes[r].render_frame_device_begin(p, groups[r].framebuffer, stream=main[r].cuda_stream)))
                groups[r].signal_ready(main[r].cuda_stream)
                if i > 0:
                    main[r].wait_event(done[r])
    torch.cuda.synchronize()
    assert not any(scenes[r].frame_wait(t) for r, t in tickets[-2 * world:])
    for f in range(frames):
        assert np.array_equal(got[f].view(np.uint32), want[f].view(np.uint32)), f
    for g in groups:
        g.close()
    for sc in scenes:
        sc.close()


def test_queued_device_frame_reports_a_rerender(rt):
    """rt_render_frame_device_begin on a fresh scene: the queued attempt outgrows the initial pools, rt_frame_wait renders the
    frame again into the caller's buffer and says so (RT_FRAME_RERENDERED); the second time it does not."""
    torch = pytest.importorskip("torch")
    data = resized(scene_bytes("hw15_scene2"), 192, 108)               # closed box: every diffuse hit spawns three GI children
    ref = rt.Scene.from_rtsc(data)
    p = rt.default_params(max_ray_depth=3, diffuse_reflection_ray_count=3, flags=rt.FLAG_ORDERED)
    want = ref.render_frame(p)
    assert ref.counters().nodes_pool > 3 * 192 * 108                 # more than the initial pool factor (2x) allows
    ref.close()
    fresh = rt.Scene.from_rtsc(data)
    st = torch.cuda.current_stream()
    fb = torch.zeros((108, 192, 3), dtype=torch.float32, device="cuda")
    assert fresh.frame_wait(fresh.render_frame_device_begin(p, fb.data_ptr(), stream=st.cuda_stream)) is True
    assert np.array_equal(fb.cpu().numpy().view(np.uint32), want.view(np.uint32))
    fb.zero_()
    assert fresh.frame_wait(fresh.render_frame_device_begin(p, fb.data_ptr(), stream=st.cuda_stream)) is False
    st.synchronize()
    assert np.array_equal(fb.cpu().numpy().view(np.uint32), want.view(np.uint32))
    fresh.close()


def test_empty_level_skipping_never_changes_a_frame(rt):
    """The host launches only the recursion levels that held rays in the previous pass with the same frame parameters; if a
    skipped level turns out to be needed (GI paths of another sample slice reach deeper) the device flags the pass and it is
    rendered again with every level.  Frames from a scene with a warm hint must equal frames from a fresh scene."""
    data = resized(scene_bytes("hw15_scene2"), 48, 48)
    warm = rt.Scene.from_rtsc(data)
    for mode in (0, rt.FLAG_ORDERED):
        for depth in (2, 6):
            for off in (0, 5, 11, 3, 0):
                p = rt.default_params(samples_per_pixel=1, sample_offset=off, spp_total=16, max_ray_depth=depth,
                                      diffuse_reflection_ray_count=1, flags=mode | rt.FLAG_RAW_SUM)
                got = warm.render_frame(p)
                cw = warm.counters()
                fresh = rt.Scene.from_rtsc(data)
                want = fresh.render_frame(p)
                cf = fresh.counters()
                fresh.close()
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (mode, depth, off)
                assert (cw.primary, cw.shadow, cw.secondary, cw.secondary_hits) == (cf.primary, cf.shadow, cf.secondary, cf.secondary_hits)
    warm.close()


def test_frame_sequence_equals_single_frames(rt):
    """rt_render_frame_begin / rt_frame_wait (a caller looping over render_frame, src/main.cpp:13-25): frame i's download runs
    behind frame i+1's render from one of two device frames.  Every frame of a sequence whose parameters change from frame to
    frame (recursion depth, tile, sample slice, query mode) must equal the frame rt_render_frame returns, bit for bit, also when
    tickets are waited for late or out of order; untouched pixels of a tile frame stay untouched."""
    torch = pytest.importorskip("torch")
    s, _ = gpu_scene(rt, "hw15_scene2", size=(160, 120))
    seq = [rt.default_params(max_ray_depth=d, flags=f, x0=x0, y0=y0, x1=x1, y1=y1, samples_per_pixel=1, sample_offset=so, spp_total=st)
           for d, f, (x0, y0, x1, y1), so, st in [(5, 0, (0, 0, 0, 0), 0, 1), (2, rt.FLAG_ORDERED, (0, 0, 0, 0), 0, 1),
                                                  (5, rt.FLAG_ORDERED, (16, 8, 120, 100), 0, 1), (3, 0, (0, 0, 0, 0), 3, 8),
                                                  (5, rt.FLAG_ORDERED, (0, 0, 0, 0), 0, 1), (1, 0, (40, 0, 160, 60), 0, 1),
                                                  (5, rt.FLAG_ORDERED | rt.FLAG_RAW_SUM, (0, 0, 0, 0), 5, 8)]]
    want = [s.render_frame(p, out=np.full((120, 160, 3), -7.0, np.float32)) for p in seq]
    bufs = [torch.full((120, 160, 3), -7.0, dtype=torch.float32).pin_memory().numpy() for _ in seq]
    # (a) the steady-state pattern: begin i+1, then wait i
    prev = None
    for p, b in zip(seq, bufs):
        t = s.render_frame_begin(p, b)
        if prev is not None:
            s.frame_wait(prev)
        prev = t
    s.frame_wait(prev)
    for i, (b, w) in enumerate(zip(bufs, want)):
        assert np.array_equal(b.view(np.uint32), w.view(np.uint32)), i
    # (b) all frames in flight before the first wait, waits in reverse order
    for b in bufs:
        b.fill(-7.0)
    tickets = [s.render_frame_begin(p, b) for p, b in zip(seq, bufs)]
    assert tickets == list(range(tickets[0], tickets[0] + len(seq)))
    for t in reversed(tickets):
        s.frame_wait(t)
    for i, (b, w) in enumerate(zip(bufs, want)):
        assert np.array_equal(b.view(np.uint32), w.view(np.uint32)), i
    with pytest.raises(rt.RtError):
        s.frame_wait(tickets[-1] + 1)
    # the synchronous entry points still work between sequences
    assert np.array_equal(s.render_frame(seq[0]).view(np.uint32), want[0].view(np.uint32))


def test_frame_sequence_on_a_fresh_scene_outgrows_its_pools(rt):
    """Frames of a sequence are queued without waiting for the device, so a pool overflow (here: the first frames of a fresh
    scene whose GI rays need several times the initial pool) is only seen
    when the frame is waited for; it is then rendered again.  The caller must get the same frames and counters either way."""
    torch = pytest.importorskip("torch")
    data = resized(scene_bytes("hw15_scene2"), 192, 108)               # closed box: every diffuse hit spawns three GI children
    ref = rt.Scene.from_rtsc(data)
    ps = [rt.default_params(max_ray_depth=3, diffuse_reflection_ray_count=3, flags=f) for f in (rt.FLAG_ORDERED, 0, rt.FLAG_ORDERED)]
    want = [ref.render_frame(p) for p in ps]
    cw = ref.counters()
    assert cw.nodes_pool > 3 * 192 * 108                             # more than the initial pool factor (2x) allows
    ref.close()
    fresh = rt.Scene.from_rtsc(data)
    bufs = [torch.zeros((108, 192, 3), dtype=torch.float32).pin_memory().numpy() for _ in ps]
    tickets = [fresh.render_frame_begin(p, b) for p, b in zip(ps, bufs)]
    for t in tickets:
        fresh.frame_wait(t)
    cf = fresh.counters()
    for i, (b, w) in enumerate(zip(bufs, want)):
        assert np.array_equal(b.view(np.uint32), w.view(np.uint32)), i
    assert (cf.primary, cf.shadow, cf.secondary, cf.secondary_hits) == (cw.primary, cw.shadow, cw.secondary, cw.secondary_hits)
    # steady state afterwards: same frames again, now without a re-render
    tickets = [fresh.render_frame_begin(p, b) for p, b in zip(ps, bufs)]
    fresh.frame_wait(tickets[-1])
    for i, (b, w) in enumerate(zip(bufs, want)):
        assert np.array_equal(b.view(np.uint32), w.view(np.uint32)), i
    fresh.close()


# ---- device-pointer entry points, threading ----------------------------------------------------------------------------------------
def test_device_pointer_api_with_torch(rt, oracle_mod):
    torch = pytest.importorskip("torch")
    s, data = gpu_scene(rt, "hw09_scene5", size=(320, 180))
    o = oracle_mod.Oracle(data)
    rays = o.primary_rays()
    d_rays = torch.from_numpy(rays).cuda()
    d_hits = torch.zeros((len(rays), 4), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream()
    s.trace_closest_device(d_rays.data_ptr(), len(rays), True, d_hits.data_ptr(), stream=st.cuda_stream)
    st.synchronize()
    hits = d_hits.cpu().numpy().view(rt.HIT_DTYPE).reshape(-1)
    assert_hits_equal(hits, *o.trace(rays, True))
    fb = torch.zeros((180, 320, 3), dtype=torch.float32, device="cuda")
    s.render_frame_device(rt.default_params(), fb.data_ptr(), stream=st.cuda_stream)
    st.synchronize()
    oi, _ = o.render(oracle_mod.default_params())
    assert np.array_equal(fb.cpu().numpy().view(np.uint32), oi.view(np.uint32))
    out8 = torch.zeros((180, 320, 3), dtype=torch.uint8, device="cuda")
    s.resolve_sum_device(fb.data_ptr(), 1, d_rgb8=out8.data_ptr(), stream=st.cuda_stream)
    st.synchronize()
    assert np.array_equal(out8.cpu().numpy(), quantise(oi))


def test_single_ray_path_is_reentrant(rt, oracle_mod):
    """the accelerator concept's intersect() is called from every tile worker at once (render/render.hpp:93-101)"""
    s, data = gpu_scene(rt, "hw09_scene5", size=(320, 180))
    o = oracle_mod.Oracle(data)
    rays = o.primary_rays()[::37][:800]
    tuv, tri = o.trace(rays, True)
    errors = []

    def worker(k):
        for i in range(k, len(rays), 8):
            h = s.trace_closest(rays[i:i + 1], True)
            if h["tri"][0] != tri[i] or (tri[i] >= 0 and h["t"][0] != tuv[i, 0]):
                errors.append(i)

    th = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors


def test_cpp_adapter_runs_the_reference_main_flow_on_the_device(rt, oracle_mod, golden, tmp_path):
    """include/b200_accel.hpp under the reference's own program flow (src/main.cpp:13-46 with `using A = b200_accel<float>`),
    compiled against the UNMODIFIED reference headers in the build container (tests/helpers/adapter_gpu.cpp, built by
    __graft_entry__.build()) and run HERE on the device: the image<float> render_frame returns, the frames of
    b200_frame_sequence and the 8-bit frame must be the golden frames of the compiled reference, and accel.intersect<bf>
    (render/accel/accel.hpp:8-12) must answer 1,000 rays as the oracle does - scene marshalling through the reference's own
    types (materials, vertex normals, mesh order) included."""
    import json
    import subprocess
    exe = os.path.join(os.path.dirname(HERE), "tests", "helpers", "_bin", "adapter_gpu")
    if not os.path.exists(exe):
        pytest.skip("tests/helpers/_bin/adapter_gpu is built where the reference headers exist (__graft_entry__.build())")
    name = "hw11_scene8"
    g = golden["scenes"][name]["configs"]["s1d5g0"]
    data = scene_bytes(name)
    (tmp_path / "scene.rtsc").write_bytes(data)
    o = oracle_mod.Oracle(data)
    n5, bx, _ = gpu_scene(rt, name)[0].tree()
    rays = random_rays(1000, 77)
    rays[:, :3] = rays[:, :3] * (bx[0, 3:] - bx[0, :3]).max() / 2 + (bx[0, :3] + bx[0, 3:]) / 2
    rays.astype(np.float32).tofile(tmp_path / "rays.bin")
    r = subprocess.run([exe, str(tmp_path / "scene.rtsc"), str(tmp_path / "out.bin"), str(tmp_path / "rays.bin")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    print("adapter timings:", info)
    raw = (tmp_path / "out.bin").read_bytes()
    h, w = np.frombuffer(raw, np.uint32, 2)
    assert (h, w) == (1080, 1920)
    n = int(h) * int(w) * 3
    off = 8
    frames = []
    for _ in range(3):                                           # render_frame, then two frames of the sequence
        frames.append(np.frombuffer(raw, np.float32, n, off)); off += 4 * n
    for f in frames:
        assert sha(f) == g["sha256_f32"]
    rgb8 = np.frombuffer(raw, np.uint8, n, off); off += n
    assert sha(rgb8) == g["sha256_rgb8"] == golden["published"]["refractive_dragon.png"]["sha256_rgb8"]
    rec = np.dtype([("hit", "<u4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("p", "<f4", 3), ("n", "<f4", 3), ("mesh", "<u4")])
    for cull in (False, True):
        got = np.frombuffer(raw, rec, len(rays), off); off += rec.itemsize * len(rays)
        tuv, tri = o.trace(rays, cull)
        hit = tri >= 0
        assert np.array_equal(got["hit"] != 0, hit)
        for k, f in enumerate(("t", "u", "v")):
            assert np.array_equal(got[f][hit].view(np.uint32), tuv[hit, k].view(np.uint32))
        pos = rays[:, :3] + got["t"][:, None] * rays[:, 3:]      # hit.position = origin + t * direction (kd_tree_simd.hpp:254)
        assert np.allclose(got["p"][hit], pos[hit], rtol=0, atol=1e-5)
        assert np.allclose(np.linalg.norm(got["n"][hit], axis=1), 1.0, atol=1e-5)
    assert off == len(raw)
