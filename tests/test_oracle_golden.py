"""The oracle (oracle/rt_oracle.c) against everything the reference pins: its published golden image, the known-answer
frames / query counts of the unmodified reference compiled into oracle/_ref (tests/golden/golden.json, written by
tests/golden/make_fixtures.py), samples of the reference's own query stream (records_*.npz), and the Random123
known-answer vectors for Philox4x32-10."""
from __future__ import annotations

import os

import numpy as np
import pytest

from .conftest import HERE, SCENES, check_textures_png, quantise, resized, scene_bytes, sha

CONFIGS = [(n, "s1d5g0", 5) for n in SCENES] + [("hw11_scene8", "s1d10g0", 10)]


@pytest.mark.parametrize("name,key,depth", CONFIGS)
def test_oracle_frame_equals_reference(oracle_mod, golden, name, key, depth):
    """float frame, 8-bit frame and query counts of the C restatement == the compiled reference's, bit for bit."""
    g = golden["scenes"][name]["configs"][key]
    o = oracle_mod.Oracle(scene_bytes(name))
    assert o.n_nodes == g["n_nodes"]
    assert o.n_tris == golden["scenes"][name]["n_triangles"]
    img, counts = o.render(oracle_mod.default_params(max_ray_depth=depth))
    assert sha(img) == g["sha256_f32"]
    assert sha(quantise(img)) == g["sha256_rgb8"]
    assert int(counts[0]) == g["counts"]["cull"] and int(counts[1]) == g["counts"]["cull_hit"]
    assert int(counts[2] + counts[4]) == g["counts"]["nocull"]
    assert int(counts[3] + counts[5]) == g["counts"]["nocull_hit"]


def test_oracle_reproduces_published_golden(oracle_mod, golden):
    """outputs/refractive_dragon.png == scenes/hw11/scene8.crtscene at the default config (README.md:60-62)."""
    o = oracle_mod.Oracle(scene_bytes("hw11_scene8"))
    img, _ = o.render(oracle_mod.default_params())
    pub = golden["published"]["refractive_dragon.png"]
    assert list(img.shape) == pub["shape"]
    assert sha(quantise(img)) == pub["sha256_rgb8"]


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("tag", ["default", "gi"])
def test_oracle_matches_reference_query_stream(oracle_mod, records, name, tag):
    """ray in -> (t, u, v, triangle) out, on a strided sample of the queries the reference itself issued (primary with
    culling, shadow / reflection / refraction / GI without), GI variant included."""
    rec = records(name)[f"{tag}_records"]
    o = oracle_mod.Oracle(scene_bytes(name))
    rays = np.concatenate([rec["o"], rec["d"]], axis=1)
    for cull in (0, 1):
        sel = rec["cull"] == cull
        if not sel.any():
            continue
        tuv, tri = o.trace(rays[sel], bool(cull))
        assert np.array_equal(tri, rec["tri"][sel])
        hit = tri >= 0
        for k, f in enumerate(("t", "u", "v")):
            assert np.array_equal(tuv[hit, k].view(np.uint32), rec[f][sel][hit].view(np.uint32))


def test_philox_known_answers(oracle_mod):
    """Random123 kat_vectors for philox4x32 with 10 rounds."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in oracle_mod.philox(ctr, key)) == want


def test_oracle_tile_and_slice_consistency(oracle_mod):
    """tiles and sample slices of the oracle compose to the full frame (what the multi-GPU tests lean on)."""
    from .conftest import resized
    data = resized(scene_bytes("hw15_scene2"), 96, 64)
    o = oracle_mod.Oracle(data)
    full, _ = o.render(oracle_mod.default_params(spp=4, gi_rays=1, max_ray_depth=3))
    tiled = np.zeros_like(full)
    for rect in ((0, 0, 50, 64), (50, 0, 96, 30), (50, 30, 96, 64)):
        o.render(oracle_mod.default_params(spp=4, gi_rays=1, max_ray_depth=3), rect=rect, out=tiled)
    assert np.array_equal(full, tiled)
    acc = np.zeros_like(full)
    for first, n in ((0, 1), (1, 3)):
        part, _ = o.render(oracle_mod.default_params(spp=n, gi_rays=1, max_ray_depth=3, sample_offset=first, spp_total=4, raw_sum=1))
        acc += part
    np.testing.assert_allclose(acc / np.float32(4), full, rtol=1e-6, atol=2e-6)


# ---- GI: the oracle's sampler against the reference (render/render.hpp:151-182) ---------------------------
@pytest.mark.parametrize("name", SCENES)
def test_oracle_minstd_gi_equals_reference_single_tile(oracle_mod, golden, name):
    """RO_RNG_MINSTD is the reference's own random sequence (utils/rand.hpp:5-19: one minstd_rand seeded 42, generate_canonical)
    consumed in the reference's order (jitter :43-44, GI directions :151-182).  With ONE worker and one tile (SINGLE_TILE) the
    reference is deterministic, and its 8 spp / GI 1 float frame - recorded by tests/golden/make_fixtures.py from the compiled
    unmodified reference - must be reproduced bit for bit, query count included."""
    g = golden["scenes"][name]["configs"]["small_gi"]
    o = oracle_mod.Oracle(resized(scene_bytes(name), g["width"], g["height"]))
    img, counts = o.render(oracle_mod.default_params(spp=g["spp"], gi_rays=g["gi"], max_ray_depth=g["depth"], rng=oracle_mod.RNG_MINSTD))
    assert sha(img) == g["sha256_f32"]
    assert int(counts[0] + counts[2] + counts[4]) == g["n_records"]


def test_oracle_philox_gi_within_the_reference_run_to_run_floor(oracle_mod):
    """The Philox specification the CUDA path shares with the oracle is a DIFFERENT random sequence than the reference's, so it
    is held to the reference statistically: hw15/scene2 at 1080x1080, 128 spp, depth 5, GI 1 (outputs/gi_128spp_5_1.png,
    README.md:46-51) on the central 200x200 crop (CPU time; the -m gpu test holds the CUDA path to the whole frame).
    Bound: PSNR >= PSNR(reference run A, reference run B) - 1 dB; 8x8 box-filtered PSNR >= the two
    runs' own box-filtered PSNR - 2.76 dB (the whole-frame bound, 48 dB against a floor of 50.76, restated for a crop whose
    floor is lower because it holds only lit, noisy pixels); channel means within 0.5 % - against the published render
    (run A) and against the compiled reference's render (run B)."""
    from .helpers import stats
    z = np.load(os.path.join(HERE, "golden", "gi_hw15_scene2_1080_s128d5g1.npz"))
    y0, y1, x0, x1 = 440, 640, 440, 640
    o = oracle_mod.Oracle(resized(scene_bytes("hw15_scene2"), 1080, 1080))
    img, _ = o.render(oracle_mod.default_params(spp=128, gi_rays=1, max_ray_depth=5), rect=(x0, y0, x1, y1))
    q = quantise(img)[y0:y1, x0:x1]
    a, b = z["published_rgb8"][y0:y1, x0:x1], z["ref_rgb8"][y0:y1, x0:x1]
    floor, floor_box = stats.psnr_u8(a, b), stats.psnr_box(a, b, 8)
    for ref in (a, b):
        assert stats.psnr_u8(q, ref) >= floor - 1.0
        assert stats.psnr_box(q, ref, 8) >= floor_box - 2.76
        assert np.all(np.abs(stats.channel_means(q) / stats.channel_means(ref) - 1.0) <= 0.005)


def test_oracle_reproduces_the_exact_quadrants_of_textures_png(oracle_mod):
    """outputs/textures.png == scenes/hw12/scene4.crtscene at the default config (README.md:64-65).  The whole frame is a
    golden: the albedo, edges and checker quadrants, and the bitmap quadrant as well now that the fixture's texel bytes are decoded
    with the reference decoder's arithmetic (tests/helpers/jpeg_stb.py) - 0 differing pixels."""
    tex = np.load(os.path.join(HERE, "golden", "textures_png.npz"))["rgb8"]
    o = oracle_mod.Oracle(scene_bytes("hw12_scene4"))
    img, _ = o.render(oracle_mod.default_params())
    check_textures_png(quantise(img), tex)
