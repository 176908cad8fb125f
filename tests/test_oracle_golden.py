"""The oracle (oracle/rt_oracle.c) against everything the reference pins: its published golden image, the known-answer
frames / query counts of the unmodified reference compiled into oracle/_ref (tests/golden/golden.json, written by
tests/golden/make_fixtures.py), samples of the reference's own query stream (records_*.npz), and the Random123
known-answer vectors for Philox4x32-10."""
from __future__ import annotations

import numpy as np
import pytest

from .conftest import SCENES, quantise, scene_bytes, sha

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
