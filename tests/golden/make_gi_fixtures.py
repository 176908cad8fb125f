#!/usr/bin/env python
"""GI goldens (run HERE, where /root/reference exists; the outputs are committed and are all a test ever reads).

    python tests/golden/make_gi_fixtures.py            # needs `make -C oracle ref` first (s128d5g1 variant)

The reference's GI renders are not reproducible bit for bit - every worker thread owns a minstd_rand seeded 42 and tiles are
claimed dynamically (utils/rand.hpp:16, render/render.hpp:93-101) - so GI parity is a STATISTICAL bound
(SURVEY.md section 8d), and the bound needs reference renders to compare with.  Written to gi_hw15_scene2_1080_s128d5g1.npz:

  published_rgb8   decoded pixels of the reference's own published render outputs/gi_128spp_5_1.png (README.md:46-51:
                   scenes/hw15/scene2.crtscene at 1080x1080, 128 spp, max_ray_depth 5, 1 GI ray) - "run A", the author's machine
  ref_rgb8         the same frame rendered here by the UNMODIFIED reference (oracle/_ref, canonical FP build, BUCKET_TILES, all
                   host threads) - "run B"
  ref_mean_f32     channel means of run B's float frame
  floor            what the two reference runs measure against each other: psnr (full resolution, 8-bit), psnr_box8 (8x8
                   box-filtered), channel means - the run-to-run floor every bound is stated against

Also textures_png.npz: the decoded pixels of outputs/textures.png (README.md:64-65, scenes/hw12/scene4.crtscene at spp 1), whose
albedo / edges / checker quadrants are exact goldens for config 4 (the bitmap quadrant depends on the JPEG decoder).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from tests.helpers import crtscene, refimpl, stats  # noqa: E402

REF = os.environ.get("RT_REFERENCE", "/root/reference")


def main() -> None:
    from PIL import Image
    pub = np.asarray(Image.open(os.path.join(REF, "outputs", "gi_128spp_5_1.png")).convert("RGB"), np.uint8)
    assert pub.shape == (1080, 1080, 3)
    sc = crtscene.load_crtscene(os.path.join(REF, "scenes/hw15/scene2.crtscene"), root=REF)
    sc.width, sc.height = 1080, 1080
    tmp = "/tmp/rt_gi_fixture_tmp.rtsc"
    crtscene.save_rtsc(sc, tmp)
    ref = refimpl.RefImpl(tmp, spp=128, depth=5, gi=1)
    img, sec = ref.render()
    ref.close()
    os.remove(tmp)
    q = refimpl.quantise(img)
    floor = {"psnr": stats.psnr_u8(pub, q), "psnr_box8": stats.psnr_box(pub, q, 8),
             "mean_published": pub.reshape(-1, 3).mean(0).tolist(), "mean_ref": q.reshape(-1, 3).mean(0).tolist()}
    print(f"reference render: {sec:.1f} s; floor between the two reference runs: {floor}")
    np.savez_compressed(os.path.join(HERE, "gi_hw15_scene2_1080_s128d5g1.npz"), published_rgb8=pub, ref_rgb8=q,
                        ref_mean_f32=img.reshape(-1, 3).mean(0).astype(np.float64),
                        floor_psnr=np.float64(floor["psnr"]), floor_psnr_box8=np.float64(floor["psnr_box8"]))
    tex = np.asarray(Image.open(os.path.join(REF, "outputs", "textures.png")).convert("RGB"), np.uint8)
    np.savez_compressed(os.path.join(HERE, "textures_png.npz"), rgb8=tex)
    for f in ("gi_hw15_scene2_1080_s128d5g1.npz", "textures_png.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
