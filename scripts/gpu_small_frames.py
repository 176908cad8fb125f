#!/usr/bin/env python
"""developer check: small accelerated-mode frames (sparse level 0, tile culling, multi-sample and later one-sample passes,
tile rectangles, frame sequences) compared with the reference-order mode bit for bit (small cases; compute-sanitizer is not available on the GPU pool)"""
import gzip, importlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
rt = importlib.import_module("simd-raytracer_b200")
from tests.helpers import crtscene
for name in ("hw09_scene5", "hw15_scene2", "hw12_scene4"):
    sc = crtscene.from_rtsc_bytes(gzip.open(os.path.join(REPO, "tests/golden/scenes", name + ".rtsc.gz")).read())
    sc.width, sc.height = 203, 117
    s = rt.Scene.from_rtsc(crtscene.to_rtsc_bytes(sc), device=0)
    for kw in (dict(), dict(samples_per_pixel=3), dict(samples_per_pixel=2, diffuse_reflection_ray_count=1, max_ray_depth=2),
               dict(x0=77, y0=33, x1=203, y1=117), dict(flags=rt.FLAG_RAW_SUM, spp_total=3, sample_offset=1)):
        flags = kw.pop("flags", 0)
        a = s.render_frame(rt.default_params(flags=flags | rt.FLAG_ORDERED, **kw))
        b = s.render_frame(rt.default_params(flags=flags, **kw))
        if "x0" not in kw:
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, kw)
    bufs = [np.zeros((117, 203, 3), np.float32) for _ in range(2)]
    p = rt.default_params(flags=rt.FLAG_ORDERED)
    t0 = s.render_frame_begin(p, bufs[0]); t1 = s.render_frame_begin(p, bufs[1])
    s.frame_wait(t0); s.frame_wait(t1)
    assert np.array_equal(bufs[0], bufs[1])
    s.close()
    print(name, "ok", flush=True)
