#!/usr/bin/env python
"""developer probe: lane occupancy of the stream kernels (needs a library built with -DRT_STREAM_STATS, RT_B200_LIB=...)"""
import ctypes as C, gzip, importlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
rt = importlib.import_module("simd-raytracer_b200")
names = ["node-step slots", "lanes walking", "lanes parked", "lanes finished", "lanes without a query", "leaf phases", "lanes in leaf phases",
         "refill rounds", "completion phases", "bursts"]
for scene, kw in (("hw09_scene5", {}), ("hw11_scene8", dict(max_ray_depth=10)), ("hw15_scene2", {})):
    data = gzip.open(os.path.join(REPO, "tests/golden/scenes", scene + ".rtsc.gz")).read()
    s = rt.Scene.from_rtsc(data, device=0)
    p = rt.default_params(flags=rt.FLAG_ORDERED, **kw)
    s.render_frame(p)
    out = (C.c_ulonglong * 16)()
    rt.lib.rt_debug_stream_stats(out, 1)
    s.render_frame(p)
    c = s.counters()
    rt.lib.rt_debug_stream_stats(out, 1)
    v = list(out)
    print(scene, "rays", c.primary + c.shadow + c.secondary)
    for n, x in zip(names, v):
        print(f"   {n:24s} {x:12d}")
    if v[0]:
        print(f"   per node-step slot: walking {v[1]/v[0]:.1f} parked {v[2]/v[0]:.1f} finished {v[3]/v[0]:.1f} empty {v[4]/v[0]:.1f};  lanes per leaf phase {v[6]/max(v[5],1):.1f};"
              f" node slots per leaf phase {v[0]/max(v[5],1):.2f}; slots per burst {v[0]/max(v[9],1):.1f}")
    s.close()
