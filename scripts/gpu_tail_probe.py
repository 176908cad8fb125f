#!/usr/bin/env python
"""developer probe: how the trace kernels scale with the number of queries per launch (config 2 at 1, 2, 4, 8 samples per pass):
a kernel whose Mrays/s rises with the launch size is paying for ramp-up and tail, not for steady-state traversal"""
import gzip, importlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
rt = importlib.import_module("simd-raytracer_b200")
scenes = sys.argv[1:] or ["hw09_scene5"]
for scene in scenes:
    data = gzip.open(os.path.join(REPO, "tests/golden/scenes", scene + ".rtsc.gz")).read()
    s = rt.Scene.from_rtsc(data, device=0)
    for spp in (1, 2, 4):
        p = rt.default_params(flags=rt.FLAG_ORDERED, samples_per_pixel=spp)
        best = None
        for i in range(8):
            s.render_frame(p)
            c = s.counters()
            v = (c.ms_total, c.ms_primary, c.ms_secondary, c.ms_shadow, c.ms_shade, c.ms_resolve)
            best = v if best is None else tuple(min(a, b) for a, b in zip(best, v))
        print(f"{scene} spp {spp}: total {best[0]:.3f} ms | primary {c.primary/best[1]/1e3:8.0f} Mrays/s ({best[1]:.3f} ms) | secondary {c.secondary/max(best[2],1e-9)/1e3:8.0f} ({best[2]:.3f}) | "
              f"shadow {c.shadow/max(best[3],1e-9)/1e3:8.0f} ({best[3]:.3f}) | shade {best[4]:.3f} resolve {best[5]:.3f} | launches {c.kernel_launches} passes {c.passes}", flush=True)
    s.close()
