#!/usr/bin/env python
"""developer probe (library built with -DRT_STREAM_STATS, RT_B200_LIB=...): when do the warps of a stream kernel finish?
Uses rt_trace_closest / rt_trace_occluded-free path: renders config-2 frames whose LAST stream kernel is the shadow kernel, and
level-1-only frames (max_ray_depth 1, no lights is not possible) - so the log is read after frames with different depth limits."""
import ctypes as C, gzip, importlib, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
rt = importlib.import_module("simd-raytracer_b200")
NW = 740 * 4
def report(tag):
    out = (C.c_ulonglong * (NW * 4))()
    rt.lib.rt_debug_stream_log(out, NW)
    a = np.frombuffer(out, dtype=np.uint64).reshape(NW, 4).astype(np.int64)
    a = a[a[:, 1] > 0]
    t0 = a[:, 0].min(); start = (a[:, 0] - t0) / 1e3; end = (a[:, 1] - t0) / 1e3
    total = end.max()
    print(f"{tag}: {len(a)} warps, kernel {total:.1f} us, warp start p50 {np.median(start):.1f} max {start.max():.1f} us; queries {a[:,2].sum()}")
    qs = np.percentile(end, [10, 25, 50, 75, 90, 99, 100])
    print("   warp end times us  p10 %.1f p25 %.1f p50 %.1f p75 %.1f p90 %.1f p99 %.1f max %.1f" % tuple(qs))
    for frac in (0.5, 0.6, 0.7, 0.8, 0.9):
        print(f"   warps still running at {int(frac*100)} % of the kernel: {(end > frac * total).sum():5d}   (mean queries of those {a[end > frac*total, 2].mean() if (end > frac*total).any() else 0:.1f}, slots {a[end > frac*total, 3].mean() if (end > frac*total).any() else 0:.0f})")
    print(f"   mean queries per warp {a[:,2].mean():.1f}, mean node-step slots per warp {a[:,3].mean():.0f}, max slots {a[:,3].max()}, mean busy time {np.mean(end-start):.1f} us")
for name, depth in (("hw15_scene2", 5), ("hw11_scene8", 10)):          # configs 1 and 3: the shadow kernel of frames with many levels
    sx = rt.Scene.from_rtsc(gzip.open(os.path.join(REPO, f"tests/golden/scenes/{name}.rtsc.gz")).read(), device=0)
    px = rt.default_params(flags=rt.FLAG_ORDERED, max_ray_depth=depth)
    for _ in range(3): sx.render_frame(px)
    report(f"{name} depth {depth}: shadow kernel")
    sx.close()
data = gzip.open(os.path.join(REPO, "tests/golden/scenes/hw09_scene5.rtsc.gz")).read()
# shadow kernel last
s = rt.Scene.from_rtsc(data, device=0)
p = rt.default_params(flags=rt.FLAG_ORDERED)
for _ in range(3): s.render_frame(p)
report("cfg2 shadow kernel")
s.close()
# no lights: the last stream kernel of the frame is the level-1 trace (max_ray_depth 1: level-1 hits return the background)
from tests.helpers import crtscene
sc = crtscene.from_rtsc_bytes(data)
sc.lights = sc.lights[:0]
s2 = rt.Scene.from_rtsc(crtscene.to_rtsc_bytes(sc), device=0)
p1 = rt.default_params(flags=rt.FLAG_ORDERED, max_ray_depth=1)
for _ in range(3): s2.render_frame(p1)
report("cfg2 without lights, depth 1: level-1 trace kernel")
p0 = rt.default_params(flags=rt.FLAG_ORDERED, max_ray_depth=0)
for _ in range(3): s2.render_frame(p0)
report("cfg2 depth 0: primary kernel")
s2.close()
