#!/usr/bin/env python
"""developer probe: device time of one config-2 frame as the single-GPU job renders it (spp_total = 1: pixel centres) and as one rank
of an N-GPU job renders its slice (1 sample of spp_total = N, Philox jitter, raw sum) - the part of the weak-scaling 'loss' that
is a different workload per rank, not communication"""
import gzip, importlib, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
rt = importlib.import_module("simd-raytracer_b200")
data = gzip.open(os.path.join(REPO, "tests/golden/scenes/hw09_scene5.rtsc.gz")).read()
s = rt.Scene.from_rtsc(data, device=0)
fb = torch.zeros((s.height, s.width, 3), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream()
for label, kw in (("spp_total 1 (N = 1 job)", dict()), ("slice 0 of spp_total 2", dict(spp_total=2, sample_offset=0, flags=rt.FLAG_ORDERED | rt.FLAG_RAW_SUM)),
                  ("slice 1 of spp_total 2", dict(spp_total=2, sample_offset=1, flags=rt.FLAG_ORDERED | rt.FLAG_RAW_SUM)),
                  ("slice 5 of spp_total 8", dict(spp_total=8, sample_offset=5, flags=rt.FLAG_ORDERED | rt.FLAG_RAW_SUM))):
    kw.setdefault("flags", rt.FLAG_ORDERED)
    p = rt.default_params(samples_per_pixel=1, **kw)
    ms = []
    for i in range(60):
        s.frame_wait(s.render_frame_device_begin(p, fb.data_ptr(), stream=st.cuda_stream))
        c = s.counters()
        if i >= 10:
            ms.append(c.ms_total)
    print(f"{label:28s} {np.mean(ms):.4f} ms (min {np.min(ms):.4f})  rays {c.primary + c.shadow + c.secondary}")
