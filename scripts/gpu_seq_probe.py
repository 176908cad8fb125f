#!/usr/bin/env python
"""developer probe: where does a pipelined frame's time go (host durations of begin / wait, device ms of the frame)"""
import gzip, importlib, os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
rt = importlib.import_module("simd-raytracer_b200")
data = gzip.open(os.path.join(REPO, "tests/golden/scenes/hw09_scene5.rtsc.gz")).read()
s = rt.Scene.from_rtsc(data, device=0)
p = rt.default_params(flags=rt.FLAG_ORDERED)
bufs = [torch.zeros((s.height, s.width, 3), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
K = 200
for mode in ("sync", "seq", "seq_wait_late"):
    tb = tw = 0.0; dev = 0.0
    for rep in range(2):
        tb = tw = dev = 0.0
        t00 = time.perf_counter()
        prev = None
        for i in range(K):
            t0 = time.perf_counter()
            if mode == "sync":
                s.render_frame(p, out=bufs[i & 1]); t1 = time.perf_counter(); t2 = t1
            else:
                t = s.render_frame_begin(p, bufs[i & 1]); t1 = time.perf_counter()
                if prev is not None: s.frame_wait(prev)
                t2 = time.perf_counter(); prev = t
            tb += t1 - t0; tw += t2 - t1
        if prev is not None: s.frame_wait(prev)
        tot = time.perf_counter() - t00
    c = s.counters()
    print(f"{mode:14s} per frame {1e3*tot/K:.3f} ms  begin/render call {1e3*tb/K:.3f}  wait {1e3*tw/K:.3f}  last frame device ms_total {c.ms_total:.3f} classes p {c.ms_primary:.3f} s {c.ms_secondary:.3f} sh {c.ms_shadow:.3f} sd {c.ms_shade:.3f} r {c.ms_resolve:.3f} sum {c.ms_primary+c.ms_secondary+c.ms_shadow+c.ms_shade+c.ms_resolve:.3f} launches {c.kernel_launches}", flush=True)
# copy alone
d = torch.zeros((s.height, s.width, 3), dtype=torch.float32, device="cuda"); h = torch.from_numpy(bufs[0])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); print(f"D2H alone {1e3*(time.perf_counter()-t0)/50:.3f} ms/frame")
# render alone to a device frame
fb = torch.zeros((s.height, s.width, 3), dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(K): s.render_frame_device(p, fb.data_ptr())
s.counters(); torch.cuda.synchronize(); print(f"render_frame_device alone (host wall) {1e3*(time.perf_counter()-t0)/K:.3f} ms/frame")
