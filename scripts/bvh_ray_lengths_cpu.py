#!/usr/bin/env python
"""developer probe (CPU only): distribution of node visits per query of the shipped BVH traversal over the oracle's recorded
query stream of a frame - how long is the longest query, and which rays are they?  usage: bvh_ray_lengths_cpu.py [scene] [W H]"""
import ctypes as C, gzip, importlib, os, subprocess, sys, tempfile
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
from tests.helpers import oracle, crtscene
from tests.conftest import resized
so = os.path.join(tempfile.mkdtemp(), "libkd8_host.so")
subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", os.path.join(REPO, "tests/helpers/kd8_host.cpp"), "-o", so])
lib = C.CDLL(so)
lib.bvh4_node_count.restype = C.c_uint64
rt = importlib.import_module("simd-raytracer_b200")
scene = sys.argv[1] if len(sys.argv) > 1 else "hw09_scene5"
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (480, 270)
data = resized(gzip.open(os.path.join(REPO, "tests/golden/scenes", scene + ".rtsc.gz")).read(), W, H)
o = oracle.Oracle(data)
rec, _ = o.record_frame(oracle.default_params(spp=1, max_ray_depth=5, gi_rays=0), cap=1 << 24)
lights = crtscene.from_rtsc_bytes(data).lights
rays = np.ascontiguousarray(np.concatenate([rec["o"], rec["d"]], axis=1), np.float32)
cull = rec["cull"] == 1
shadow = np.zeros(len(rec), bool); max_t = np.full(len(rec), np.inf, np.float32)
for L in lights:
    v = L["pos"][None, :] - rec["o"]; r = np.linalg.norm(v, axis=1)
    c = np.linalg.norm(np.cross(v, rec["d"]), axis=1) / np.maximum(r, 1e-9)
    m = (~cull) & (c < 1e-4) & ((v * rec["d"]).sum(1) > 0)
    shadow |= m; max_t[m] = r[m] + 1e-4
s = rt.Scene.from_rtsc(data, device=rt.DEVICE_HOST_ONLY)
nodes, tris, root = s.bvh_layout()
print(f"{scene} {W}x{H}: bvh nodes {s.info.bvh_n_nodes} depth {s.info.bvh_depth}, {len(rec)} queries")
for kind, m, cu, far, ah in (("primary", cull, 1, None, 0), ("secondary", (~cull) & ~shadow, 0, None, 0), ("shadow", shadow, 0, max_t, 1)):
    r = np.ascontiguousarray(rays[m]); n = len(r)
    nv = np.zeros(n, np.uint32); tt = np.zeros(n, np.uint32)
    f = None if far is None else np.ascontiguousarray(far[m])
    lib.bvh_trace_batch_counts(C.c_void_p(nodes.ctypes.data), C.c_void_p(tris.ctypes.data), C.c_void_p(root.ctypes.data), C.c_void_p(r.ctypes.data),
                               C.c_uint64(n), cu, C.c_float(1e-6), None if f is None else C.c_void_p(f.ctypes.data), ah, C.c_void_p(nv.ctypes.data), C.c_void_p(tt.ctypes.data))
    work = nv.astype(np.int64)
    q = np.percentile(work, [50, 90, 99, 99.9, 100])
    big = work > 4 * max(work.mean(), 1)
    print(f"  {kind:9s} n {n:7d}: node visits mean {work.mean():6.1f} p50 {q[0]:.0f} p90 {q[1]:.0f} p99 {q[2]:.0f} p99.9 {q[3]:.0f} max {q[4]:.0f}; "
          f"tri tests mean {tt.mean():5.1f} max {tt.max()}; queries over 4x the mean: {big.sum()} ({100*big.mean():.2f} %) holding {100*work[big].sum()/max(work.sum(),1):.1f} % of the visits")
    # the same queries over the four-wide hierarchy (csrc/rt_bvh4.cuh, collapsed from the two-wide nodes)
    nv4 = np.zeros(n, np.uint32); tuv = np.zeros((n, 3), np.float32); tri = np.zeros(n, np.int32)
    lib.kd8_counters(None, None, 1)
    lib.bvh4_trace_batch(C.c_void_p(nodes.ctypes.data), C.c_uint64(int(s.info.bvh_n_nodes)), C.c_void_p(tris.ctypes.data), C.c_void_p(root.ctypes.data),
                         C.c_void_p(r.ctypes.data), C.c_uint64(n), cu, 0, C.c_float(1e-6), None if f is None else C.c_void_p(f.ctypes.data), ah,
                         C.c_void_p(tuv.ctypes.data), C.c_void_p(tri.ctypes.data), None, C.c_void_p(nv4.ctypes.data))
    a, b = C.c_uint64(0), C.c_uint64(0); lib.kd8_counters(C.byref(a), C.byref(b), 1)
    w4 = (nv4 & 0xFFFF).astype(np.int64); q4 = np.percentile(w4, [50, 90, 99, 99.9, 100])
    sp4 = (nv4 >> 16).astype(np.int64)          # deepest traversal stack of the query (entries)
    print(f"  {'':9s} four-wide : node visits mean {w4.mean():6.1f} p50 {q4[0]:.0f} p90 {q4[1]:.0f} p99 {q4[2]:.0f} p99.9 {q4[3]:.0f} max {q4[4]:.0f}; tri tests mean {b.value/max(n,1):5.1f}")
    print(f"  {'':9s} four-wide stack depth: mean {sp4.mean():.2f} p99 {np.percentile(sp4, 99):.0f} p99.9 {np.percentile(sp4, 99.9):.0f} max {sp4.max()}; queries deeper than 8 / 12 / 16 entries: "
          f"{100 * (sp4 > 8).mean():.4f} % / {100 * (sp4 > 12).mean():.4f} % / {100 * (sp4 > 16).mean():.4f} %")
    for width in (8, 16):      # statistics only: how many dependent node visits would wider nodes need?
        nvw = np.zeros(n, np.uint32)
        lib.bvh_wide_visit_counts(C.c_void_p(nodes.ctypes.data), C.c_uint64(int(s.info.bvh_n_nodes)), width, C.c_void_p(tris.ctypes.data), C.c_void_p(root.ctypes.data),
                                  C.c_void_p(r.ctypes.data), C.c_uint64(n), cu, C.c_float(1e-6), None if f is None else C.c_void_p(f.ctypes.data), ah, C.c_void_p(nvw.ctypes.data))
        ww = nvw.astype(np.int64); qw = np.percentile(ww, [50, 90, 99, 99.9, 100])
        print(f"  {'':9s} {width:2d}-wide   : node visits mean {ww.mean():6.1f} p50 {qw[0]:.0f} p90 {qw[1]:.0f} p99 {qw[2]:.0f} p99.9 {qw[3]:.0f} max {qw[4]:.0f}")
