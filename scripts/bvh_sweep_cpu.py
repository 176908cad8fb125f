#!/usr/bin/env python
"""developer sweep (CPU only): node visits / leaf visits / triangle tests per ray of the shipped BVH traversal for builder settings
RT_B200_BVH_LEAF x RT_B200_BVH_COST over the oracle's recorded query streams.  usage: bvh_sweep_cpu.py [scene ...]"""
import ctypes as C, gzip, importlib, os, subprocess, sys, tempfile
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
from tests.helpers import oracle, crtscene
so = os.path.join(tempfile.mkdtemp(), "libkd8_host.so")
subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", os.path.join(REPO, "tests/helpers/kd8_host.cpp"), "-o", so])
lib = C.CDLL(so); lib.bvh_leaf_visits.restype = C.c_uint64
scenes = sys.argv[1:] or ["hw09_scene5", "hw11_scene8", "hw15_scene2"]
W, H = 480, 270
for scene in scenes:
    rtsc = crtscene_bytes = None
    data = gzip.open(os.path.join(REPO, "tests/golden/scenes", scene + ".rtsc.gz")).read()
    from tests.conftest import resized
    data = resized(data, W, H if scene != "hw15_scene2" else W)
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
    for leaf in (2, 4, 8):
        for cost in (0.5, 1.0, 2.0, 4.0):
            env = dict(os.environ, RT_B200_BVH_LEAF=str(leaf), RT_B200_BVH_COST=str(cost))
            # the builder reads the environment once per process: count in a child process
            code = f"""
import ctypes as C, importlib, sys, numpy as np
sys.path.insert(0, {REPO!r})
rt = importlib.import_module("simd-raytracer_b200")
lib = C.CDLL({so!r}); lib.bvh_leaf_visits.restype = C.c_uint64
d = np.load({os.path.join(tempfile.gettempdir(), 'sweep_in.npz')!r})
s = rt.Scene.from_rtsc(open({os.path.join(tempfile.gettempdir(), 'sweep_scene.rtsc')!r}, 'rb').read(), device=rt.DEVICE_HOST_ONLY)
nodes, tris, root = s.bvh_layout()
out = []
for kind, m, cu, far, ah in (("primary", d["cull"], 1, None, 0), ("secondary", d["sec"], 0, None, 0), ("shadow", d["shadow"], 0, d["max_t"], 1)):
    r = np.ascontiguousarray(d["rays"][m]); n = len(r)
    tuv = np.zeros((n, 3), np.float32); tri = np.zeros(n, np.int32)
    f = None if far is None else np.ascontiguousarray(far[m])
    lib.kd8_counters(None, None, 1); lib.bvh_leaf_visits(1)
    lib.bvh_trace_batch(C.c_void_p(nodes.ctypes.data), C.c_void_p(tris.ctypes.data), C.c_void_p(root.ctypes.data), C.c_void_p(r.ctypes.data), C.c_uint64(n), cu, 0, C.c_float(1e-6), None if f is None else C.c_void_p(f.ctypes.data), ah, C.c_void_p(tuv.ctypes.data), C.c_void_p(tri.ctypes.data), None)
    a, b = C.c_uint64(0), C.c_uint64(0); lib.kd8_counters(C.byref(a), C.byref(b), 1)
    out.append((kind, n, a.value, lib.bvh_leaf_visits(1), b.value))
print(s.info.bvh_n_nodes, s.info.bvh_depth, out)
"""
            np.savez(os.path.join(tempfile.gettempdir(), "sweep_in.npz"), rays=rays, cull=cull, sec=(~cull) & ~shadow, shadow=shadow, max_t=max_t)
            open(os.path.join(tempfile.gettempdir(), "sweep_scene.rtsc"), "wb").write(data)
            res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
            if res.returncode: print(res.stderr[-800:]); continue
            line = res.stdout.strip().splitlines()[-1]
            nn, dep, out = eval(line.split(" ", 2)[0]), eval(line.split(" ", 2)[1]), eval(line.split(" ", 2)[2])
            tot = 0; desc = []
            for kind, n, a, lv, b in out:
                est = 80 * a + 35 * lv + 45 * b           # rough instruction model: node step, leaf phase entry, triangle test
                tot += est
                desc.append(f"{kind[:3]} n/r {a/max(n,1):5.1f} l/r {lv/max(n,1):5.1f} t/r {b/max(n,1):5.1f}")
            print(f"{scene} leaf {leaf} cost {cost}: nodes {nn} depth {dep} | " + " | ".join(desc) + f" | model {tot/1e6:.1f} M", flush=True)
