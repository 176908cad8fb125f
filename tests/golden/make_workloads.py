#!/usr/bin/env python
"""Writes tests/golden/workloads.json: the reference ALGORITHM's own work per frame on the BASELINE.json configs -
closest-hit queries by kind, node visits (slab tests) and triangle tests by kind - counted by the oracle
(oracle/rt_oracle.c, pinned bit-exact to the compiled reference).  These are the fixed denominators of bench.py's
roofline (bytes/ray = 8*N_nodes + 36*N_tri + 16 (+24 for a ray read from memory),
flop/ray = 24*N_nodes + 45*N_tri); bench.py reads the JSON, it never runs the oracle for them.

    python tests/golden/make_workloads.py [--synthetic N ...]

Each kind also carries "own": the node visits and triangle tests of the SHIPPED accelerated query (the BVH traversal of
simd-raytracer_b200/csrc/rt_bvh.cuh) for the same queries, counted by running that very source on the CPU
(tests/helpers/kd8_host.cpp) over the oracle's recorded query stream; own_bytes = 64 B/node + 48 B/triangle-record + ray/hit
I/O.  Shadow queries are recognised as the non-culling queries that point at a light; they run any-hit with t_far = distance.

--synthetic N: also (re)count the synthetic config-5 workload at N triangles (bench.py --workload cfg5 [--tris N]; 3840x2160,
spp 1, GI 1, depth 5, kd<24,64>).  A full frame is ~63 M queries of ~500 triangle tests each for the CPU oracle, so the counts
are ESTIMATED: every 64th image row is rendered and the totals are multiplied by 64 (entry field "estimated_from").
"""
import gzip, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.helpers import oracle  # noqa: E402

CONFIGS = {
    "cfg1_hw15_scene2": ("hw15_scene2", dict(spp=1, max_ray_depth=5, gi_rays=0)),
    "cfg2_hw09_scene5": ("hw09_scene5", dict(spp=1, max_ray_depth=5, gi_rays=0)),
    "cfg3_hw11_scene8_d10": ("hw11_scene8", dict(spp=1, max_ray_depth=10, gi_rays=0)),
    "cfg3_hw11_scene8_d5": ("hw11_scene8", dict(spp=1, max_ray_depth=5, gi_rays=0)),
    "cfg4_hw12_scene4": ("hw12_scene4", dict(spp=1, max_ray_depth=5, gi_rays=0)),
}
import ctypes as C
import importlib
import subprocess
import tempfile

import numpy as np


def own_counter():
    """the product's BVH traversal source compiled for the host, with its visit counters"""
    repo = os.path.dirname(os.path.dirname(HERE))
    so = os.path.join(tempfile.mkdtemp(), "libkd8_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                           os.path.join(repo, "tests", "helpers", "kd8_host.cpp"), "-o", so])
    lib = C.CDLL(so)
    rt = importlib.import_module("simd-raytracer_b200")

    def count(rtsc, rec, lights, kd=(8, 64)):
        s = rt.Scene.from_rtsc(rtsc, kd_max_depth=kd[0], kd_max_leaf_size=kd[1], device=rt.DEVICE_HOST_ONLY)
        nodes, tris, root = s.bvh_layout()
        rays = np.ascontiguousarray(np.concatenate([rec["o"], rec["d"]], axis=1), np.float32)
        cull = rec["cull"] == 1
        shadow = np.zeros(len(rec), bool)
        max_t = np.full(len(rec), np.inf, np.float32)
        for L in lights:
            v = L["pos"][None, :] - rec["o"]
            r = np.linalg.norm(v, axis=1)
            c = np.linalg.norm(np.cross(v, rec["d"]), axis=1) / np.maximum(r, 1e-9)
            m = (~cull) & (c < 1e-4) & ((v * rec["d"]).sum(1) > 0)
            shadow |= m
            max_t[m] = r[m] + 1e-4
        out = {}
        for kind, m, kw in (("primary", cull, dict(cull=1)), ("secondary", (~cull) & ~shadow, dict(cull=0)),
                            ("shadow", shadow, dict(cull=0, far=max_t, any_hit=1))):
            r = np.ascontiguousarray(rays[m])
            n = len(r)
            tuv = np.zeros((n, 3), np.float32); tri = np.zeros(n, np.int32)
            far = np.ascontiguousarray(kw["far"][m]) if "far" in kw else None
            lib.kd8_counters(None, None, 1)
            lib.bvh_trace_batch(C.c_void_p(nodes.ctypes.data), C.c_void_p(tris.ctypes.data), C.c_void_p(root.ctypes.data),
                                C.c_void_p(r.ctypes.data), C.c_uint64(n), kw["cull"], 0, C.c_float(1e-6),
                                None if far is None else C.c_void_p(far.ctypes.data), kw.get("any_hit", 0),
                                C.c_void_p(tuv.ctypes.data), C.c_void_p(tri.ctypes.data), None)
            a, b = C.c_uint64(0), C.c_uint64(0)
            lib.kd8_counters(C.byref(a), C.byref(b), 1)
            out[kind] = dict(rays=int(n), nodes=int(a.value), tris=int(b.value))
        s.close()
        return out
    return count


out = {}
path = os.path.join(HERE, "workloads.json")
if os.path.exists(path):
    with open(path) as fh:
        out = {k: v for k, v in json.load(fh).items() if k.startswith("cfg5_synthetic_")}      # kept unless recounted below
for key, (scene, kw) in CONFIGS.items():
    rtsc = gzip.open(os.path.join(HERE, "scenes", scene + ".rtsc.gz")).read()
    o = oracle.Oracle(rtsc)
    _, c = o.render(oracle.default_params(**kw))
    c = [int(x) for x in c]
    kinds = {
        "primary": dict(rays=c[0], hits=c[1], nodes=c[6] - c[8] - c[10], tris=c[7] - c[9] - c[11], ray_in_bytes=0),
        "shadow": dict(rays=c[2], hits=c[3], nodes=c[8], tris=c[9], ray_in_bytes=24),
        "secondary": dict(rays=c[4], hits=c[5], nodes=c[10], tris=c[11], ray_in_bytes=24),
    }
    for k in kinds.values():
        k["alg_bytes"] = 8 * k["nodes"] + 36 * k["tris"] + (16 + k["ray_in_bytes"]) * k["rays"]
        k["alg_flop"] = 24 * k["nodes"] + 45 * k["tris"]
    if "--no-own" not in sys.argv:
        from tests.helpers import crtscene  # noqa: E402
        if "count_own" not in globals():
            count_own = own_counter()
        rec, _ = o.record_frame(oracle.default_params(**kw), cap=1 << 24)
        own = count_own(rtsc, rec, crtscene.from_rtsc_bytes(rtsc).lights)
        for name, k in kinds.items():
            w = own[name]
            k["own"] = dict(structure="bvh2 (64 B nodes, 48 B triangle records)", rays_classified=w["rays"], nodes=w["nodes"], tris=w["tris"],
                            own_bytes=64 * w["nodes"] + 48 * w["tris"] + (16 + k["ray_in_bytes"]) * k["rays"])
    out[key] = dict(scene=scene, width=o.width, height=o.height, n_triangles=o.n_tris, kd=[8, 64], **kw, kinds=kinds,
                    alg_bytes=sum(k["alg_bytes"] for k in kinds.values()), alg_flop=sum(k["alg_flop"] for k in kinds.values()),
                    rays=sum(k["rays"] for k in kinds.values()))
    print(key, out[key]["rays"], out[key]["alg_bytes"], out[key]["alg_flop"])

def kinds_of(c):
    kinds = {
        "primary": dict(rays=c[0], hits=c[1], nodes=c[6] - c[8] - c[10], tris=c[7] - c[9] - c[11], ray_in_bytes=0),
        "shadow": dict(rays=c[2], hits=c[3], nodes=c[8], tris=c[9], ray_in_bytes=24),
        "secondary": dict(rays=c[4], hits=c[5], nodes=c[10], tris=c[11], ray_in_bytes=24),
    }
    for k in kinds.values():
        k["alg_bytes"] = 8 * k["nodes"] + 36 * k["tris"] + (16 + k["ray_in_bytes"]) * k["rays"]
        k["alg_flop"] = 24 * k["nodes"] + 45 * k["tris"]
    return kinds


if "--synthetic" in sys.argv:
    import numpy as np
    from tests.helpers import crtscene  # noqa: E402
    for n in (int(a) for a in sys.argv[sys.argv.index("--synthetic") + 1:]):
        W, H, STEP = 3840, 2160, 64
        kw = dict(spp=1, max_ray_depth=5, gi_rays=1)
        o = oracle.Oracle(crtscene.to_rtsc_bytes(crtscene.synthetic_scene(n_tris=n, seed=1234, width=W, height=H)), 24, 64)
        tot = np.zeros(12, np.uint64)
        for y in range(0, H, STEP):
            _, c = o.render(oracle.default_params(**kw), rect=(0, y, W, y + 1))
            tot += c
        c = [int(x) * STEP for x in tot]
        c[0] = W * H                                               # the primary count is known exactly
        kinds = kinds_of(c)
        key = f"cfg5_synthetic_{n}"
        out[key] = dict(scene=f"synthetic_{n}", width=W, height=H, n_triangles=o.n_tris, kd=[24, 64], **kw, kinds=kinds,
                        alg_bytes=sum(k["alg_bytes"] for k in kinds.values()), alg_flop=sum(k["alg_flop"] for k in kinds.values()),
                        rays=sum(k["rays"] for k in kinds.values()),
                        estimated_from=f"every {STEP}th image row by the oracle, totals x{STEP} (primary count exact)")
        print(key, out[key]["rays"], out[key]["alg_bytes"], out[key]["alg_flop"])

with open(path, "w") as fh:
    json.dump(out, fh, indent=1, sort_keys=True)
