#!/usr/bin/env python
"""Writes tests/golden/workloads.json: the reference ALGORITHM's own work per frame on the BASELINE.json configs -
closest-hit queries by kind, node visits (slab tests) and triangle tests by kind - counted by the oracle
(oracle/rt_oracle.c, pinned bit-exact to the compiled reference).  These are the fixed denominators of bench.py's
roofline (bytes/ray = 8*N_nodes + 36*N_tri + 16 (+24 for a ray read from memory),
flop/ray = 24*N_nodes + 45*N_tri); bench.py reads the JSON, it never runs the oracle for them.

    python tests/golden/make_workloads.py [--synthetic N ...]

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
out = {}
path = os.path.join(HERE, "workloads.json")
if os.path.exists(path):
    with open(path) as fh:
        out = {k: v for k, v in json.load(fh).items() if k.startswith("cfg5_synthetic_")}      # kept unless recounted below
for key, (scene, kw) in CONFIGS.items():
    o = oracle.Oracle(gzip.open(os.path.join(HERE, "scenes", scene + ".rtsc.gz")).read())
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
