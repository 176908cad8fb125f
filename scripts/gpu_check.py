#!/usr/bin/env python
"""Developer check on a GPU box: CUDA path vs oracle on the four reference scenes (parity + timing)."""
import gzip, hashlib, importlib, json, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
rt = importlib.import_module("simd-raytracer_b200")
from tests.helpers import oracle

def sha(a): return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
def quant(rgb):
    c = np.clip(rgb.astype(np.float32), np.float32(0), np.float32(1)).astype(np.float64)
    return (255.999 * c).astype(np.uint8)

golden = json.load(open(os.path.join(REPO, "tests/golden/golden.json")))
full = "--full" in sys.argv
for name in ["hw09_scene5", "hw15_scene2", "hw11_scene8", "hw12_scene4"]:
    data = gzip.open(os.path.join(REPO, f"tests/golden/scenes/{name}.rtsc.gz")).read()
    s = rt.Scene.from_rtsc(data)
    o = oracle.Oracle(data)
    print(f"== {name} {s.width}x{s.height} tris={s.info.n_triangles} nodes={s.info.n_nodes} packets={s.info.n_packets}", flush=True)
    for flags, tag in ((0, "exact"), (rt.FLAG_ORDERED, "ordered"), (rt.FLAG_FAST_MATH, "fast"), (rt.FLAG_FAST_MATH | rt.FLAG_ORDERED, "fast+ordered")):
        p = rt.default_params(flags=flags)
        hits = s.trace_primary(p).reshape(-1)
        rays = o.primary_rays()
        tuv, tri = o.trace(rays, True)
        same_tri = hits["tri"] == tri
        hit = tri >= 0
        both = hit & (hits["tri"] >= 0)
        rel = np.abs(hits["t"][both] - tuv[both, 0]) / np.maximum(np.abs(tuv[both, 0]), 1e-30)
        exact = same_tri & np.where(hit, (hits["t"] == tuv[:, 0]) & (hits["u"] == tuv[:, 1]) & (hits["v"] == tuv[:, 2]), True)
        print(f"  primary[{tag}]: tri agree {same_tri.mean()*100:.5f}% bit-exact {exact.mean()*100:.5f}% max rel t {rel.max() if len(rel) else 0:.3e} hits {int((hits['tri']>=0).sum())}", flush=True)
    cfgs = [("s1d5g0", dict(samples_per_pixel=1, max_ray_depth=5))]
    if name == "hw11_scene8": cfgs.append(("s1d10g0", dict(samples_per_pixel=1, max_ray_depth=10)))
    for key, kw in cfgs:
        g = golden["scenes"][name]["configs"][key]
        for flags, tag in ((0, "exact"), (rt.FLAG_ORDERED, "ordered"), (rt.FLAG_FAST_MATH | rt.FLAG_ORDERED, "fast+ordered")):
            p = rt.default_params(flags=flags, **kw)
            img = s.render_frame(p)
            c = s.counters()
            for _ in range(3):
                t0 = time.perf_counter(); img = s.render_frame(p); t1 = time.perf_counter()
            c = s.counters()
            ok8 = sha(quant(img)) == g["sha256_rgb8"]; ok32 = sha(img) == g["sha256_f32"]
            cnt = g["counts"]
            print(f"  frame {key}[{tag}]: rgb8 {'OK' if ok8 else 'DIFF'} f32 {'OK' if ok32 else 'DIFF'} | primary {c.primary}/{c.primary_hits} (ref {cnt['cull']}/{cnt['cull_hit']}) "
                  f"other {c.shadow + c.secondary}/{c.shadow_hits + c.secondary_hits} (ref {cnt['nocull']}/{cnt['nocull_hit']}) | "
                  f"ms total {c.ms_total:.3f} prim {c.ms_primary:.3f} sec {c.ms_secondary:.3f} shadow {c.ms_shadow:.3f} shade {c.ms_shade:.3f} resolve {c.ms_resolve:.3f} "
                  f"launches {c.kernel_launches} wall {1e3*(t1-t0):.2f} ms | primary {c.primary/c.ms_primary/1e3:.1f} Mrays/s", flush=True)
            if not ok8 and full:
                oi, _ = o.render(oracle.default_params(spp=kw["samples_per_pixel"], max_ray_depth=kw["max_ray_depth"]))
                d = (quant(img) != quant(oi)).any(axis=2)
                print(f"     differing px vs oracle: {int(d.sum())}", flush=True)
    # GI / multi-sample vs the oracle's Philox render at reduced size
    s.close()
print("done")
