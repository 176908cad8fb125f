#!/usr/bin/env python
"""Generate tests/golden/* from the reference tree (run HERE, where /root/reference exists; outputs are committed).

    python tests/golden/make_fixtures.py            # needs `make -C oracle ref` first

Writes
  scenes/<name>.rtsc.gz    the four BASELINE.json config scenes converted to the flat RTSC layout
                           (tests/helpers/crtscene.py; every number float(double) as io/json/loader.hpp:9-17;
                           dragon.jpg decoded with PIL - see decode_bitmap for the stb_image caveat)
  golden.json              * sha256 of the decoded RGB bytes of the reference's published outputs/*.png
                           * known answers from the UNMODIFIED reference compiled into oracle/_ref (canonical
                             -ffp-contract=off build): sha256 of the 8-bit frame, of the raw float frame, query
                             counts (cull / no-cull, hits), tree sizes, per config
  records_<name>.npz       strided sample of the reference's own closest-hit query stream (ray in, t/u/v/triangle
                             index out), default config and the spp 8 / GI 1 variant, at reduced resolution

Nothing here is read at test time from /root/reference: tests only open the files this script wrote.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from tests.helpers import crtscene, refimpl  # noqa: E402

REF = os.environ.get("RT_REFERENCE", "/root/reference")
SCENES = {"hw15_scene2": "scenes/hw15/scene2.crtscene", "hw09_scene5": "scenes/hw09/scene5.crtscene",
          "hw11_scene8": "scenes/hw11/scene8.crtscene", "hw12_scene4": "scenes/hw12/scene4.crtscene"}
SMALL = {"hw09_scene5": (480, 270), "hw11_scene8": (320, 180), "hw15_scene2": (240, 240), "hw12_scene4": (320, 180)}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    from PIL import Image
    os.makedirs(os.path.join(HERE, "scenes"), exist_ok=True)
    golden: dict = {"published": {}, "scenes": {}}
    for png in ("refractive_dragon.png", "textures.png"):
        g = np.asarray(Image.open(os.path.join(REF, "outputs", png)).convert("RGB"), np.uint8)
        golden["published"][png] = {"sha256_rgb8": sha(g), "shape": list(g.shape)}

    tmp = "/tmp/rt_fixture_tmp.rtsc"
    for name, rel in SCENES.items():
        sc = crtscene.load_crtscene(os.path.join(REF, rel), root=REF)
        crtscene.save_rtsc(sc, os.path.join(HERE, "scenes", name + ".rtsc.gz"))
        crtscene.save_rtsc(sc, tmp)
        entry: dict = {"source": rel, "width": sc.width, "height": sc.height, "n_triangles": sc.n_triangles,
                       "configs": {}}
        variants = [dict(spp=1, depth=5, gi=0)]
        if name == "hw11_scene8":
            variants.append(dict(spp=1, depth=10, gi=0))
        for var in variants:
            ref = refimpl.RefImpl(tmp, **var)
            img, _ = ref.render()
            cnt = ref.count()
            entry["configs"]["s{spp}d{depth}g{gi}".format(**var)] = {
                "sha256_rgb8": sha(refimpl.quantise(img)), "sha256_f32": sha(img), "counts": cnt,
                "n_nodes": ref.n_nodes, "n_packs_w8": ref.n_packs, "kd": [ref.kd_max_depth, ref.kd_max_leaf]}
            print(name, var, cnt)
            ref.close()

        # strided samples of the reference's query stream at reduced resolution
        w, h = SMALL[name]
        sc.width, sc.height = w, h
        crtscene.save_rtsc(sc, tmp)
        out = {}
        for tag, var, cap in (("default", dict(spp=1, depth=5, gi=0), 1 << 24), ("gi", dict(spp=8, depth=5, gi=1), 1 << 26)):
            ref = refimpl.RefImpl(tmp, **var)
            rec, img = ref.record(cap)
            stride = max(1, len(rec) // 6000)
            idx = np.arange(0, len(rec), stride)
            out[f"{tag}_index"] = idx.astype(np.uint32)
            out[f"{tag}_records"] = rec[idx]
            entry["configs"][f"small_{tag}"] = {"width": w, "height": h, "n_records": int(len(rec)),
                                                "sha256_records": sha(rec), "sha256_f32": sha(img), **var}
            ref.close()
        np.savez_compressed(os.path.join(HERE, f"records_{name}.npz"), **out)
        golden["scenes"][name] = entry

    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(golden, fh, indent=1, sort_keys=True)
    os.remove(tmp)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
