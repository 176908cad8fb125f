#!/bin/bash
# developer helper: variant libraries (simd-raytracer_b200/variants/) against the default build, cfg2 / cfg3 / cfg1 queued-frame times
out=gpurun_out; tag=${1:-sw3}
run() { local name=$1; shift
  for w in cfg2 cfg3 cfg1; do env "$@" python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline > $out/${tag}_${name}_$w.json 2> $out/${tag}_${name}_$w.err || echo "FAILED $name $w"; done; }
run base X=1
shopt -s nullglob
for v in simd-raytracer_b200/variants/librt_*.so; do n=$(basename $v .so); run ${n#librt_} RT_B200_LIB=$PWD/$v; done
run base2 X=1
python - <<PY
import json,glob,collections
t=collections.defaultdict(dict)
for f in sorted(glob.glob("$out/${tag}_*.json")):
    name, w = f[len("$out/${tag}_"):-5].rsplit("_",1)
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); ms=d["rays"]["ms"]
        t[name][w]=f"{d['ms_per_step']:7.4f} (p {ms['ms_primary']:.3f} s {ms['ms_secondary']:.3f} sh {ms['ms_shadow']:.3f})"
    except Exception as e: t[name][w]="ERR"
for name in t: print(f"{name:8s}", " | ".join(f"{w} {t[name].get(w,'-')}" for w in ("cfg2","cfg3","cfg1")))
PY
