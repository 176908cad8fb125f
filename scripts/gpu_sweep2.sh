#!/bin/bash
# developer helper: variant libraries (simd-raytracer_b200/variants/) on chosen workloads, with the GPU tests for the default build
tag=${1:-sw}; shift
wls=${@:-cfg2 cfg5}
out=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
run() {
  local name=$1; shift
  for w in $wls; do
    extra=""; [ $w = cfg5 ] && extra="--tris 1000000 --steps 6"
    env "$@" python bench.py --workload $w --mode ordered --steps 30 --warmup 3 --no-cpu-baseline $extra > $out/${tag}_${name}_$w.json 2> $out/${tag}_${name}_$w.err || echo "FAILED $name $w"
  done
}
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
        t[name][w]=f"{d['ms_per_step']:8.3f} (p {ms['ms_primary']:.3f} s {ms['ms_secondary']:.3f} sh {ms['ms_shadow']:.3f} sd {ms['ms_shade']:.3f} r {ms['ms_resolve']:.3f})"
    except Exception as e: t[name][w]="ERR"
for name in t:
    print(f"{name:8s}", " | ".join(f"{w} {t[name].get(w,'-')}" for w in "$wls".split()))
PY
