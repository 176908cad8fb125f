#!/bin/bash
# developer helper: parameter sweep of the accelerated (stream) kernels - variant libraries from simd-raytracer_b200/variants/
# (build.py --define=... --out=...) and the SAH / accel-tree environment knobs; one bench line per combination
tag=${1:-sweep}
out=gpurun_out
if [ "$2" = "tests" ]; then timeout 900 python -m pytest tests -q -m gpu -x > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log; fi
run() {  # name, env assignments...
  local name=$1; shift
  for w in cfg2 cfg3 cfg1 cfg5; do
    extra=""; [ $w = cfg5 ] && extra="--tris 1000000 --steps 4"
    env "$@" python bench.py --workload $w --mode ordered --steps 20 --warmup 3 --no-cpu-baseline $extra > $out/${tag}_${name}_$w.json 2> $out/${tag}_${name}_$w.err || echo "FAILED $name $w"
  done
}
run base X=1
shopt -s nullglob
for v in simd-raytracer_b200/variants/librt_*.so; do n=$(basename $v .so); run ${n#librt_} RT_B200_LIB=$PWD/$v; done
for e in $SWEEP_ENVS; do run $(echo $e | tr -c 'A-Za-z0-9\n' '_') $e; done
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
    print(f"{name:8s}", " | ".join(f"{w} {t[name].get(w,'-')}" for w in ("cfg2","cfg3","cfg1","cfg5")))
PY
