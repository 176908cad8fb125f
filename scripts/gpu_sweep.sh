#!/bin/bash
# tuning sweep: variant libraries under simd-raytracer_b200/variants against the default build, configs 2, 3 and 5 (1 M triangles)
out=gpurun_out; tag=${1:-sweep}; V=$PWD/simd-raytracer_b200/variants
run() { # label, env...
  label=$1; shift
  for cfg in "cfg2:--workload cfg2 --steps 60" "cfg3:--workload cfg3 --steps 20" "cfg5:--workload cfg5 --tris 1000000 --steps 4"; do
    c=${cfg%%:*}; a=${cfg#*:}
    env "$@" timeout 300 python bench.py $a --warmup 4 --no-cpu-baseline --ns-tris 0 > $out/${tag}_${label}_$c.json 2> $out/${tag}_${label}_$c.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_${label}_$c.json").read().strip().splitlines()[-1]); r=d["rays"]["ms"]
    print("%-10s %-5s ms/step %8.4f  prim %.3f sec %.3f shad %.3f"%("$label","$c",d["ms_per_step"],r["ms_primary"],r["ms_secondary"],r["ms_shadow"]))
except Exception as e: print("$label $c failed",e)
PY
  done
}
run default RT_B200_ACCEL_WIDTH=4
for f in $V/librt_*.so; do n=$(basename $f .so); n=${n#librt_}; run $n RT_B200_ACCEL_WIDTH=4 RT_B200_LIB=$f; done
