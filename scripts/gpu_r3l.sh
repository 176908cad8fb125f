#!/bin/bash
# prefetch of the next-nearest child's node (variants pf1 = L1, pf2 = L2) vs none: configs 2, 3 and 5
out=gpurun_out; V=$PWD/simd-raytracer_b200/variants
run() { label=$1; shift
  for cfg in "cfg2:--workload cfg2 --steps 60" "cfg3:--workload cfg3 --steps 20" "cfg5_1M:--workload cfg5 --tris 1000000 --steps 4" "cfg5_10M:--workload cfg5 --tris 10000000 --accel-build device --steps 4"; do
    c=${cfg%%:*}; a=${cfg#*:}
    env "$@" timeout 400 python bench.py $a --warmup 4 --no-cpu-baseline --ns-tris 0 > $out/r3l_${label}_$c.json 2> $out/r3l_${label}_$c.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/r3l_${label}_$c.json").read().strip().splitlines()[-1]); r=d["rays"]["ms"]
    print("%-8s %-8s ms/step %8.4f  prim %.3f sec %.3f shad %.3f"%("$label","$c",d["ms_per_step"],r["ms_primary"],r["ms_secondary"],r["ms_shadow"]))
except Exception as e: print("$label $c failed",e)
PY
  done
}
run none
for f in $V/librt_*.so; do n=$(basename $f .so); n=${n#librt_}; run $n RT_B200_LIB=$f; done
