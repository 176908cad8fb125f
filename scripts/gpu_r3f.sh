#!/bin/bash
# device-built hierarchy: triangles per leaf (default 4; variants in simd-raytracer_b200/variants)
out=gpurun_out; V=$PWD/simd-raytracer_b200/variants
run() { # label, env...
  label=$1; shift
  for cfg in "cfg1:--workload cfg1 --steps 40" "cfg2:--workload cfg2 --steps 40" "cfg3:--workload cfg3 --steps 20" "cfg5_1M:--workload cfg5 --tris 1000000 --steps 4"; do
    c=${cfg%%:*}; a=${cfg#*:}
    env "$@" timeout 400 python bench.py $a --accel-build device --warmup 4 --no-cpu-baseline --ns-tris 0 > $out/r3f_${label}_$c.json 2> $out/r3f_${label}_$c.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/r3f_${label}_$c.json").read().strip().splitlines()[-1]); r=d["rays"]["ms"]; sc=d["scene"]
    print("%-8s %-8s ms/step %8.4f  prim %.3f sec %.3f shad %.3f  nodes %d depth %d need %d build %.3f s"%("$label","$c",d["ms_per_step"],r["ms_primary"],r["ms_secondary"],r["ms_shadow"],sc["bvh_nodes"],sc["bvh_depth"],sc["bvh4_stack_need"],sc["accel_build_s"]))
except Exception as e: print("$label $c failed",e); print(open("$out/r3f_${label}_$c.err").read()[-500:])
PY
  done
}
run leaf4
for f in $V/librt_*.so; do n=$(basename $f .so); n=${n#librt_}; run $n RT_B200_LIB=$f; done
