#!/bin/bash
# triangles per leaf on config 5: host SAH builder (RT_B200_BVH_LEAF) and device builder (variants leaf1 / leaf2)
out=gpurun_out; V=$PWD/simd-raytracer_b200/variants
one() { # label tris env...
  label=$1; tris=$2; shift 2
  env "$@" timeout 600 python bench.py --workload cfg5 --tris $tris --steps 4 --warmup 4 --no-cpu-baseline --ns-tris 0 $EXTRA > $out/r3g_${label}_$tris.json 2> $out/r3g_${label}_$tris.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/r3g_${label}_$tris.json").read().strip().splitlines()[-1]); r=d["rays"]["ms"]; sc=d["scene"]
    print("%-12s tris %-9s ms/step %8.4f  prim %.3f sec %.3f shad %.3f  nodes %d depth %d need %d accel build %.3f s dev bytes %d"%("$label","$tris",d["ms_per_step"],r["ms_primary"],r["ms_secondary"],r["ms_shadow"],sc["bvh_nodes"],sc["bvh_depth"],sc["bvh4_stack_need"],sc["accel_build_s"],sc["device_bytes"]))
except Exception as e: print("$label $tris failed",e); print(open("$out/r3g_${label}_$tris.err").read()[-500:])
PY
}
for tris in 1000000 10000000; do
  EXTRA=""
  for l in 4 2 1; do one host_leaf$l $tris RT_B200_BVH_LEAF=$l; done
  EXTRA="--accel-build device"
  one dev_leaf4 $tris
  for l in 2 1; do one dev_leaf$l $tris RT_B200_LIB=$V/librt_leaf$l.so; done
done
