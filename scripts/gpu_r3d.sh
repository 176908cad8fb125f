#!/bin/bash
# node fetch as four 256-bit loads (default) vs seven 128-bit loads (variant ld128): parity subset, then configs 1-3, 5 at 1 M and 10 M
out=gpurun_out; V=$PWD/simd-raytracer_b200/variants
timeout 900 python -m pytest tests -m gpu -x -q -k "frames or hits or config5 or device_built" > $out/r3d_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/r3d_pytest.log)"
run() { # label, env...
  label=$1; shift
  for cfg in "cfg1:--workload cfg1 --steps 60" "cfg2:--workload cfg2 --steps 60" "cfg3:--workload cfg3 --steps 20" "cfg5_1M:--workload cfg5 --tris 1000000 --steps 4" "cfg5_10M:--workload cfg5 --tris 10000000 --accel-build device --steps 4"; do
    c=${cfg%%:*}; a=${cfg#*:}
    env "$@" timeout 400 python bench.py $a --warmup 4 --no-cpu-baseline --ns-tris 0 > $out/r3d_${label}_$c.json 2> $out/r3d_${label}_$c.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/r3d_${label}_$c.json").read().strip().splitlines()[-1]); r=d["rays"]["ms"]
    print("%-10s %-8s ms/step %8.4f  e2e %8.4f  prim %.3f sec %.3f shad %.3f"%("$label","$c",d["ms_per_step"],d["e2e"]["ms_per_frame"],r["ms_primary"],r["ms_secondary"],r["ms_shadow"]))
except Exception as e: print("$label $c failed",e)
PY
  done
}
run default
for f in $V/librt_*.so; do n=$(basename $f .so); n=${n#librt_}; run $n RT_B200_LIB=$f; done
