#!/bin/bash
# merged level + shadow launches, shadow jobs first: A/B against separate launches
out=gpurun_out
run() { # label, env...
  label=$1; shift
  for cfg in "cfg1:--workload cfg1 --steps 60" "cfg2:--workload cfg2 --steps 60" "cfg3:--workload cfg3 --steps 20" "cfg5_1M:--workload cfg5 --tris 1000000 --steps 4"; do
    c=${cfg%%:*}; a=${cfg#*:}
    env "$@" timeout 400 python bench.py $a --warmup 4 --no-cpu-baseline --ns-tris 0 > $out/r2t_${label}_$c.json 2> $out/r2t_${label}_$c.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/r2t_${label}_$c.json").read().strip().splitlines()[-1]); r=d["rays"]["ms"]
    print("%-10s %-8s ms/step %8.4f  e2e %8.4f  prim %.3f sec %.3f shad %.3f"%("$label","$c",d["ms_per_step"],d["e2e"]["ms_per_frame"],r["ms_primary"],r["ms_secondary"],r["ms_shadow"]))
except Exception as e: print("$label $c failed",e)
PY
  done
}
run separate RT_B200_MERGE_SHADOW=0
run merged RT_B200_MERGE_SHADOW=1
