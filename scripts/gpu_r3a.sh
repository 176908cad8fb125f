#!/bin/bash
# config 5: how many shared stack rows (the start value; adaptation doubles when more than 1 query in 2^14 overflows)
out=gpurun_out
for tris in 1000000 10000000; do
  for rows in 16 20 24 28 32; do
    RT_B200_VERBOSE=1 RT_B200_STACK_ROWS=$rows timeout 400 python bench.py --workload cfg5 --tris $tris --accel-build device --steps 4 --warmup 4 --no-cpu-baseline --ns-tris 0 \
        > $out/r3a_${tris}_$rows.json 2> $out/r3a_${tris}_$rows.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/r3a_${tris}_$rows.json").read().strip().splitlines()[-1]); r=d["rays"]["ms"]
    print("tris %-9s rows %-3s ms/step %8.4f  prim %.3f sec %.3f shad %.3f"%("$tris","$rows",d["ms_per_step"],r["ms_primary"],r["ms_secondary"],r["ms_shadow"]))
except Exception as e: print("$tris $rows failed",e)
PY
    grep -h "outgrew" $out/r3a_${tris}_$rows.err | sort | uniq -c | head -3
  done
done
