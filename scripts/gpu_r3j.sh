#!/bin/bash
# split shadow launches (level 0's jobs on a second stream beside the deeper levels): GPU parity suite, then A/B
out=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r3j_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/r3j_pytest.log)"
for sp in 1 0; do
  for cfg in "cfg1:--workload cfg1 --steps 60" "cfg2:--workload cfg2 --steps 60" "cfg3:--workload cfg3 --steps 20" "cfg5_1M:--workload cfg5 --tris 1000000 --steps 4"; do
    c=${cfg%%:*}; a=${cfg#*:}
    RT_B200_SHADOW_SPLIT=$sp timeout 400 python bench.py $a --warmup 4 --no-cpu-baseline --ns-tris 0 > $out/r3j_sp${sp}_$c.json 2> $out/r3j_sp${sp}_$c.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/r3j_sp${sp}_$c.json").read().strip().splitlines()[-1]); e=d["e2e"]
    print("split %s %-8s ms/step %8.4f  e2e %8.4f ms/frame  float seq %8.4f"%("$sp","$c",d["ms_per_step"],e["ms_per_frame"],e["float_sequence"]["ms_per_frame"]))
except Exception as ex: print("$sp $c failed",ex); print(open("$out/r3j_sp${sp}_$c.err").read()[-800:])
PY
  done
done
