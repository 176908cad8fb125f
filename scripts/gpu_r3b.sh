#!/bin/bash
# two frames of a sequence on two streams and two pool sets: GPU parity suite, then e2e with and without the overlap
out=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r3b_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/r3b_pytest.log)"
for ov in 1 0; do
  for cfg in "cfg1:--workload cfg1 --steps 60" "cfg2:--workload cfg2 --steps 60" "cfg3:--workload cfg3 --steps 20" "cfg4:--workload cfg4 --steps 6" "cfg5_1M:--workload cfg5 --tris 1000000 --steps 4"; do
    c=${cfg%%:*}; a=${cfg#*:}
    RT_B200_SEQ_OVERLAP=$ov timeout 400 python bench.py $a --warmup 4 --no-cpu-baseline --ns-tris 0 > $out/r3b_ov${ov}_$c.json 2> $out/r3b_ov${ov}_$c.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/r3b_ov${ov}_$c.json").read().strip().splitlines()[-1]); e=d["e2e"]
    print("overlap %s %-8s ms/step %8.4f  e2e %8.4f ms/frame  (%s)"%("$ov","$c",d["ms_per_step"],e["ms_per_frame"],{k:v for k,v in e.items() if k.startswith("ms_") or "sequence" in k}))
except Exception as e: print("$ov $c failed",e)
PY
  done
done
