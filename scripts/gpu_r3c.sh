#!/bin/bash
# pass budget (level-0 entries per pass): multi-pass frames of configs 4 and 5
out=gpurun_out
for e in 33554432 67108864 134217728; do
  for cfg in "cfg4:--workload cfg4 --steps 6" "cfg5_1M_spp16:--workload cfg5 --tris 1000000 --spp 16 --steps 2"; do
    c=${cfg%%:*}; a=${cfg#*:}
    RT_B200_PASS_ENTRIES=$e timeout 600 python bench.py $a --warmup 3 --no-cpu-baseline --ns-tris 0 > $out/r3c_${e}_$c.json 2> $out/r3c_${e}_$c.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/r3c_${e}_$c.json").read().strip().splitlines()[-1])
    print("entries %-10s %-14s ms/step %9.4f  e2e %9.4f  passes %s device_bytes %s"%("$e","$c",d["ms_per_step"],d["e2e"]["ms_per_frame"],d["rays"]["passes_per_frame"],d["scene"]["device_bytes"]))
except Exception as ex: print("$e $c failed",ex); print(open("$out/r3c_${e}_$c.err").read()[-600:])
PY
  done
done
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
