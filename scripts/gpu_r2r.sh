#!/bin/bash
# gpurun --gpus 8: the driver's scaling run at N = 8 and N = 4 (config 2 weak scaling + the north-star strong-scaling record)
out=gpurun_out
for n in 8 4; do
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 100 --warmup 5 \
      > $out/r2r_n$n.json 2> $out/r2r_n$n.err; echo "n$n rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$out/r2r_n$n.json").read().strip().splitlines()[-1])
    print("N=$n ms/step %.4f value %.1f e2e %s verify %s" % (d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"), d.get("verify")))
    ns = d.get("north_star_scaling") or {}
    print({k: ns.get(k) for k in ("n1", "sample_slices", "row_bands")})
except Exception as e:
    print("ERR", e); print(open("$out/r2r_n$n.err").read()[-2000:])
PY
done
