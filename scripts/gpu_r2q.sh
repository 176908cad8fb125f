#!/bin/bash
# row bands in one call: parity test on one GPU, then the N = 2 bench (north-star record by slices and by bands)
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "row_bands or peer or tile" > $out/r2q_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/r2q_pytest.log)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 \
    > $out/r2q_n2.json 2> $out/r2q_n2.err; echo "n2 rc=$?"
python - <<PY
import json
d = json.loads(open("$out/r2q_n2.json").read().strip().splitlines()[-1])
print("N=2 ms/step %.4f value %.1f verify %s" % (d["ms_per_step"], d["value"], d.get("verify")))
ns = d.get("north_star_scaling") or {}
print({k: ns.get(k) for k in ("n1", "sample_slices", "row_bands")})
PY
