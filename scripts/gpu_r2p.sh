#!/bin/bash
# bounded peer waits: the peer tests on one GPU, then the N = 2 bench (peer combine) on two
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "peer" > $out/r2p_pytest_peer.log 2>&1; echo "pytest peer rc=$? $(tail -1 $out/r2p_pytest_peer.log)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 \
    > $out/r2p_n2.json 2> $out/r2p_n2.err; echo "n2 rc=$?"
python - <<PY
import json
d = json.loads(open("$out/r2p_n2.json").read().strip().splitlines()[-1])
print("N=2 ms/step %.4f value %.1f verify %s" % (d["ms_per_step"], d["value"], d.get("verify")))
ns = d.get("north_star_scaling") or {}
print({k: ns.get(k) for k in ("sample_slices", "row_bands")})
PY
