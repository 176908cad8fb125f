#!/bin/bash
# gpurun --gpus N (NG=N bash scripts/gpu_r3m.sh): the north-star record at the size BASELINE.json states - config 5 at 10 M triangles, 4K, GI 1, 64 spp in total, strong scaling
out=gpurun_out
timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-4} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NG:-4} --steps 50 --warmup 5 \
    --ns-tris 10000000 --accel-build device > $out/r3m_n${NG:-4}_ns10M.json 2> $out/r3m_n${NG:-4}_ns10M.err; echo "n${NG:-4} rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$out/r3m_n${NG:-4}_ns10M.json").read().strip().splitlines()[-1])
    print("N=${NG:-4} ms/step %.4f value %.1f verify %s" % (d["ms_per_step"], d["value"], d.get("verify")))
    ns = d.get("north_star_scaling") or {}
    print(ns.get("workload")); print({k: ns.get(k) for k in ("n1", "sample_slices", "row_bands", "host_build_s_per_rank")})
except Exception as e:
    print("ERR", e); print(open("$out/r3m_n${NG:-4}_ns10M.err").read()[-2000:])
PY
