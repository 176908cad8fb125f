#!/bin/bash
# developer helper (gpurun --gpus N): the default bench at N ranks, as the driver launches it
n=${1:-8}; tag=${2:-scale}
out=gpurun_out
nvidia-smi -L | head -8
for c in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 100 --warmup 5 --combine $c > $out/${tag}_n${n}_$c.json 2> $out/${tag}_n${n}_$c.err; echo "n$n $c rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_n${n}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],4), d.get('e2e'), d.get('verify'), d['rays']['ms'])
    except Exception as e: print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
