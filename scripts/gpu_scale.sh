#!/bin/bash
# developer helper (gpurun --gpus N): the default bench at N ranks, as the driver launches it, then config 5 (1M triangles)
n=${1:-8}; tag=${2:-scale}
out=gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 200 --warmup 10 > $out/${tag}_n${n}_cfg2.json 2> $out/${tag}_n${n}_cfg2.err; echo "n$n cfg2 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $n --workload cfg5 --tris 1000000 --steps 10 --warmup 3 > $out/${tag}_n${n}_cfg5_1M.json 2> $out/${tag}_n${n}_cfg5_1M.err; echo "n$n cfg5 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_n${n}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],4), d.get('e2e'), d.get('verify'), d['rays']['ms'])
    except Exception as e: print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
