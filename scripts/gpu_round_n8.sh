#!/bin/bash
# developer helper (gpurun --gpus 8): N=8 bench of the default workload with the pipelined peer combine
tag=${1:-r1g}
out=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > $out/${tag}_bench_n8_cfg2.json 2> $out/${tag}_bench_n8_cfg2.err; echo "n8 rc=$?"
python - <<PY
import json
f="$out/${tag}_bench_n8_cfg2.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],4), d.get('e2e',{}).get('value'), d.get('e2e',{}).get('ms_per_frame'), d.get('verify'))
except Exception as e: print("ERR", e); print(open(f.replace('.json','.err')).read()[-3000:])
PY
