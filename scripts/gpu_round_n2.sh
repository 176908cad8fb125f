#!/bin/bash
# developer helper (gpurun --gpus 2): N=1 bench, N=2 bench (pipelined peer combine, NCCL baseline), cfg5 at 1 M triangles on one GPU
tag=${1:-r1g}
out=gpurun_out
python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "n1 rc=$?"
for c in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --combine $c > $out/${tag}_bench_n2_$c.json 2> $out/${tag}_bench_n2_$c.err; echo "n2 $c rc=$?"
done
timeout 600 python bench.py --workload cfg5 --tris 1000000 --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_cfg5_1M.json 2> $out/${tag}_bench_cfg5_1M.err; echo "cfg5 1M rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_bench_n*.json")+glob.glob("$out/${tag}_bench_cfg5_1M.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],4), d.get('e2e',{}).get('value'), d.get('e2e',{}).get('ms_per_frame'), d.get('verify'), d['rays']['ms'])
    except Exception as e: print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
