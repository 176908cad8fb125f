#!/bin/bash
# developer helper (gpurun --gpus 2): all GPU tests, N=1 bench, N=2 bench (pipelined peer combine, NCCL baseline)
tag=${1:-r1e}
out=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $out/${tag}_pytest_gpu.log
python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "n1 rc=$?"
for c in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --combine $c > $out/${tag}_bench_n2_$c.json 2> $out/${tag}_bench_n2_$c.err; echo "n2 $c rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],4), d.get('e2e'), d.get('verify'), d['rays']['ms'])
    except Exception as e: print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
