#!/bin/bash
# developer helper (gpurun --gpus 2): GPU tests, cfg5 single-GPU bench at 1M and 10M triangles, 2-GPU bench with the fused
# peer combine and with the NCCL baseline
tag=${1:-r1b}
out=gpurun_out
nvidia-smi -L; nproc; free -g | head -2
timeout 900 python -m pytest tests -q -m gpu -x > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
for m in ordered exact; do
timeout 600 python bench.py --workload cfg5 --tris 1000000 --mode $m --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_cfg5_1M_$m.json 2> $out/${tag}_bench_cfg5_1M_$m.err; echo "cfg5 1M $m rc=$?"
done
RT_B200_VERBOSE=1 timeout 900 python bench.py --workload cfg5 --mode ordered --steps 5 --warmup 3 > $out/${tag}_bench_cfg5_10M_ordered.json 2> $out/${tag}_bench_cfg5_10M_ordered.err; echo "cfg5 10M ordered rc=$?"
for c in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --combine $c > $out/${tag}_bench_cfg2_n2_$c.json 2> $out/${tag}_bench_cfg2_n2_$c.err; echo "cfg2 n2 $c rc=$?"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg5 --tris 1000000 --steps 5 --warmup 3 > $out/${tag}_bench_cfg5_1M_n2.json 2> $out/${tag}_bench_cfg5_1M_n2.err; echo "cfg5 1M n2 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],3), round((d.get('e2e') or {}).get('value',0),1), d.get('verify'), d.get('scene'), d['rays']['ms'])
    except Exception as e: print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
