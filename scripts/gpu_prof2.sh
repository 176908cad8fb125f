#!/bin/bash
# developer helper: cfg5 (10M) bench + ncu full captures (source counters) of the trace kernels on cfg2 and cfg5-1M
tag=${1:-p3}
out=gpurun_out
RT_B200_VERBOSE=1 timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_cfg5_10M.json 2> $out/${tag}_bench_cfg5_10M.err; echo "cfg5 10M rc=$?"
cmd="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline"
$cmd > $out/${tag}_plain_cfg2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_stream' -s 21 -c 3 -o $out/${tag}_cfg2 $cmd > $out/${tag}_ncu_cfg2.log 2>&1
echo "ncu cfg2 rc=$?"
cmd="python bench.py --workload cfg5 --tris 1000000 --steps 1 --warmup 3 --no-cpu-baseline"
$cmd > $out/${tag}_plain_cfg5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_stream' -s 42 -c 7 -o $out/${tag}_cfg5 $cmd > $out/${tag}_ncu_cfg5.log 2>&1
echo "ncu cfg5 rc=$?"
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_cfg5_10M.json").read().strip().splitlines()[-1]); print(round(d['value'],1), d['ms_per_step'], d['rays']['ms'], d['scene'], d['e2e'])
PY
