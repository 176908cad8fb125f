#!/bin/bash
# developer helper (one gpurun call, 1 GPU): GPU parity tests, default bench + reference arm, ncu launch list of the bench
# command and ncu --set full captures of one default-mode frame (cfg2).  Outputs land in gpurun_out/ with the given tag.
tag=${1:-r1g}
out=gpurun_out
nvidia-smi -L; nproc
timeout 900 python -m pytest tests -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $out/${tag}_pytest_gpu.log
python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err; echo "bench default rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "reference arm rc=$?"
for wl in cfg1 cfg3 cfg4; do
python bench.py --workload $wl --no-cpu-baseline > $out/${tag}_bench_$wl.json 2> $out/${tag}_bench_$wl.err
done
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$cmd > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$cmd > $out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:^k_|rtb::k_' -s 45 -c 12 -o $out/${tag}_prof_cfg2 $cmd > $out/${tag}_ncu_full.log 2>&1
echo "full cfg2 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'), (d.get('rays') or {}).get('ms'))
    except Exception as e: print(f, "ERR", e)
PY
