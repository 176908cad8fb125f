#!/bin/bash
# developer helper (one gpurun call): GPU parity tests, the default bench + reference arm, ncu launch lists and full captures
# of the trace kernels in both query modes.  Outputs land in gpurun_out/ with the given tag.
tag=${1:-r1}
out=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $out/${tag}_pytest_gpu.log
python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err; echo "bench default rc=$?"
python bench.py --mode exact --no-cpu-baseline > $out/${tag}_bench_exact.json 2> $out/${tag}_bench_exact.err
python bench.py --mode ordered --no-cpu-baseline > $out/${tag}_bench_ordered.json 2> $out/${tag}_bench_ordered.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "reference arm rc=$?"
for m in exact ordered; do
  cmd="python bench.py --mode $m --steps 2 --warmup 3 --no-cpu-baseline"
  $cmd > $out/${tag}_plain_$m.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_$m.csv $cmd > $out/${tag}_ncu_launches_$m.log 2>&1
  echo "launch list $m rc=$?"
done
cmd="python bench.py --mode exact --steps 2 --warmup 3 --no-cpu-baseline"
$cmd > $out/${tag}_plain2_exact.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_primary|k_trace_level|k_shadow|k_shade|k_resolve' -s 60 -c 14 -o $out/${tag}_prof_exact $cmd > $out/${tag}_ncu_full_exact.log 2>&1
echo "full exact rc=$?"
cmd="python bench.py --mode ordered --steps 2 --warmup 3 --no-cpu-baseline"
$cmd > $out/${tag}_plain2_ordered.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_stream' -s 21 -c 7 -o $out/${tag}_prof_ordered $cmd > $out/${tag}_ncu_full_ordered.log 2>&1
echo "full ordered rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_bench_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, "ERR", e)
PY
