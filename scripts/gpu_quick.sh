#!/bin/bash
# developer helper: GPU parity tests, default bench (and the same with RT_B200_NO_PDL=1), launch-size probe
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/g1_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/g1_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/g1_bench.json 2> gpurun_out/g1_bench.err; echo rc=$?
RT_B200_NO_PDL=1 python bench.py --no-cpu-baseline > gpurun_out/g1_bench_nopdl.json 2> gpurun_out/g1_bench_nopdl.err; echo rc=$?
for w in cfg1 cfg3; do python bench.py --no-cpu-baseline --workload $w > gpurun_out/g1_bench_$w.json 2> gpurun_out/g1_bench_$w.err; RT_B200_NO_PDL=1 python bench.py --no-cpu-baseline --workload $w > gpurun_out/g1_bench_${w}_nopdl.json 2>/dev/null; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/g1_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],4), round(d["value"],1), round(d["e2e"]["value"],1), round(d["e2e"]["one_call_per_frame"]["ms_per_frame"],3), {k: round(v,4) for k,v in d["rays"]["ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/g1_bench.err
