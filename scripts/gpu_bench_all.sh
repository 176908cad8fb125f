#!/bin/bash
# developer helper: gpu tests (optional) + a short bench of every workload / mode into gpurun_out/
tag=${1:-x}
if [ "$2" = "tests" ]; then timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_gpu_$tag.log; fi
for w in cfg2 cfg1 cfg3 cfg4; do for m in exact ordered; do
python bench.py --workload $w --mode $m --steps 30 --no-cpu-baseline > gpurun_out/bench_${tag}_${w}_$m.json 2>gpurun_out/bench_${tag}_${w}_$m.err || tail -3 gpurun_out/bench_${tag}_${w}_$m.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_${tag}_*.json")):
    try:
        d=json.load(open(f)); ms=d["rays"]["ms"]
        print(f"{f.split('bench_')[1][:-5]:24s} {d['value']:8.0f} Mrays/s {d['ms_per_step']:7.3f} ms/frame e2e {d['e2e']['ms_per_frame']:7.3f} | prim {ms['ms_primary']:.3f} sec {ms['ms_secondary']:.3f} shad {ms['ms_shadow']:.3f} shade {ms['ms_shade']:.3f} res {ms['ms_resolve']:.3f}")
    except Exception as e: print(f, "ERR", e)
PY
