#!/bin/bash
# device-built hierarchy (rt_build_opts.accel_build = device): parity tests, then build time and frame time beside the host build
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "device_built" > $out/r2n_pytest_lbvh.log 2>&1; echo "pytest lbvh rc=$? $(tail -1 $out/r2n_pytest_lbvh.log)"
for build in host device; do
  for wl in "cfg2" "cfg3" "cfg5 --tris 1000000" "cfg5 --tris 10000000"; do
    tag=$(echo $wl | tr -d ' -' )
    RT_B200_VERBOSE=1 timeout 900 python bench.py --workload $wl --accel-build $build --steps 20 --warmup 5 --no-cpu-baseline --ns-tris 0 \
        > $out/r2n_${tag}_${build}.json 2> $out/r2n_${tag}_${build}.err
    echo "$wl $build rc=$? $(python - <<PY
import json
try:
    d = json.loads(open("$out/r2n_${tag}_${build}.json").read().strip().splitlines()[-1])
    print("ms/step %.4f value %.1f" % (d["ms_per_step"], d["value"]), d["scene"])
except Exception as e:
    print("no line", e)
PY
)"
    grep "device bvh\|bvh build" $out/r2n_${tag}_${build}.err | head -3
  done
done
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2n_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/r2n_pytest.log)"
