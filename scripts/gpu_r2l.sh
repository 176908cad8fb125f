#!/bin/bash
out=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2l_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/r2l_pytest.log)"
bash scripts/gpu_sweep.sh r2l 2>&1 | tee $out/r2l_sweep.txt
RT_B200_VERBOSE=1 timeout 600 python bench.py --workload cfg5 --tris 1000000 --steps 4 --warmup 4 --no-cpu-baseline --ns-tris 0 2>&1 | grep "rt_b200\]" | sort | uniq -c | head
