timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/g1_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/g1_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/g1_bench.json 2> gpurun_out/g1_bench.err; echo rc=$?
python scripts/gpu_tail_probe.py hw09_scene5 hw11_scene8 hw15_scene2
python - <<PY
import json
d=json.loads(open("gpurun_out/g1_bench.json").read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["rays"]["ms"])
PY
