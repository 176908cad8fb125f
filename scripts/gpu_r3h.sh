#!/bin/bash
# config 5 with one triangle per leaf (the scene-size rule of finish_create): parity tests, ncu counts, bench lines, --set full capture
out=gpurun_out; tag=r3h
timeout 1500 python -m pytest tests -m gpu -x -q -k "config5 or device_built or accelerated_mode_same_hits" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/${tag}_pytest.log)"
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
key=cfg5_synthetic_10000000
timeout 1500 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ns-tris 0 --workload cfg5 > $out/${tag}_plain_$key.log 2>&1 &&
timeout 1500 ncu --metrics $M --clock-control none --csv --log-file $out/${tag}_counts_$key.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ns-tris 0 --workload cfg5 > $out/${tag}_counts_$key.log 2>&1
echo "ncu counts rc=$?"
python scripts/make_kernel_counts.py $out/${tag}_counts_$key.csv $key ordered/bvh4/spp1 1 > $out/${tag}_counts_$key.txt 2>&1; head -5 $out/${tag}_counts_$key.txt
cp profiles/kernel_counts.json $out/${tag}_kernel_counts.json
timeout 1500 python bench.py --steps 3 --warmup 3 --workload cfg5 --ns-tris 0 > $out/${tag}_bench_cfg5_10M.json 2> $out/${tag}_bench_cfg5_10M.err; echo "bench cfg5 10M rc=$?"
timeout 1500 python bench.py --steps 2 --warmup 2 --workload cfg5 --spp 16 --ns-tris 0 --no-cpu-baseline > $out/${tag}_bench_cfg5_10M_spp16.json 2> $out/${tag}_bench_cfg5_10M_spp16.err; echo "bench cfg5 10M spp16 rc=$?"
timeout 1500 python bench.py --steps 3 --warmup 3 --workload cfg5 --accel-build device --ns-tris 0 --no-cpu-baseline > $out/${tag}_bench_cfg5_10M_device_built.json 2> $out/${tag}_bench_cfg5_10M_device_built.err; echo "bench cfg5 10M device-built rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --workload cfg5 --tris 1000000 --ns-tris 0 --no-cpu-baseline > $out/${tag}_bench_cfg5_1M.json 2> $out/${tag}_bench_cfg5_1M.err; echo "bench cfg5 1M rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); rf=d.get("roofline") or {}; sc=d["scene"]
        print(f.split("/")[-1], "ms/step %.4f value %.0f e2e %.4f"%(d["ms_per_step"],d["value"],d["e2e"].get("ms_per_frame",0)), "roofline", rf.get("bound"), rf.get("frac"), "leaf", sc.get("bvh_leaf_size"), "build", sc.get("accel_build"), sc.get("accel_build_s"), "create", sc.get("host_build_s"))
    except Exception as e: print(f, "parse failed", e)
PY
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_stream_shadow" -s 3 -c 1 -o $out/${tag}_prof_cfg5_10M python bench.py --steps 1 --warmup 3 --workload cfg5 --no-cpu-baseline --ns-tris 0 > $out/${tag}_prof_cfg5_10M.log 2>&1; echo "ncu full cfg5 rc=$?"
