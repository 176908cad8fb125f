#!/bin/bash
# round 2: the evidence run - parity suite, smoke, ncu kernel counts (feed bench.py's roofline), bench lines of every config, the
# reference arm, ncu --set full of the dominant kernels, the launch list of a config-2 frame
out=gpurun_out; tag=${1:-r2}; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/${tag}_pytest.log)"
timeout 600 python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $out/${tag}_smoke.log)"
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
counts() { # key variant passes bench-args...
  key=$1; var=$2; passes=$3; shift 3
  timeout 1500 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ns-tris 0 "$@" > $out/${tag}_plain_$key.log 2>&1 &&
  timeout 1500 ncu --metrics $M --clock-control none --csv --log-file $out/${tag}_counts_$key.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ns-tris 0 "$@" > $out/${tag}_counts_$key.log 2>&1
  echo "ncu counts $key rc=$?"
  python scripts/make_kernel_counts.py $out/${tag}_counts_$key.csv $key $var $passes > $out/${tag}_counts_$key.txt 2>&1; head -4 $out/${tag}_counts_$key.txt
}
counts cfg2_hw09_scene5 ordered/bvh4/spp1 1
counts cfg1_hw15_scene2 ordered/bvh4/spp1 1 --workload cfg1
counts cfg3_hw11_scene8_d10 ordered/bvh4/spp1 1 --workload cfg3
counts cfg4_hw12_scene4 ordered/bvh4/spp128 8 --workload cfg4
counts cfg5_synthetic_10000000 ordered/bvh4/spp1 1 --workload cfg5
cp profiles/kernel_counts.json $out/${tag}_kernel_counts.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "reference arm rc=$?"
timeout 900 python bench.py --steps 200 --warmup 10 > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err; echo "bench default rc=$?"
timeout 900 python bench.py --steps 100 --warmup 5 --workload cfg1 --ns-tris 0 > $out/${tag}_bench_cfg1.json 2> $out/${tag}_bench_cfg1.err; echo "bench cfg1 rc=$?"
timeout 900 python bench.py --steps 40 --warmup 5 --workload cfg3 --ns-tris 0 > $out/${tag}_bench_cfg3.json 2> $out/${tag}_bench_cfg3.err; echo "bench cfg3 rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 --workload cfg4 --ns-tris 0 > $out/${tag}_bench_cfg4_128spp.json 2> $out/${tag}_bench_cfg4_128spp.err; echo "bench cfg4 rc=$?"
timeout 1500 python bench.py --steps 3 --warmup 3 --workload cfg5 --ns-tris 0 > $out/${tag}_bench_cfg5_10M.json 2> $out/${tag}_bench_cfg5_10M.err; echo "bench cfg5 10M rc=$?"
timeout 1500 python bench.py --steps 2 --warmup 2 --workload cfg5 --spp 16 --ns-tris 0 --no-cpu-baseline > $out/${tag}_bench_cfg5_10M_spp16.json 2> $out/${tag}_bench_cfg5_10M_spp16.err; echo "bench cfg5 10M spp16 rc=$?"
timeout 900 python bench.py --steps 100 --warmup 5 --accel-width 2 --ns-tris 0 --no-cpu-baseline > $out/${tag}_bench_default_bvh2.json 2> $out/${tag}_bench_default_bvh2.err; echo "bench bvh2 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get("rays",{}); rf=d.get("roofline") or {}
        print(f.split("/")[-1], "ms/step %.4f value %.0f e2e %.4f"%(d["ms_per_step"],d["value"],d["e2e"].get("ms_per_frame",0)), "roofline", rf.get("bound"), rf.get("frac"), "lanes", rf.get("lanes_per_inst"), "| cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "parse failed", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stream" -s 24 -c 3 -o $out/${tag}_prof_cfg2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ns-tris 0 > $out/${tag}_prof_cfg2.log 2>&1; echo "ncu full cfg2 rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_stream_shadow" -s 3 -c 1 -o $out/${tag}_prof_cfg5_10M python bench.py --steps 1 --warmup 3 --workload cfg5 --no-cpu-baseline --ns-tris 0 > $out/${tag}_prof_cfg5_10M.log 2>&1; echo "ncu full cfg5 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 120 --csv --log-file $out/${tag}_cfg2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ns-tris 0 > $out/${tag}_cfg2_launches.log 2>&1; echo "ncu launches cfg2 rc=$?"
