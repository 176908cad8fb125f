#!/bin/bash
# round 2: config 5 at 10 M triangles (bench, ncu counts, ncu --set full of the trace kernels) and the --set full capture of config 2
out=gpurun_out; mkdir -p $out
timeout 1200 python bench.py --steps 3 --warmup 3 --workload cfg5 --ns-tris 0 > $out/r2i_bench_cfg5_10M.json 2> $out/r2i_bench_cfg5_10M.err; echo "bench cfg5 10M rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$out/r2i_bench_cfg5_10M.json").read().strip().splitlines()[-1]); r=d["rays"]["ms"]
    print("cfg5 10M ms/step %.3f Mrays/s %.0f  prim %.2f sec %.2f shad %.2f shade %.2f res %.2f | scene %s | cpu %s"%(d["ms_per_step"],d["value"],r["ms_primary"],r["ms_secondary"],r["ms_shadow"],r["ms_shade"],r["ms_resolve"],d["scene"],d.get("cpu_baseline")))
except Exception as e: print("parse failed",e)
PY
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 1200 ncu --metrics $M --clock-control none --csv --log-file $out/r2i_counts_cfg5_10M.csv python bench.py --steps 1 --warmup 3 --workload cfg5 --no-cpu-baseline --ns-tris 0 > $out/r2i_counts_cfg5_10M.log 2>&1; echo "ncu counts cfg5 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_stream_shadow|k_stream_level" -s 12 -c 3 -o $out/r2i_prof_cfg5_10M python bench.py --steps 1 --warmup 3 --workload cfg5 --no-cpu-baseline --ns-tris 0 > $out/r2i_prof_cfg5_10M.log 2>&1; echo "ncu full cfg5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_stream" -s 24 -c 3 -o $out/r2i_prof_cfg2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ns-tris 0 > $out/r2i_prof_cfg2.log 2>&1; echo "ncu full cfg2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 120 --csv --log-file $out/r2i_cfg2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ns-tris 0 > $out/r2i_cfg2_launches.log 2>&1; echo "ncu launches cfg2 rc=$?"
ls -la $out/r2i_*
