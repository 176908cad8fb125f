#!/bin/bash
# round 2, third GPU call: full parity suite (adapter binary, queued rgb8, crtscene/JPEG), the new bench line, ncu kernel counts
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q -s > $out/r2c_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/r2c_pytest.log)"; grep "adapter timings" $out/r2c_pytest.log
timeout 900 python bench.py --steps 100 --warmup 5 > $out/r2c_bench_cfg2.json 2> $out/r2c_bench_cfg2.err; echo "bench cfg2 rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 --workload cfg4 --ns-tris 0 > $out/r2c_bench_cfg4.json 2> $out/r2c_bench_cfg4.err; echo "bench cfg4 rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --workload cfg5 --tris 1000000 --spp 8 --ns-tris 0 --no-cpu-baseline > $out/r2c_bench_cfg5_1M_spp8.json 2> $out/r2c_bench_cfg5_1M_spp8.err; echo "bench cfg5 rc=$?"
python - <<PY
import json
for t in ("cfg2","cfg4","cfg5_1M_spp8"):
    try:
        d=json.loads(open("$out/r2c_bench_%s.json"%t).read().strip().splitlines()[-1])
        e=d["e2e"]; r=d["rays"]
        print(t,"ms/step %.4f Mrays/s %.0f passes %s | e2e %.4f ms (x%.3f of device) float-seq %.4f one-call %.4f | ms %s"%(d["ms_per_step"],d["value"],r.get("passes_per_frame"),e["ms_per_frame"],e["vs_device_time"],e["float_sequence"]["ms_per_frame"],e["one_call_per_frame"]["ms_per_frame"],{k:round(v,4) for k,v in r["ms"].items()}))
        print("   clocks",d["clocks"]," ns:",json.dumps(d.get("north_star_scaling"))[:400])
    except Exception as ex: print(t,"parse failed",ex)
PY
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $out/r2c_counts_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ns-tris 0 > $out/r2c_counts_cfg2.log 2>&1; echo "ncu counts cfg2 rc=$?"
python scripts/make_kernel_counts.py $out/r2c_counts_cfg2.csv cfg2_hw09_scene5 ordered/bvh4/spp1 | tee $out/r2c_counts_cfg2.txt
cp profiles/kernel_counts.json $out/r2c_kernel_counts.json
