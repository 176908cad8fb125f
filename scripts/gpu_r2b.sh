#!/bin/bash
# round 2, second GPU call: parity after the queued multi-pass change; four-wide stack variants vs two-wide
out=gpurun_out; mkdir -p $out
V=$PWD/simd-raytracer_b200/variants
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2b_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/r2b_pytest.log)"
# every multi-sample frame of the suite as several queued passes (tiny pass budget), both stack homes of the four-wide traversal
RT_B200_PASS_ENTRIES=40000 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not peer and not gi_128 and not later_one_sample" > $out/r2b_pytest_smallpass.log 2>&1; echo "pytest small passes rc=$? $(tail -1 $out/r2b_pytest_smallpass.log)"
RT_B200_LIB=$V/librt_ss0.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "accelerated or config5 or sparse or sequence" > $out/r2b_pytest_ss0.log 2>&1; echo "pytest ss0 rc=$? $(tail -1 $out/r2b_pytest_ss0.log)"
RT_B200_LIB=$V/librt_ss36nt.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "accelerated or config5 or sparse or sequence" > $out/r2b_pytest_ss36nt.log 2>&1; echo "pytest ss36nt rc=$? $(tail -1 $out/r2b_pytest_ss36nt.log)"
b() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline $ARGS > $out/r2b_$tag.json 2> $out/r2b_$tag.err; echo "$tag rc=$? $(python - <<PY
import json
try:
    d=json.loads(open("$out/r2b_$tag.json").read().strip().splitlines()[-1])
    r=d["rays"]; print("ms/step %.4f  Mrays/s %.0f  prim %.3f sec %.3f shad %.3f shade %.3f res %.3f  e2e %.3f"%(d["ms_per_step"],d["value"],r["ms"]["ms_primary"],r["ms"]["ms_secondary"],r["ms"]["ms_shadow"],r["ms"]["ms_shade"],r["ms"]["ms_resolve"],d["e2e"]["ms_per_frame"]))
except Exception as e: print("parse failed",e)
PY
)"
}
for cfg in cfg2 cfg3; do
  ARGS="--workload $cfg"
  b ${cfg}_w2 RT_B200_ACCEL_WIDTH=2
  b ${cfg}_w4 RT_B200_ACCEL_WIDTH=4
  b ${cfg}_w4_ss0 RT_B200_ACCEL_WIDTH=4 RT_B200_LIB=$V/librt_ss0.so
  b ${cfg}_w4_ss36nt RT_B200_ACCEL_WIDTH=4 RT_B200_LIB=$V/librt_ss36nt.so
done
ARGS="--workload cfg5 --tris 1000000 --steps 5"
b cfg5_1M_w2 RT_B200_ACCEL_WIDTH=2
b cfg5_1M_w4 RT_B200_ACCEL_WIDTH=4
b cfg5_1M_w4_ss0 RT_B200_ACCEL_WIDTH=4 RT_B200_LIB=$V/librt_ss0.so
b cfg5_1M_w4_ss36nt RT_B200_ACCEL_WIDTH=4 RT_B200_LIB=$V/librt_ss36nt.so
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.max
for v in ss0 ss36nt; do
RT_B200_LIB=$V/librt_$v.so RT_B200_ACCEL_WIDTH=4 timeout 600 ncu --metrics $M --clock-control none -k regex:k_stream -s 12 -c 3 --csv --log-file $out/r2b_ncu_$v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/r2b_ncu_$v.log 2>&1; echo "ncu $v rc=$?"
done
python - <<PY
import csv,glob
for f in sorted(glob.glob("$out/r2b_ncu_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>14 and r[0].isdigit()]
    d={}
    for r in rows:
        k=(r[0], r[4].split("(")[0][-28:]); d.setdefault(k,{})[r[12]]=float(r[14].replace(",",""))
    print(f)
    for k,v in d.items():
        wi=v.get("smsp__inst_executed.sum",0); ti=v.get("smsp__thread_inst_executed.sum",0)
        print("  ",k[1], f"us {v.get('gpu__time_duration.sum',0)/1e3:8.1f} warp-inst {wi/1e6:7.2f}M lanes/inst {ti/max(wi,1):5.2f} issue {v.get('smsp__issue_active.avg.pct_of_peak_sustained_active',0):5.1f}% warps {v.get('sm__warps_active.avg.pct_of_peak_sustained_active',0):5.1f}% sm-active {v.get('sm__cycles_active.avg',0)/max(v.get('sm__cycles_elapsed.max',1),1)*100:5.1f}%")
PY
