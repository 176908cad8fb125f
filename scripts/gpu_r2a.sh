#!/bin/bash
# round 2, first GPU call: parity of both hierarchy widths, then two-wide vs four-wide A/B on configs 1-3 and 5 (1 M triangles)
out=gpurun_out; mkdir -p $out
V=$PWD/simd-raytracer_b200/variants
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2a_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/r2a_pytest.log)"
# the thread-local tail of the four-wide stack: a build that keeps only two entries per lane in shared memory
RT_B200_LIB=$V/librt_ss2.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "accelerated or config5 or gi_128 or sparse or sequence" > $out/r2a_pytest_ss2.log 2>&1; echo "pytest ss2 rc=$? $(tail -1 $out/r2a_pytest_ss2.log)"
b() { # tag, env..., -- args
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline $ARGS > $out/r2a_$tag.json 2> $out/r2a_$tag.err; echo "$tag rc=$? $(python - <<PY
import json
try:
    d=json.loads(open("$out/r2a_$tag.json").read().strip().splitlines()[-1])
    r=d["rays"]; print("ms/step %.4f  Mrays/s %.0f  prim %.3f sec %.3f shad %.3f shade %.3f res %.3f  e2e %.3f"%(d["ms_per_step"],d["value"],r["ms"]["ms_primary"],r["ms"]["ms_secondary"],r["ms"]["ms_shadow"],r["ms"]["ms_shade"],r["ms"]["ms_resolve"],d["e2e"]["ms_per_frame"]))
except Exception as e: print("parse failed",e)
PY
)"
}
for cfg in cfg2 cfg1 cfg3; do
  ARGS="--workload $cfg"
  b ${cfg}_w2 RT_B200_ACCEL_WIDTH=2
  b ${cfg}_w4 RT_B200_ACCEL_WIDTH=4
done
ARGS="--workload cfg2"
for v in mb2 t128x5 ss8; do
  b cfg2_w4_$v RT_B200_ACCEL_WIDTH=4 RT_B200_LIB=$V/librt_$v.so
  b cfg2_w2_$v RT_B200_ACCEL_WIDTH=2 RT_B200_LIB=$V/librt_$v.so
done
ARGS="--workload cfg5 --tris 1000000 --steps 5"
b cfg5_1M_w2 RT_B200_ACCEL_WIDTH=2
b cfg5_1M_w4 RT_B200_ACCEL_WIDTH=4
b cfg5_1M_w4_mb2 RT_B200_ACCEL_WIDTH=4 RT_B200_LIB=$V/librt_mb2.so
# instruction counts / lane utilisation of the stream kernels, two-wide vs four-wide (one ncu pass each; the plain runs above exited 0)
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.max
for w in 2 4; do
RT_B200_ACCEL_WIDTH=$w timeout 600 ncu --metrics $M --clock-control none -k regex:k_stream -s 12 -c 6 --csv --log-file $out/r2a_ncu_w$w.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/r2a_ncu_w$w.log 2>&1; echo "ncu w$w rc=$?"
done
python - <<PY
import csv,glob
for f in sorted(glob.glob("$out/r2a_ncu_w*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>14 and r[0].isdigit()]
    d={}
    for r in rows:
        k=(r[0], r[4].split("(")[0][-28:]); d.setdefault(k,{})[r[12]]=float(r[14].replace(",",""))
    print(f)
    for k,v in d.items():
        wi=v.get("smsp__inst_executed.sum",0); ti=v.get("smsp__thread_inst_executed.sum",0)
        print("  ",k[1], f"us {v.get('gpu__time_duration.sum',0)/1e3:8.1f} warp-inst {wi/1e6:7.2f}M lanes/inst {ti/max(wi,1):5.2f} issue {v.get('smsp__issue_active.avg.pct_of_peak_sustained_active',0):5.1f}% warps {v.get('sm__warps_active.avg.pct_of_peak_sustained_active',0):5.1f}% sm-active {v.get('sm__cycles_active.avg',0)/max(v.get('sm__cycles_elapsed.max',1),1)*100:5.1f}%")
PY
