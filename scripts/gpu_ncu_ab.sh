#!/bin/bash
# developer helper: instruction counts / lane utilisation of the stream kernels, default build vs variant libraries
out=gpurun_out; tag=${1:-ab}
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics $M --clock-control none -k regex:k_stream -s 9 -c 6 --csv --log-file $out/${tag}_base.csv $cmd > $out/${tag}_base.log 2>&1; echo "base rc=$?"
shopt -s nullglob
for v in simd-raytracer_b200/variants/librt_*.so; do n=$(basename $v .so); n=${n#librt_}
RT_B200_LIB=$PWD/$v ncu --metrics $M --clock-control none -k regex:k_stream -s 9 -c 6 --csv --log-file $out/${tag}_$n.csv $cmd > $out/${tag}_$n.log 2>&1; echo "$n rc=$?"
done
python - <<PY
import csv,glob
for f in sorted(glob.glob("$out/${tag}_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>14 and r[0].isdigit()]
    d={}
    for r in rows:
        k=(r[0], r[4].split("(")[0][-28:]); d.setdefault(k,{})[r[12]]=float(r[14].replace(",",""))
    print(f)
    for k,v in d.items():
        wi=v.get("smsp__inst_executed.sum",0); ti=v.get("smsp__thread_inst_executed.sum",0)
        print("  ",k[1], f"us {v.get('gpu__time_duration.sum',0)/1e3:8.1f} warp-inst {wi/1e6:7.2f}M lanes/inst {ti/max(wi,1):5.2f} issue {v.get('smsp__issue_active.avg.pct_of_peak_sustained_active',0):5.1f}% warps {v.get('sm__warps_active.avg.pct_of_peak_sustained_active',0):5.1f}%")
PY
