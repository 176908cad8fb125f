#!/bin/bash
# developer helper: ncu --set full with source counters for the accelerated trace kernels (cfg2 by default)
tag=${1:-prof}; w=${2:-cfg2}; extra=${3:-}
out=gpurun_out
cmd="python bench.py --workload $w --mode ordered --steps 2 --warmup 3 --no-cpu-baseline $extra"
$cmd > $out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_stream' -s 21 -c 3 -o $out/${tag}_$w $cmd > $out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $out/${tag}_ncu.log
