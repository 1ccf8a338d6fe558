#!/bin/bash
# ncu --set full captures of several kernels in one bench process: gpu_ncu2.sh tag nx "regex:skip:count" ...
TAG=${1:-ncu2}; NX=${2:-256}; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
for spec in "$@"; do
  IFS=: read RX SKIP CNT <<< "$spec"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $CNT -o $OUT/prof_$RX python bench.py --nx $NX --steps 1 --warmup 3 --no-cpu > $OUT/ncu_$RX.log 2>&1
done
ls -la $OUT
