#!/bin/bash
# one ncu --set full capture of one kernel on the bench workload: gpu_ncu1.sh tag regex skip [nx] [lib]
TAG=${1:-ncu1}; RX=${2:-rates_pair}; SKIP=${3:-3}; NX=${4:-256}
OUT=gpurun_out/$TAG; mkdir -p $OUT
[ -n "$5" ] && export NDSPMHD_B200_LIB=$PWD/$5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c 1 -o $OUT/prof_$RX python bench.py --nx $NX --steps 1 --warmup 3 --no-cpu > $OUT/ncu_$RX.log 2>&1
ls -la $OUT
