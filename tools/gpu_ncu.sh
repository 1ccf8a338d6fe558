#!/bin/bash
# ncu --set full captures of the two hot kernels on the bench workload at reduced size
TAG=${1:-ncu}; NX=${2:-256}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rates_pair -s 3 -c 1 -o $OUT/prof_rates python bench.py --nx $NX --steps 1 --warmup 3 --no-cpu > $OUT/ncu_rates.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:density_round -s 9 -c 3 -o $OUT/prof_density python bench.py --nx $NX --steps 1 --warmup 3 --no-cpu > $OUT/ncu_density.log 2>&1
ls -la $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:build_lists -s 12 -c 4 -o $OUT/prof_lists python bench.py --nx $NX --steps 1 --warmup 3 --no-cpu > $OUT/ncu_lists.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file $OUT/launches.csv python bench.py --nx $NX --steps 1 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
ls -la $OUT
