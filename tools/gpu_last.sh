#!/bin/bash
# last seconds of the round's GPU budget: ncu --set full of the first LIGHT density round, then the launch list of one bench step (nx=256)
OUT=gpurun_out/v5ncu; mkdir -p $OUT
timeout 26 ncu --set full --clock-control none --import-source on -k regex:density_round -c 1 -o $OUT/prof_density_light_256 python bench.py --nx 256 --steps 1 --warmup 1 --no-cpu > $OUT/ncu_density.log 2>&1
timeout 22 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file $OUT/launches_256.csv python bench.py --nx 256 --steps 1 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
ls -la $OUT
