#!/bin/bash
# profiling session at the bench size: launch list of one step + ncu --set full of the dominant kernels
# (the pair-kernel capture runs the host call unchunked, so the captured launch covers every row like the resident launch bench.py times)
TAG=${1:-prof}; NX=${2:-512}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file $OUT/launches_nx$NX.csv python bench.py --nx $NX --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
python tools/launch_summary.py $OUT/launches_nx$NX.csv 120 > $OUT/launches_nx${NX}_summary.txt; cat $OUT/launches_nx${NX}_summary.txt
NDSPMHD_B200_RATE_CHUNKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rates_pair -s 3 -c 1 -o $OUT/prof_rates_nx$NX python bench.py --nx $NX --steps 1 --warmup 3 --no-cpu > $OUT/ncu_rates.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:density_round -s 9 -c 2 -o $OUT/prof_density_nx$NX python bench.py --nx $NX --steps 1 --warmup 3 --no-cpu > $OUT/ncu_density.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:build_lists -s 12 -c 3 -o $OUT/prof_lists_nx$NX python bench.py --nx $NX --steps 1 --warmup 3 --no-cpu > $OUT/ncu_lists.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rates_final -s 3 -c 1 -o $OUT/prof_final_nx$NX python bench.py --nx $NX --steps 1 --warmup 3 --no-cpu > $OUT/ncu_final.log 2>&1
ls -la $OUT
