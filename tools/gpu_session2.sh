#!/bin/bash
# Round record: parity tests, smoke, reference arm, bench at full size, launch list + ncu captures.  gpurun --timeout 1800 -- 'bash tools/gpu_session2.sh tag'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt; free -g >> $OUT/host.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-300
echo "== bench nx=512"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_512.json | cut -c1-300
echo "== bench nx=256"; timeout 600 python bench.py --nx 256 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_256.json | cut -c1-300
echo "== ncu launches (nx=512)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file $OUT/launches_512.csv python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
echo "== ncu full rates (nx=512)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rates_pair -s 3 -c 1 -o $OUT/prof_rates_512 python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/ncu_rates.log 2>&1
echo "== ncu full density + lists (nx=256)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:density_round -s 9 -c 3 -o $OUT/prof_density_256 python bench.py --nx 256 --steps 1 --warmup 3 --no-cpu > $OUT/ncu_density.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:build_lists -s 12 -c 4 -o $OUT/prof_lists_256 python bench.py --nx 256 --steps 1 --warmup 3 --no-cpu > $OUT/ncu_lists.log 2>&1
ls -la $OUT
