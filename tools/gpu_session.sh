#!/bin/bash
# One GPU session: parity tests, bench, ncu launch list + full captures.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_session.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt; free -g >> $OUT/host.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench nx=256"; timeout 600 python bench.py --nx 256 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -2 | tee $OUT/bench_256.json
echo "== bench nx=512"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 | tee $OUT/bench_512.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_256.csv python bench.py --nx 256 --steps 1 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
echo "== ncu full rates"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rates_pair -s 2 -c 1 -o $OUT/prof_rates python bench.py --nx 256 --steps 1 --warmup 3 --no-cpu > $OUT/ncu_rates.log 2>&1
echo "== ncu full density"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:density_round -s 6 -c 3 -o $OUT/prof_density python bench.py --nx 256 --steps 1 --warmup 3 --no-cpu > $OUT/ncu_density.log 2>&1
ls -la $OUT
