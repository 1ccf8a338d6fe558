#!/bin/bash
# Full-size bench exactly as the driver runs it: reference arm, our arm (N=1)
TAG=${1:-full}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-600
echo "== ours N=1"; timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_n1.json | cut -c1-3000
