#!/bin/bash
# multi-GPU session: gpurun --gpus N -- 'bash tools/gpu_slab.sh tag N nx'
TAG=${1:-slab}; N=${2:-2}; NX=${3:-256}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee $OUT/gpus.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== slab check N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/slab_check.py 32 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -20 | tee $OUT/slab_check.txt
echo "== bench N=1"; timeout 600 python bench.py --nx $NX --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_n1.json | cut -c1-400
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --nx $NX --steps 5 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3 | tee $OUT/bench_n$N.json | cut -c1-1500
