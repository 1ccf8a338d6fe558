#!/bin/bash
# multi-GPU bench only: gpurun --gpus N -- 'bash tools/gpu_slab2.sh tag N nx [nx2]'
TAG=${1:-slab}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== slab check N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/slab_check.py 32 2>&1 | grep -E "PASSED|FAIL|rror" | tail -3
for NX in $3 $4; do
echo "== bench N=$N nx=$NX"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --nx $NX --steps 5 --warmup 3 2>&1 | grep "^{" | tail -1 | tee $OUT/bench_n${N}_nx$NX.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['ms_per_step'], d['value']/1e6, d['phases_ms'], d['comm'], 'e2e', d['e2e']['ms_per_step'])"
done
