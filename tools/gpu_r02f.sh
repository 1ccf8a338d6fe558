#!/bin/bash
OUT=gpurun_out/r02f; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_step.py -q -x 2>&1 | tail -5 | tee $OUT/step_tests.txt
for t in nccl callbacks; do
  timeout 600 $TR --master-port 29551 tools/slab_step_check.py 32 $t 3 > $OUT/stepcheck_$t.txt 2>&1; grep -v "^W\|^\*\|^$" $OUT/stepcheck_$t.txt | tail -12
done
timeout 600 $TR --master-port 29552 tools/slab_step_check.py 24 nccl 8 > $OUT/stepcheck_nccl_8steps.txt 2>&1; grep -v "^W\|^\*\|^$" $OUT/stepcheck_nccl_8steps.txt | tail -8
