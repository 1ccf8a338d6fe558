#!/bin/bash
# N-GPU session: slab parity (both transports, LIGHT off/on) and the slab bench per transport.  usage: gpu_slab_ab.sh TAG NGPUS [nx]
TAG=${1:-slab}; N=${2:-2}; NX=${3:-512}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for t in nccl callbacks; do
  for light in 0 1; do
    NDSPMHD_B200_SLAB_LIGHT=$light timeout 600 $TR --master-port 29541 tools/slab_check.py 32 $t > $OUT/check_${t}_light$light.txt 2>&1
    tail -4 $OUT/check_${t}_light$light.txt
  done
done
for t in nccl callbacks; do
  for light in 0 1; do
    NDSPMHD_B200_SLAB_LIGHT=$light timeout 900 $TR --master-port 29542 bench.py --gpus $N --nx $NX --steps 10 --warmup 3 --no-cpu --transport $t 2>$OUT/bench_${t}_light$light.err | tail -1 > $OUT/bench_${t}_light$light.json
    python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${t}_light$light.json"))
    print("$t light=$light", d["ms_per_step"], "ms/step e2e", d.get("e2e", {}).get("ms_per_step"), d["phases_ms"], d["comm"], "parity ok:", (d.get("parity_vs_single") or {}).get("ok"))
except Exception as ex:
    print("$t light=$light unreadable", ex)
PY
  done
done
