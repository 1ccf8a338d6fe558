#!/bin/bash
# N-GPU session: slab parity and step-with-migration checks (both transports), then the slab bench per transport.  usage: gpu_slab_ab.sh TAG NGPUS [nx] [transports]
TAG=${1:-slab}; N=${2:-2}; NX=${3:-512}; TRS=${4:-"nccl callbacks"}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for t in $TRS; do
  timeout 600 $TR --master-port 29541 tools/slab_check.py 32 $t > $OUT/check_$t.txt 2>&1; grep "SLAB CHECK" $OUT/check_$t.txt
  timeout 600 $TR --master-port 29543 tools/slab_step_check.py 32 $t 3 > $OUT/stepcheck_$t.txt 2>&1; grep "SLAB STEP CHECK" $OUT/stepcheck_$t.txt
done
for t in $TRS; do
  timeout 900 $TR --master-port 29542 bench.py --gpus $N --nx $NX --steps 10 --warmup 3 --no-cpu --transport $t 2>$OUT/bench_$t.err | tail -1 > $OUT/bench_$t.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$t.json"))
    print("$t", round(d["ms_per_step"], 2), "ms/step  e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 2), {k: round(v, 2) for k, v in d["phases_ms"].items()}, d["comm"]["allreduces_per_step"], "allreduces  parity ok:", (d.get("parity_vs_single") or {}).get("ok"), " step:", d.get("step_resident"))
except Exception as ex:
    print("$t unreadable", ex); import subprocess; print(open("$OUT/bench_$t.err").read()[-1500:])
PY
done
