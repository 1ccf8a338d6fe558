#!/bin/bash
# 8-GPU session: slab parity + step-with-migration at 8 ranks (nx=64: slabs 8 spacings wide), then the slab bench (native transport)
TAG=${1:-n8}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/slab_check.py 64 nccl > $OUT/check_nccl.txt 2>&1; grep "SLAB CHECK" $OUT/check_nccl.txt
timeout 300 $TR --master-port 29543 tools/slab_step_check.py 64 nccl 3 > $OUT/stepcheck_nccl.txt 2>&1; grep "SLAB STEP CHECK" $OUT/stepcheck_nccl.txt
for cfgname in slab512 cube256; do
timeout 600 $TR --master-port 29542 bench.py --gpus $N --config $cfgname --steps 10 --warmup 3 --no-cpu 2>$OUT/bench_$cfgname.err | tail -1 > $OUT/bench_$cfgname.json
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$cfgname.json"))
    print("$cfgname", round(d["ms_per_step"], 2), "ms/step  e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 2), {k: round(v, 2) for k, v in d["phases_ms"].items()}, d["comm"]["allreduces_per_step"], "allreduces  parity ok:", (d.get("parity_vs_single") or {}).get("ok"), " step:", d.get("step_resident"))
except Exception as ex:
    print("$cfgname unreadable", ex); print(open("$OUT/bench_$cfgname.err").read()[-1500:])
PY
done
