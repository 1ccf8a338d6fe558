#!/bin/bash
# 4-GPU session: the 4-rank slab tests of the suite (both transports), then the slab bench (strong, and one weak-scaling line)
TAG=${1:-n4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_slab.py -q -m gpu -k "4-" 2>&1 | tail -6 | tee $OUT/pytest_slab_n4.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
for cfgname in slab512 cube256; do
timeout 600 $TR --master-port 29542 bench.py --gpus 4 --config $cfgname --steps 10 --warmup 3 --no-cpu 2>$OUT/bench_$cfgname.err | tail -1 > $OUT/bench_$cfgname.json
done
timeout 600 $TR --master-port 29544 bench.py --gpus 4 --config slab512 --nx 256 --scaling weak --steps 10 --warmup 3 --no-cpu 2>$OUT/bench_weak.err | tail -1 > $OUT/bench_weak_nx256.json
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], d["scaling"], d["run"]["npart"], round(d["ms_per_step"], 2), "ms/step  e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 2), {k: round(v, 2) for k, v in d["phases_ms"].items()}, "parity ok:", (d.get("parity_vs_single") or {}).get("ok"), " step:", (d.get("step_resident") or {}).get("ms_per_step"))
    except Exception as ex:
        print(f, "unreadable", ex); print(open(f.replace(".json", ".err").replace("_nx256", "")).read()[-1500:])
PY
