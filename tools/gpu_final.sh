#!/bin/bash
# end-of-round session on one GPU: the suite, the headline bench (with its CPU leg), smoke, then the profiles of the same build
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tail -1 > $OUT/bench_nx512.json
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/smoke.txt
bash tools/gpu_prof.sh ${1:-final}/prof 512 > $OUT/prof.log 2>&1
python - <<PY
import json
d = json.load(open("$OUT/bench_nx512.json"))
print(round(d["ms_per_step"], 2), "ms/step  e2e", round(d["e2e"]["ms_per_step"], 2), d["phases_ms"], "fp64", d["roofline"].get("fp64"), "step", d["step_resident"], "cpu", d.get("cpu_baseline", {}).get("value"))
PY
