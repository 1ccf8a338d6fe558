#!/bin/bash
# records that carry the source hash: the headline bench line and the full-launch capture of the pair kernel (run after the last source change)
OUT=gpurun_out/${1:-last}; mkdir -p $OUT
timeout 900 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tail -1 > $OUT/bench_nx512.json
NDSPMHD_B200_RATE_CHUNKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rates_pair -s 3 -c 1 -o $OUT/prof_rates_nx512 python bench.py --nx 512 --steps 1 --warmup 3 --no-cpu > $OUT/ncu_rates.log 2>&1
timeout 600 python -m pytest tests -q -m gpu -x -k "not large" 2>&1 | tail -3 | tee $OUT/pytest_gpu_quick.txt
python - <<PY
import json
d = json.load(open("$OUT/bench_nx512.json"))
print(round(d["ms_per_step"], 2), "ms/step  e2e", round(d["e2e"]["ms_per_step"], 2), d["phases_ms"], "traffic", d["roofline"]["traffic"], "fp64", (d["roofline"].get("fp64") or {}).get("frac"), "cpu", d.get("cpu_baseline", {}).get("value"))
PY
