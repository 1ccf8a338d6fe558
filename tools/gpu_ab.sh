#!/bin/bash
# A/B of the default build against the builds in ndspmhd_b200/variants/: parity first, then the bench at two sizes
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
shopt -s nullglob
timeout 1500 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
for so in ndspmhd_b200/libndspmhd_b200.so ndspmhd_b200/variants/*.so; do
  b=$(basename $so .so)
  NDSPMHD_B200_LIB=$PWD/$so timeout 300 python bench.py --nx 256 --steps 5 --warmup 3 --no-cpu 2>$OUT/b256_$b.err | tail -1 > $OUT/bench256_$b.json
  NDSPMHD_B200_LIB=$PWD/$so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>$OUT/b512_$b.err | tail -1 > $OUT/bench512_$b.json
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench*.json")):
    try:
        d = json.load(open(f)); e = d.get("e2e", {})
        print("%-52s %8.2f ms/step  e2e %8.2f ms  pair %6.2f ms  fp64frac %s %s" % (f.split("/")[-1], d["ms_per_step"], e.get("ms_per_step", 0), d["roofline"]["kernel_ms"], d["roofline"].get("fp64", {}).get("frac"), d["phases_ms"]))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
