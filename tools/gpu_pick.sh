#!/bin/bash
# Times the candidate builds in ndspmhd_b200/variants/ on the bench workload at nx=256, picks the fastest (the plain ND_DENS_LIGHT=1
# build is the incumbent at INCUMBENT_MS), then runs the whole GPU test suite, smoke() and the full-size bench on the winner.
OUT=gpurun_out/pick; mkdir -p $OUT
INC=$PWD/ndspmhd_b200/variants/libndspmhd_b200_light.so; INC_MS=${1:-10.058}
shopt -s nullglob
for so in ndspmhd_b200/variants/*_b*.so; do
  NDSPMHD_B200_LIB=$PWD/$so timeout 60 python bench.py --nx 256 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > $OUT/bench_$(basename $so .so).json
done
WIN=$(python - <<PY
import json, glob, os
best, ms = "$INC", float("$INC_MS")
for f in sorted(glob.glob("$OUT/bench_*_b*.json")):
    try:
        d = json.load(open(f))
    except Exception:
        continue
    if d["ms_per_step"] < ms * 0.995:
        ms = d["ms_per_step"]; best = os.path.join("$PWD", "ndspmhd_b200", "variants", os.path.basename(f)[6:-5] + ".so")
print(best)
PY
)
echo "winner $WIN" | tee $OUT/winner.txt
export NDSPMHD_B200_LIB=$WIN
timeout 120 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > $OUT/bench_512.json
head -c 300 $OUT/bench_512.json
