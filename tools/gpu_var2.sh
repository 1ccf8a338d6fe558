#!/bin/bash
# Times tuning builds (ndspmhd_b200/variants/*.so) against the default on the bench workload at nx (default 256); no tests.
TAG=${1:-var}; NX=${2:-256}
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ -x tools/micro/fp64_pipe ] && [ ! -f $OUT/fp64_pipe.txt ]; then timeout 120 tools/micro/fp64_pipe > $OUT/fp64_pipe.txt 2>&1; fi
shopt -s nullglob
for so in ndspmhd_b200/libndspmhd_b200.so ndspmhd_b200/variants/*.so; do
  b=$(basename $so .so)
  NDSPMHD_B200_LIB=$PWD/$so timeout 600 python bench.py --nx $NX --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > $OUT/bench_$b.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$b.json"))
    print("%-32s ms/step %.2f  phases %s" % ("$b", d["ms_per_step"], {k:round(v,2) for k,v in d["phases_ms"].items()}))
except Exception as e:
    print("$b", "FAILED", e, open("$OUT/bench_$b.json").read()[-300:])
PY
done
