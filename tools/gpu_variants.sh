#!/bin/bash
# Times differently tuned builds (ndspmhd_b200/variants/*.so) on the bench workload at a reduced size.
TAG=${1:-var}; NX=${2:-256}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest gpu (default build)"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
shopt -s nullglob
for so in ndspmhd_b200/libndspmhd_b200.so ndspmhd_b200/variants/*.so; do
  echo "== $so"
  NDSPMHD_B200_LIB=$PWD/$so timeout 600 python bench.py --nx $NX --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > $OUT/bench_$(basename $so .so).json
  python - <<PY
import json
d=json.load(open("$OUT/bench_$(basename $so .so).json"))
print("%-40s ms/step %.2f value %.1fM  its %d phases %s e2e %.1fM" % ("$(basename $so .so)", d["ms_per_step"], d["value"]/1e6, d["config"]["itsdensity"], {k:round(v,2) for k,v in d["phases_ms"].items()}, d["e2e"]["value"]/1e6))
PY
done
