#!/bin/bash
# One GPU session: tests, smoke, the bench configs, the ncu launch list and one full capture of the dominant kernel.
#   gpurun --timeout 2400 -- 'bash tools/gpu_session.sh r02b [tests] [bench] [configs] [ncu]'
TAG=${1:-sess}; shift; WHAT="${*:-tests bench configs ncu}"
OUT=gpurun_out/$TAG; mkdir -p $OUT
has() { [[ " $WHAT " == *" $1 "* ]]; }
if has tests; then
  timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
  timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
fi
if has bench; then
  timeout 900 python bench.py --steps 5 --warmup 3 2>$OUT/bench_slab512.err | tail -1 > $OUT/bench_slab512.json
fi
if has configs; then
  for cfg in cube256 cube128 ot2d dust2m; do
    timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 2>$OUT/bench_$cfg.err | tail -1 > $OUT/bench_$cfg.json
  done
fi
if has ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_slab512.csv python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rates_pair -s 6 -c 1 -o $OUT/prof_rates_slab512 python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/ncu_rates.log 2>&1
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench*.json")):
    try:
        d = json.load(open(f))
        e = d.get("e2e", {})
        print("%-28s %8.2f ms/step  e2e lean %8.2f full %8.2f ms  pair %6.2f ms  fp64 frac %s  cpu %s" % (f.split("/")[-1], d["ms_per_step"], e.get("ms_per_step", 0), e.get("full_contract", {}).get("ms_per_step", 0),
              d.get("roofline", {}).get("kernel_ms", 0), d.get("roofline", {}).get("fp64", {}).get("frac"), d.get("cpu_baseline", {}).get("value")))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
ls -la $OUT
