#!/bin/bash
# BASELINE.json configs C1-C5 on one GPU, each with its serial-oracle CPU sample: the lines of BASELINE.md section 4
TAG=${1:-configs}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for cfgname in ot2d cube128 dust2m cube256 slab512; do
  timeout 900 python bench.py --config $cfgname --steps 10 --warmup 3 2>$OUT/bench_$cfgname.err | tail -1 > $OUT/bench_$cfgname.json
done
# C1: the reference's own CPU-sized case (1-D Brio-Wu, ~1000 particles): launch-latency bound on a GPU; timed for the record
timeout 300 python - > $OUT/c1.json 2>$OUT/c1.err <<'PY'
import json, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from ndspmhd_b200 import lib, setups
from oracle import oracle
o, p = setups.shock1d(nright=125, mhd=True, iener=3)     # Brio-Wu, ~1000 particles (src/setup_shock1D_mhd.f90 layout)
o.device_ghosts = 1; o.want_aux = 0
hot = lib.Hotpath(o, p.ndim, 0)
hot.upload(p)
for _ in range(5):
    hot.rewind(); s = hot.derivs()
torch.cuda.synchronize(); t0 = time.perf_counter()
K = 50
for _ in range(K):
    hot.rewind(); s = hot.derivs()
torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / K * 1e3
hot.close()
po = p.copy(); t0 = time.perf_counter()
for _ in range(20):
    q = p.copy(); oracle.derivs(o, q)
cpu_ms = (time.perf_counter() - t0) / 20 * 1e3
print(json.dumps({"config": "C1 1-D Brio-Wu shock tube", "npart": int(p.npart), "gpu_ms_per_derivs": ms, "gpu_updates_per_s": p.npart / (ms * 1e-3),
                  "cpu_oracle_ms_per_derivs": cpu_ms, "cpu_updates_per_s": p.npart / (cpu_ms * 1e-3), "itsdensity": s["itsdensity"]}))
PY
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/*.json")):
    try:
        d = json.load(open(f))
        if "metric" in d:
            print(f.split("/")[-1], d["run"]["npart"], round(d["ms_per_step"], 2), "ms  value %.3e" % d["value"], " e2e %.3e" % d["e2e"]["value"], " cpu", d.get("cpu_baseline", {}).get("value"), " frac", round(d["roofline"]["frac"], 4), d["roofline"].get("fp64", {}).get("frac"))
        else:
            print(d)
    except Exception as ex:
        print(f, "unreadable", ex)
PY
