#!/bin/bash
# First GPU session after round 1 (everything below was prepared without a GPU; see DESIGN.md "what the v5 profiles say to do next"):
#   here:   make -C ndspmhd_b200/csrc variant TAG=intcmp DEFS="-DND_SQRT_INTGUARD=1 -DND_FMAX_INT=1"
#           make -C ndspmhd_b200/csrc variant TAG=dens768 DEFS="-DND_DENS_BLOCK_FAST=768"   (two-gather rounds: slab contexts, phase-by-phase calls)
#           make -C ndspmhd_b200/csrc variant TAG=fp64lean DEFS="-DND_SQRT_INTGUARD=1 -DND_FMAX_INT=1 -DND_RATES_FUSEDR=1 -DND_TABIDX_SAT=1"
#   gpurun --timeout 900 -- 'bash tools/gpu_next_session.sh r02a'
# 1. the whole GPU suite (tests/test_gpu_vectors.py has never run on a GPU), 2. the integer-compare variant of the pair kernel against the
# default at nx=256, 3. the end-to-end number with ND_DL_REAL_ROWS, 4. (needs --gpus 2) LIGHT density rounds in slab contexts.
TAG=${1:-next}; OUT=gpurun_out/$TAG; mkdir -p $OUT
shopt -s nullglob
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
for so in ndspmhd_b200/variants/*.so; do   # parity of every variant build before its timing counts
  NDSPMHD_B200_LIB=$PWD/$so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py -q -m gpu 2>&1 | tail -3 | tee $OUT/pytest_$(basename $so .so).txt
done
for so in ndspmhd_b200/libndspmhd_b200.so ndspmhd_b200/variants/*.so; do
  NDSPMHD_B200_LIB=$PWD/$so timeout 300 python bench.py --nx 256 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > $OUT/bench256_$(basename $so .so).json
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > $OUT/bench512.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-real-rows 2>&1 | tail -1 > $OUT/bench512_real_rows.json
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  for light in 0 1; do
    NDSPMHD_B200_SLAB_LIGHT=$light timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
      tools/slab_check.py 32 2>&1 | tail -6 > $OUT/slab_check_light$light.txt
    NDSPMHD_B200_SLAB_LIGHT=$light timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -1 > $OUT/bench512_n2_light$light.json
  done
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench*.json")):
    try:
        d = json.load(open(f))
        print("%-60s %8.2f ms/step  e2e %8.2f ms  pair %6.2f ms" % (f.split("/")[-1], d["ms_per_step"], d.get("e2e", {}).get("ms_per_step", 0), d.get("roofline", {}).get("kernel_ms", 0)))
    except Exception as e:
        print(f, "unreadable", e)
PY
