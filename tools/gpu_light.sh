#!/bin/bash
# Checks the ND_DENS_LIGHT=1 build (ndspmhd_b200/variants/libndspmhd_b200_light.so; csrc/Makefile `variant`) against the oracle on the
# fast-tuple parity cases and the device-step tests, then times it on the bench workload at a reduced size.
OUT=gpurun_out/light; mkdir -p $OUT
export NDSPMHD_B200_LIB=$PWD/ndspmhd_b200/variants/libndspmhd_b200_light.so
timeout 90 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py -q -m gpu -k "noaux or step" 2>&1 | tail -15 > $OUT/pytest.txt
timeout 120 python bench.py --nx ${1:-256} --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > $OUT/bench_light.json
cat $OUT/pytest.txt | tail -5
