#!/bin/bash
# First GPU call of the next round: broad validation of K1s (the tape-specialised kernel) before AUTO may select it.
#   gpurun --timeout 900 -- 'bash scripts/gpu_spec_round.sh'
mkdir -p gpurun_out
# 1. every uniform render case + the measured archetypes, bit-identical to the sample-by-sample interpreter (xfail/xpass listed)
python -m pytest tests/test_zz_gpu_spec.py -m gpu -q -rxX 2>&1 | tail -40 > gpurun_out/spec_sweep.txt
tail -3 gpurun_out/spec_sweep.txt
# 2. the random-graph families through the specialised kernel as well (oracle parity, 20 seeds per family)
QG_FUZZ_SPEC=1 QG_FUZZ_SEEDS=20 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/spec_fuzz.txt
tail -2 gpurun_out/spec_fuzz.txt
# 3. speed on the configs[4] archetypes: default form, then the experimental block form for delay lines
python scripts/spec_check.py 262144 9600 > gpurun_out/spec_check_default.txt 2>&1; tail -4 gpurun_out/spec_check_default.txt
cp gpurun_out/spec_check.json gpurun_out/spec_check_default.json
QG_SPEC_PREFETCH=1 python scripts/spec_check.py 262144 9600 > gpurun_out/spec_check_prefetch.txt 2>&1; tail -4 gpurun_out/spec_check_prefetch.txt
cp gpurun_out/spec_check.json gpurun_out/spec_check_prefetch.json
