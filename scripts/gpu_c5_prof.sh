#!/bin/bash
# by-kernel profile of the four tape-specialised lane kernels of configs[4] (archetypes a, b, c, d) at 1/10 length
export QG_SPEC_MIN_WORK=1
python scripts/prof_wl.py c5 9600 > gpurun_out/c5_plain.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spec -c 8 -o gpurun_out/prof_c5 -f python scripts/prof_wl.py c5 9600 > gpurun_out/ncu_c5.log 2>&1
cat gpurun_out/c5_plain.txt
