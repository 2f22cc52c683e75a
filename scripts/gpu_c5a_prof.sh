#!/bin/bash
# by-line profile of the tape-specialised lane kernel on configs[4]'s archetype a (quantised oscillator)
export QG_SPEC_MIN_WORK=1
python scripts/prof_wl.py c5 9600 > gpurun_out/c5_plain.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spec -c 1 -o gpurun_out/prof_c5a -f python scripts/prof_wl.py c5 9600 > gpurun_out/ncu_c5a.log 2>&1
cat gpurun_out/c5_plain.txt
