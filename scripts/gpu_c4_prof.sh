#!/bin/bash
# K5 / K5s evidence on one B200: c4 at 1/10 length plain, launch list, one full capture of each kernel with source
export QG_SPECTRAL_SPEC=1
python scripts/prof_wl.py c4 144000 > gpurun_out/c4_plain.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sp_frames -s 40 -c 1 -o gpurun_out/prof_c4f -f python scripts/prof_wl.py c4 144000 > gpurun_out/ncu_c4f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sp_post -s 40 -c 1 -o gpurun_out/prof_c4p -f python scripts/prof_wl.py c4 144000 > gpurun_out/ncu_c4p.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv python scripts/prof_wl.py c4 144000 > gpurun_out/ncu_c4l.log 2>&1
cat gpurun_out/c4_plain.txt
