#!/bin/bash
# K5 evidence on one B200: spectral tests, c4 at 1/10 length plain, launch list, one full capture of each K5 kernel with source
python -m pytest tests/test_gpu_spectral.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t_k5.log
python scripts/prof_wl.py c4 144000 > gpurun_out/c4_plain.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spectral_frames -s 40 -c 1 -o gpurun_out/prof_c4f -f python scripts/prof_wl.py c4 144000 > gpurun_out/ncu_c4f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spectral_post -s 40 -c 1 -o gpurun_out/prof_c4p -f python scripts/prof_wl.py c4 144000 > gpurun_out/ncu_c4p.log 2>&1
cat gpurun_out/t_k5.log gpurun_out/c4_plain.txt
