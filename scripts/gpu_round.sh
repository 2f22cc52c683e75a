#!/bin/bash
# Round evidence on one B200: GPU tests, bench lines, launch list and full ncu captures of the bench kernels.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/b_c2.err
for w in c3 c4 c5; do timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/b_$w.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/b_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_noise_svf_scan -c 1 -o gpurun_out/prof_c2 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_polysynth -c 1 -o gpurun_out/prof_c3 -f python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c3_full.log 2>&1
tail -c 300 gpurun_out/b_c2.err gpurun_out/b_c3.err gpurun_out/b_c4.err gpurun_out/b_c5.err gpurun_out/b_ref.err
