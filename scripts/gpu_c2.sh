python -m pytest tests/test_gpu_fused.py -q -x 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b2.json 2> gpurun_out/b2.err
python -c "import json; d=json.loads(open('gpurun_out/b2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'])"
ncu --set full --clock-control none --import-source on -k regex:k_noise_svf_scan -c 1 -o gpurun_out/prof_c2 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2_full.log 2>&1
