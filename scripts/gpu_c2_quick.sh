python -m pytest tests/test_gpu_fused.py -q -x 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b2.json 2> gpurun_out/b2.err
python -c "import json; d=json.loads(open('gpurun_out/b2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'])"
