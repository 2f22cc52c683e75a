#!/bin/bash
# Full GPU validation: every -m gpu test (parity margins logged), smoke, default bench, reference arm
rm -f gpurun_out/parity_margins.jsonl
QG_PARITY_LOG=gpurun_out/parity_margins.jsonl python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/t_full.log
cat gpurun_out/t_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 300 gpurun_out/bench_full.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
wc -l gpurun_out/parity_margins.jsonl
