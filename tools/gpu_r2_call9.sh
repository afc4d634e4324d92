#!/bin/bash
# Round 2, call 9 (one GPU): 16-byte copies of tau -- parity, ladder timing, bench, smoke.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_i.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_i.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_i.json 2> gpurun_out/r2_bench_n1_i.log
cat gpurun_out/r2_bench_n1_i.json | cut -c1-260
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1_i.json')); print('pp ladder', d['roofline']['ms_per_launch'], d['roofline']['achieved'])
PY
