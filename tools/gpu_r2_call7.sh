#!/bin/bash
# Round 2, call 7 (one GPU): sparse clearing of the generated A tiles -- tests, ladder timing, bench.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_g.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_g.txt
timeout 300 python tools/profile_pp_virtual.py 20 3 2>&1 | tail -1
timeout 300 python tools/profile_pp_virtual.py 25 2 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_g.json 2> gpurun_out/r2_bench_n1_g.log
cat gpurun_out/r2_bench_n1_g.json | cut -c1-260
