#!/bin/bash
# Round 2, third GPU call (one GPU): tail-wave split + shared T1 products, sweep profile, C4.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_c.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_c.txt
timeout 600 python tools/profile_sweep.py 25 virtual gpurun_out/r2_sweep_profile_n1 > gpurun_out/r2_sweep_profile_n1.log 2>&1
head -30 gpurun_out/r2_sweep_profile_n1_rank0.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.log
cat gpurun_out/r2_bench_n1_c.json | cut -c1-400
timeout 600 python tools/bench_excited.py davidson 13 10 80 > gpurun_out/r2_c4_davidson_54e_203.json 2> gpurun_out/r2_c4_203.log
tail -3 gpurun_out/r2_c4_203.log; cat gpurun_out/r2_c4_davidson_54e_203.json
timeout 900 python tools/bench_excited.py davidson 20 10 12 > gpurun_out/r2_c4_davidson_54e_389.json 2> gpurun_out/r2_c4_389.log
tail -3 gpurun_out/r2_c4_389.log; cat gpurun_out/r2_c4_davidson_54e_389.json
