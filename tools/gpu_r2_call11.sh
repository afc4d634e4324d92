#!/bin/bash
# Round 2, call 11 (one GPU): final verification -- whole -m gpu suite, smoke, the default bench line.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_j.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_j.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.log
tail -2 gpurun_out/r2_bench_n1_final.log; cat gpurun_out/r2_bench_n1_final.json | cut -c1-300
timeout 600 python tools/profile_sweep.py 25 virtual gpurun_out/r2_sweep_profile_n1_final > /dev/null 2>&1
head -22 gpurun_out/r2_sweep_profile_n1_final_rank0.txt | cut -c1-130
