#!/bin/bash
# Round 2, first GPU call (one GPU): the CUDA cases written after round 1's GPU budget was spent.
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_next.py -q -m gpu_next ) > gpurun_out/r2_pytest_gpu_next.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_next.txt
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu.txt
timeout 400 python tools/bench_eom.py 20 2 8 virtual > gpurun_out/r2_eom_sigma_54e_389_virtual.json 2> gpurun_out/r2_eom_389.log
tail -6 gpurun_out/r2_eom_389.log
timeout 500 python tools/bench_eom.py 25 2 2 virtual > gpurun_out/r2_eom_sigma_54e_515_virtual.json 2> gpurun_out/r2_eom_515.log
tail -6 gpurun_out/r2_eom_515.log
timeout 300 python bench.py --dcsd --no-cpu --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_dcsd.json 2> gpurun_out/r2_bench_n1_dcsd.log
cat gpurun_out/r2_bench_n1_dcsd.json
