#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
tail -5 gpurun_out/pytest_gpu.txt
timeout 300 python tools/profile_sweep.py > gpurun_out/sweep_profile.txt 2>&1
cat gpurun_out/sweep_profile.txt
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log
cat gpurun_out/bench_n1.json
