#!/bin/bash
# round-1 re-entry: microbenchmark + per-term profile + bench + gpu tests + launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt
timeout 120 tools/micro/dmma_issue > gpurun_out/dmma_issue.txt 2>&1
timeout 300 python tools/profile_terms.py 314 27 2 > gpurun_out/terms_314.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
tail -3 gpurun_out/pytest_gpu.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
cat gpurun_out/dmma_issue.txt gpurun_out/terms_314.txt gpurun_out/bench_r1_n1.json
