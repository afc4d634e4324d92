#!/bin/bash
# Round 2, second eight-GPU call: bench + per-rank profile after the stacked tau pair / gemv / sparse clearing.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_n8_b.json 2> gpurun_out/r2_bench_n8_b.log
cat gpurun_out/r2_bench_n8_b.json | cut -c1-300
timeout 400 $TR --master-port 29532 tools/profile_sweep.py 25 virtual gpurun_out/r2_sweep_profile_n8_b > gpurun_out/r2_sweep_profile_n8_b.log 2>&1
head -16 gpurun_out/r2_sweep_profile_n8_b_rank0.txt | cut -c1-150; grep -A12 "kernels by name" gpurun_out/r2_sweep_profile_n8_b_rank0.txt | cut -c1-150
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR4 --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2_bench_n4_b.json 2> gpurun_out/r2_bench_n4_b.log
cat gpurun_out/r2_bench_n4_b.json | cut -c1-300
