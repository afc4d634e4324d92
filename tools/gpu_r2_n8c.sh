#!/bin/bash
# Round 2, final eight-GPU call: bench at N=8, C3 synthetic, then N=4 on four of the GPUs.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_n8_c.json 2> gpurun_out/r2_bench_n8_c.log
cat gpurun_out/r2_bench_n8_c.json | cut -c1-300
timeout 500 $TR --master-port 29552 bench.py --workload synthetic --gpus 8 --steps 2 --warmup 2 > gpurun_out/r2_bench_synthetic_n8_c.json 2> gpurun_out/r2_bench_synthetic_n8_c.log
cat gpurun_out/r2_bench_synthetic_n8_c.json | cut -c1-300
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR4 --master-port 29553 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2_bench_n4_c.json 2> gpurun_out/r2_bench_n4_c.log
cat gpurun_out/r2_bench_n4_c.json | cut -c1-300
