#!/bin/bash
# Round 2, four-GPU call with the momentum-blocked products: bench at N=4.
mkdir -p gpurun_out
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR4 --master-port 29553 bench.py --gpus 4 --steps 5 --warmup 3 --no-calibration > gpurun_out/r2_bench_n4_d.json 2> gpurun_out/r2_bench_n4_d.log
cat gpurun_out/r2_bench_n4_d.json | cut -c1-300
