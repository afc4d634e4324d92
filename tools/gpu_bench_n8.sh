#!/bin/bash
# round-1 call 14 (8 GPUs): the same 515-orbital workload over 8 ranks
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.log
tail -6 gpurun_out/bench_n8.log; cat gpurun_out/bench_n8.json
