#!/bin/bash
# round-1 call 13 (2 GPUs): sharded CCSD with the generated V_abcd rows, same 515-orbital workload
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.log
tail -8 gpurun_out/bench_n2.log; cat gpurun_out/bench_n2.json
