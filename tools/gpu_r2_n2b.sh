#!/bin/bash
# Round 2, final two-GPU call: NCCL parity with the final kernels, bench at N=2.
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_rank_nccl" ) > gpurun_out/r2_pytest_gpu_n2_b.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_n2_b.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.log
cat gpurun_out/r2_bench_n2_b.json | cut -c1-300
