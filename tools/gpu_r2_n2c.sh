#!/bin/bash
# Round 2, two-GPU call with the momentum-blocked products: NCCL parity, bench at N=2, per-rank profile.
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_rank_nccl" ) > gpurun_out/r2_pytest_gpu_n2_c.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_n2_c.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_c.json 2> gpurun_out/r2_bench_n2_c.log
cat gpurun_out/r2_bench_n2_c.json | cut -c1-300
timeout 600 $TR --master-port 29542 tools/profile_sweep.py 25 virtual gpurun_out/r2_sweep_profile_n2_c > /dev/null 2>&1
head -12 gpurun_out/r2_sweep_profile_n2_c_rank0.txt | cut -c1-130
