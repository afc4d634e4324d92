#!/bin/bash
# Round 2, two-GPU call: NCCL parity, sharded sweep (bench + per-rank profile), row-sharded EOM sigma.
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
( timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_rank_nccl or tail_wave or bit_identical" ) > gpurun_out/r2_pytest_gpu_n2.txt 2>&1
tail -5 gpurun_out/r2_pytest_gpu_n2.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.log
cat gpurun_out/r2_bench_n2.json | cut -c1-300
timeout 600 $TR --master-port 29512 tools/profile_sweep.py 25 virtual gpurun_out/r2_sweep_profile_n2 > gpurun_out/r2_sweep_profile_n2.log 2>&1
head -12 gpurun_out/r2_sweep_profile_n2_rank0.txt; grep -A12 "kernels by name" gpurun_out/r2_sweep_profile_n2_rank1.txt | cut -c1-150
timeout 900 $TR --master-port 29513 tools/bench_eom_sharded.py 20 2 4 > gpurun_out/r2_eom_sigma_rows_n2_389.json 2> gpurun_out/r2_eom_rows_n2.log
tail -4 gpurun_out/r2_eom_rows_n2.log; cat gpurun_out/r2_eom_sigma_rows_n2_389.json | cut -c1-600
