#!/bin/bash
# Round 2, eight-GPU call: per-rank sweep profile, bench at N=8, C3 (synthetic, stored dense V_abcd),
# C5 (FEAST contour 16 x 32 dealt out over the ranks).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 tools/profile_sweep.py 25 virtual gpurun_out/r2_sweep_profile_n8 > gpurun_out/r2_sweep_profile_n8.log 2>&1
head -40 gpurun_out/r2_sweep_profile_n8_rank0.txt | cut -c1-150; grep -A14 "kernels by name" gpurun_out/r2_sweep_profile_n8_rank7.txt | cut -c1-150
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/r2_nccl_n8_%h_%p.log timeout 400 $TR --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.log
cat gpurun_out/r2_bench_n8.json | cut -c1-300
grep -h "NVLS\|nranks\|Connected all" gpurun_out/r2_nccl_n8_*.log | sort | uniq -c | sort -rn | head -8 > gpurun_out/r2_nccl_n8_summary.txt; rm -f gpurun_out/r2_nccl_n8_*_*.log; cat gpurun_out/r2_nccl_n8_summary.txt | cut -c1-200
timeout 500 $TR --master-port 29523 bench.py --workload synthetic --gpus 8 --steps 2 --warmup 2 > gpurun_out/r2_bench_synthetic_n8.json 2> gpurun_out/r2_bench_synthetic_n8.log
tail -5 gpurun_out/r2_bench_synthetic_n8.log | cut -c1-300; cat gpurun_out/r2_bench_synthetic_n8.json | cut -c1-1500
timeout 700 $TR --master-port 29524 tools/bench_excited.py feast 13 16 32 1 -0.156 0.03 12 12 > gpurun_out/r2_c5_feast_16x32_n8_203.json 2> gpurun_out/r2_c5_n8.log
tail -5 gpurun_out/r2_c5_n8.log | cut -c1-300; cat gpurun_out/r2_c5_feast_16x32_n8_203.json | cut -c1-2500
