#!/bin/bash
# Round 2, call 20 (one GPU): pmb_gather_expand (T1 products of the o.v^3 blocks through partner tables) -- parity, bench.
mkdir -p gpurun_out
( PYMES_B200_T1_GATHER=1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "momentum_gather or lockstep or ccsd_matches_dense" ) > gpurun_out/r2_pytest_gpu_gather.txt 2>&1
tail -4 gpurun_out/r2_pytest_gpu_gather.txt
PYMES_B200_T1_GATHER=1 timeout 200 python bench.py --no-cpu --no-calibration > gpurun_out/r2_bench_n1_gather.json 2> gpurun_out/r2_bench_n1_gather.log
cut -c1-260 gpurun_out/r2_bench_n1_gather.json
