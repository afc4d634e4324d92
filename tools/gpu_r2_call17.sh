#!/bin/bash
# Round 2, call 17 (one GPU): C2 DCSD iteration with the momentum-blocked products.
mkdir -p gpurun_out
timeout 300 python bench.py --dcsd --no-cpu --no-calibration > gpurun_out/r2_bench_n1_dcsd_blocked.json 2> gpurun_out/r2_bench_n1_dcsd_blocked.log
cut -c1-300 gpurun_out/r2_bench_n1_dcsd_blocked.json
