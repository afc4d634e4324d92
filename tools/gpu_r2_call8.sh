#!/bin/bash
# Round 2, call 8 (one GPU): consumer loop pipelined across k-tiles -- parity + ladder / ring timings.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_h.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_h.txt
for dbg in 0 1; do
PMB_WS_DEBUG=$dbg timeout 300 python tools/profile_pp_virtual.py 20 3 2>&1 | tail -1
PMB_WS_DEBUG=$dbg timeout 300 python tools/profile_pp_virtual.py 13 5 dense 2>&1 | tail -1
done
timeout 300 python tools/profile_pp_virtual.py 25 2 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_h.json 2> gpurun_out/r2_bench_n1_h.log
cat gpurun_out/r2_bench_n1_h.json | cut -c1-260
