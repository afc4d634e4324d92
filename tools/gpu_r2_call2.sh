#!/bin/bash
# Round 2, second GPU call (one GPU): new tests, the reworked bench line, C4 (Davidson, 10 roots).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_b.txt 2>&1
tail -3 gpurun_out/r2_pytest_gpu_b.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.log
tail -4 gpurun_out/r2_bench_n1_b.log; cat gpurun_out/r2_bench_n1_b.json
timeout 600 python tools/bench_excited.py davidson 13 10 80 > gpurun_out/r2_c4_davidson_54e_203.json 2> gpurun_out/r2_c4_203.log
tail -3 gpurun_out/r2_c4_203.log; cat gpurun_out/r2_c4_davidson_54e_203.json
timeout 900 python tools/bench_excited.py davidson 20 10 12 > gpurun_out/r2_c4_davidson_54e_389.json 2> gpurun_out/r2_c4_389.log
tail -3 gpurun_out/r2_c4_389.log; cat gpurun_out/r2_c4_davidson_54e_389.json
