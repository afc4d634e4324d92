#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_eom.py 13 3 32 > gpurun_out/eom_sigma_54e_203.json 2> gpurun_out/eom_sigma_54e_203.log
cat gpurun_out/eom_sigma_54e_203.log | tail -20; cat gpurun_out/eom_sigma_54e_203.json
