#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 900 python tools/bench_eom.py 13 3 64 > gpurun_out/eom_sigma_54e_203.json 2> gpurun_out/eom_sigma_54e_203.log
tail -8 gpurun_out/eom_sigma_54e_203.log
